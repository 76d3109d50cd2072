/*
 * kcgpu_kernels.cu -- sm_100a kernels of the full k-mer counting mode (the kc-c4 path).
 *
 *  kc_scan_kernel<MODE>    every thread owns the 16 stream positions of one 128-bit chunk, warms
 *                          its two rolling words up on the 32 bytes before them (k - 1 <= 30),
 *                          and for each position where a k-mer ends (kc-c4.c:80-87) hashes the
 *                          canonical word (kc-c4.c:40-50).  Then
 *    KC_PARTITION          appends it to the list of its region in its owner's allocation
 *                          (count_seq_buf / c4x_insert_buf, kc-c4.c:64-90); a list that is full
 *                          sends the k-mer straight to the table instead;
 *    KC_PUSH               several owners: appends it to its owner's inbox (the owner's list area
 *                          used as one list).  The CTA counts its k-mers per owner in shared
 *                          memory and reserves one range per owner with one atomic on the
 *                          owner's cursor, so a warp's k-mers for one owner are neighbours and
 *                          leave as whole sectors -- over NVLink when the owner is a peer;
 *    KC_DIRECT             adds it to its owner's table with 64-bit compare-and-swap;
 *    KC_EXTRACT            appends it to one list per owner (warp-aggregated), for an exchange
 *                          by NCCL all-to-all.
 *                          In the first two the owner's memory may be a peer's: the same
 *                          instructions then travel over NVLink, which is the all-to-all of
 *                          the partition step fused into the producer.
 *  kc_route_kernel         several owners: what arrived in the inbox, filed under its region
 *  kc_flush_kernel         worker_for (kc-c4.c:116-128): the region lists into the table, region
 *                          by region so that the slice being filled stays in L2
 *  kc_insert_kernel        the same for lists that came from an exchange
 *  kc_hist_kernel          worker_hist (kc-c4.c:186-197) over the slots
 *
 * No tensor cores: shifts, multiplies and atomics.
 */
#include "kcgpu_kernels.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace kcgpu {

namespace {

#define KC_FULL 0xFFFFFFFFu
#define KC_THREADS 256

/* A C G T U in either case: bit (b & 31) of this word, for bytes 0x40..0x7F */
#define KC_BASE_BITS ((1u << 1) | (1u << 3) | (1u << 7) | (1u << 20) | (1u << 21))

__device__ __forceinline__ bool kc_is_base(uint32_t b)
{
	return ((b & 0xC0u) == 0x40u) && ((KC_BASE_BITS >> (b & 31u)) & 1u);
}

__device__ __forceinline__ uint64_t ld_slot(const uint64_t *p)
{
	/* L2 is the point of coherence for the atomics; a stale value is harmless because the
	 * compare-and-swap that follows is the arbiter (a slot never changes its tag) */
	return __ldcg(reinterpret_cast<const unsigned long long *>(p));
}

__device__ __forceinline__ uint64_t cas_slot(uint64_t *p, uint64_t expect, uint64_t want)
{
	return atomicCAS(reinterpret_cast<unsigned long long *>(p), (unsigned long long)expect, (unsigned long long)want);
}

enum { KC_PARTITION = 0, KC_DIRECT = 1, KC_EXTRACT = 2, KC_PUSH = 3 };

__device__ __forceinline__ uint64_t kc_home(uint64_t tag, uint32_t rslot_bits) { return (tag * 0x9E3779B97F4A7C15ull) >> (64 - rslot_bits); }

/* The Bloom pre-filter of yak-count's first pass (yak-count.c:71-104,159-161): has this k-mer
 * been seen before?  One 64-bit word per k-mer and ONE atomic: the word comes from the low bits
 * of the tag, the n bit positions from the bits above them (double hashing with an odd step), and
 * the k-mer counts as seen when every one of its bits was set before.  (The reference spreads
 * the bits over a 512-bit block and sets them one by one, which is exact only because one thread
 * owns a partition; with one atomic per k-mer two occurrences of a k-mer can never both find it
 * new-ish, however they race -- the second one sees all the bits of the first.  Which k-mers end
 * up with an entry differs from the reference only by false positives, and those are dropped
 * again by the shrink, yak-count.c:453.)  `slice` is the region's part of the filter,
 * 2^slice_bits bits. */
__device__ __forceinline__ bool kc_bloom_seen(uint32_t *slice, uint32_t slice_bits, uint32_t n_hashes, uint64_t tag)
{
	const uint32_t words_log = slice_bits - 6;
	unsigned long long *w = reinterpret_cast<unsigned long long *>(slice) + (tag & ((1ull << words_log) - 1ull));
	const uint32_t first = (uint32_t)(tag >> words_log) & 63u, step = ((uint32_t)(tag >> (words_log + 6)) & 63u) | 1u;
	unsigned long long mask = 0;
	for (uint32_t i = 0, z = first; i < n_hashes; ++i, z = (z + step) & 63u) mask |= 1ull << z;
	return (atomicOr(w, mask) & mask) == mask;
}

/* the region's slice of an owner's Bloom filter, NULL if there is no filter (or none that leaves
 * a region at least one word: every k-mer then gets an entry, as when the reference cannot
 * create its filter, yak-count.c:75,117-121) */
__device__ __forceinline__ uint32_t *kc_bloom_slice(uint32_t *bloom, const InsertCtl &ctl, uint32_t region_bits, uint64_t region)
{
	if (!bloom || ctl.bloom_bits < region_bits + 6) return nullptr;
	return bloom + (region << (ctl.bloom_bits - region_bits - 5));
}

/* kc-c4.c:116-128 / yak-count.c:150-177 on one region of one table: find (or, unless the mode
 * is KC_INS_LOOKUP, claim) the slot of `tag` and count up to 1023 (KC_INS_CLAIM: make the entry,
 * leave its count at 0).  `v` is what the home slot `pos` held when the caller looked (callers
 * look at several home slots before they resolve the first, so that the loads overlap). */
__device__ __forceinline__ void kc_insert_from(uint64_t *base, uint32_t rslot_bits, uint64_t tag, uint64_t pos, uint64_t v,
                                               const int mode, uint32_t &n_new, uint32_t &n_overflow)
{
	const uint64_t rmask = (1ull << rslot_bits) - 1ull, key = tag + 1ull;
	for (int tries = 0; tries < KC_MAX_PROBES; ++tries) {
		uint64_t *p = base + pos;
		if (tries) v = ld_slot(p);
		if (v == 0) {
			if (mode == KC_INS_LOOKUP) return; /* no entry: not counted */
			v = cas_slot(p, 0, (key << KC_COUNT_BITS) | (mode == KC_INS_CLAIM ? 0ull : 1ull));
			if (v == 0) {
				++n_new;
				return;
			}
		}
		while ((v >> KC_COUNT_BITS) == key) {
			if (mode == KC_INS_CLAIM) return;
			if ((v & KC_COUNT_MAX) == KC_COUNT_MAX) return; /* kc-c4.c:125 */
			const uint64_t old = cas_slot(p, v, v + 1);
			if (old == v) return;
			v = old;
		}
		pos = (pos + 1) & rmask;
	}
	if (mode != KC_INS_LOOKUP) ++n_overflow;
}

__device__ __forceinline__ void kc_insert_region(uint64_t *base, uint32_t rslot_bits, uint64_t tag, uint32_t *bloom_slice,
                                                 const InsertCtl &ctl, uint32_t slice_bits, uint32_t &n_new, uint32_t &n_overflow)
{
	if (ctl.mode == KC_INS_CLAIM && bloom_slice && !kc_bloom_seen(bloom_slice, slice_bits, ctl.bloom_hashes, tag)) return;
	const uint64_t pos = kc_home(tag, rslot_bits);
	kc_insert_from(base, rslot_bits, tag, pos, ld_slot(base + pos), ctl.mode, n_new, n_overflow);
}

__device__ __forceinline__ void kc_insert(uint64_t *table, uint32_t region_bits, uint32_t rslot_bits, uint64_t q, uint32_t *bloom,
                                          const InsertCtl &ctl, uint32_t &n_new, uint32_t &n_overflow)
{
	const uint64_t region = q & ((1ull << region_bits) - 1ull);
	kc_insert_region(table + (region << rslot_bits), rslot_bits, q >> region_bits, kc_bloom_slice(bloom, ctl, region_bits, region), ctl,
	                 ctl.bloom_bits - region_bits, n_new, n_overflow);
}

__device__ __forceinline__ void kc_owner(uint64_t h, uint32_t n_parts, int part_shift, uint32_t &owner, uint64_t &q)
{
	if (part_shift >= 0) {
		owner = (uint32_t)h & (n_parts - 1u);
		q = h >> part_shift;
	} else {
		q = h / n_parts;
		owner = (uint32_t)(h - q * n_parts);
	}
}

__device__ __forceinline__ uint64_t kc_unsplit(uint64_t q, uint32_t owner, uint32_t n_parts, int part_shift)
{
	return part_shift >= 0 ? (q << part_shift) | owner : q * n_parts + owner;
}

template <int MODE>
__global__ void __launch_bounds__(KC_THREADS) kc_scan_kernel(const CountArgs a, const int part_shift)
{
	const uint64_t c = a.first_chunk + (uint64_t)blockIdx.x * KC_THREADS + threadIdx.x;
	uint32_t n_kmers = 0, n_new = 0, n_overflow = 0, n_dropped = 0, n_direct = 0;
	const bool live = c < a.end_chunk;
	__shared__ uint32_t s_cnt[2][KC_MAX_PARTS];
	__shared__ unsigned long long s_base[2][KC_MAX_PARTS];
	if (MODE == KC_PUSH) {
		if (threadIdx.x < KC_MAX_PARTS) s_cnt[0][threadIdx.x] = 0;
		__syncthreads();
	}
	if (live || MODE == KC_PUSH) { /* KC_PUSH: the whole CTA walks in step (barriers below) */
		const uint4 *chunks = reinterpret_cast<const uint4 *>(a.bytes);
		const uint4 sep = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
		const uint4 w0 = live && c >= 2 ? __ldg(chunks + c - 2) : sep;
		const uint4 w1 = live && c >= 1 ? __ldg(chunks + c - 1) : sep;
		uint4 own = live ? __ldg(chunks + c) : sep;
		const int k = a.k;
		const uint64_t mask = (1ull << 2 * k) - 1ull;
		const int top = 2 * (k - 1);
		uint64_t fw = 0, rv = 0;
		int run = 0;
		/* warm up: after these 32 bytes run, fw and rv are what a scan from the start of the
		 * read would hold wherever a k-mer can end inside the chunk (run is capped by the
		 * window, and k - 1 <= 30 < 32) */
		const uint32_t warm[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
		for (int i = 0; i < 32; ++i) {
			const uint32_t b = (warm[i >> 2] >> (8 * (i & 3))) & 0xFFu;
			uint64_t code = (b >> 1) & 3u;
			code ^= code >> 1; /* A0 C1 T2 G3 -> A0 C1 G2 T3 (kc-c4.c:21-38) */
			fw = (fw << 2 | code) & mask;
			rv = rv >> 2 | (3ull - code) << top;
			run = kc_is_base(b) ? run + 1 : 0;
		}
		/* four positions at a time: the push form pays one pair of barriers and one cursor
		 * atomic per owner for four positions of every thread; for the other forms four
		 * independent round trips per thread measured the same as one (the list stores are
		 * throughput-bound, profiles/r1_kc_ablation.txt) */
#pragma unroll 1
		for (int g = 0; g < 4; ++g) {
			uint32_t word = own.x;
			own.x = own.y, own.y = own.z, own.z = own.w;
			uint64_t q[4];
			uint32_t owner[4];
			bool ok[4];
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const uint32_t b = word & 0xFFu;
				word >>= 8;
				uint64_t code = (b >> 1) & 3u;
				code ^= code >> 1;
				fw = (fw << 2 | code) & mask;
				rv = rv >> 2 | (3ull - code) << top;
				run = kc_is_base(b) ? run + 1 : 0;
				ok[j] = run >= k;
				kc_owner(kc_hash64(fw < rv ? fw : rv, mask), a.n_parts, part_shift, owner[j], q[j]);
				n_kmers += ok[j];
			}
			if (MODE == KC_PARTITION) {
				/* file q under its region in the owner's allocation; the cursor runs on past the
				 * capacity so that the flush knows the list was full, the excess goes to the table */
				uint64_t at[4];
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const uint64_t region = q[j] & ((1ull << a.region_bits) - 1ull);
					unsigned long long *cursor =
					    kc_cursors_of(a.tables[owner[j]], a.n_slots, a.list_cap, a.region_bits) + region * KC_CURSOR_STRIDE;
					at[j] = ok[j] ? atomicAdd(cursor, 1ull) : 0ull;
				}
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					if (!ok[j]) continue;
					uint64_t *base = a.tables[owner[j]];
					const uint64_t region = q[j] & ((1ull << a.region_bits) - 1ull);
					if (at[j] < a.list_cap) {
						kc_lists_of(base, a.n_slots)[region * a.list_cap + at[j]] = q[j];
					} else {
						++n_direct;
						kc_insert(base, a.region_bits, a.rslot_bits, q[j], kc_bloom_of(base, a.n_slots, a.list_cap, a.region_bits), a.ctl,
						          n_new, n_overflow);
					}
				}
			} else if (MODE == KC_PUSH) {
				const int buf = g & 1;
				uint32_t idx[4];
#pragma unroll
				for (int j = 0; j < 4; ++j) idx[j] = ok[j] ? atomicAdd(&s_cnt[buf][owner[j]], 1u) : 0u;
				__syncthreads();
				if (threadIdx.x < a.n_parts) {
					const uint32_t n = s_cnt[buf][threadIdx.x];
					if (n) s_base[buf][threadIdx.x] =
						atomicAdd(kc_inbox_cursor(a.tables[threadIdx.x], a.n_slots, a.list_cap, a.region_bits), (unsigned long long)n);
					s_cnt[buf ^ 1][threadIdx.x] = 0;
				}
				__syncthreads();
				const uint64_t cap = kc_inbox_cap(a.list_cap, a.region_bits);
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					if (!ok[j]) continue;
					uint64_t *base = a.tables[owner[j]];
					const uint64_t pos = s_base[buf][owner[j]] + idx[j];
					if (pos < cap) {
						kc_lists_of(base, a.n_slots)[pos] = q[j];
					} else { /* inbox full: straight to the owner's table */
						++n_direct;
						kc_insert(base, a.region_bits, a.rslot_bits, q[j], kc_bloom_of(base, a.n_slots, a.list_cap, a.region_bits), a.ctl,
						          n_new, n_overflow);
					}
				}
			} else if (MODE == KC_DIRECT) {
#pragma unroll
				for (int j = 0; j < 4; ++j)
					if (ok[j])
						kc_insert(a.tables[owner[j]], a.region_bits, a.rslot_bits, q[j],
						          kc_bloom_of(a.tables[owner[j]], a.n_slots, a.list_cap, a.region_bits), a.ctl, n_new, n_overflow);
			} else {
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					if (!ok[j]) continue;
					/* one atomic per owner and warp: the lanes that have a k-mer for the same owner
					 * reserve consecutive entries of its list */
					cg::coalesced_group active = cg::coalesced_threads();
					cg::coalesced_group same = cg::labeled_partition(active, owner[j]);
					uint32_t at = 0;
					if (same.thread_rank() == 0) at = atomicAdd(a.part_counts + owner[j], same.size());
					at = same.shfl(at, 0) + same.thread_rank();
					if (at < a.cap_per_part) a.out_keys[(uint64_t)owner[j] * a.cap_per_part + at] = kc_unsplit(q[j], owner[j], a.n_parts, part_shift);
					else ++n_dropped;
				}
			}
		}
	}
	for (int o = 16; o; o >>= 1) {
		n_kmers += __shfl_xor_sync(KC_FULL, n_kmers, o);
		n_new += __shfl_xor_sync(KC_FULL, n_new, o);
		n_overflow += __shfl_xor_sync(KC_FULL, n_overflow, o);
		n_dropped += __shfl_xor_sync(KC_FULL, n_dropped, o);
		n_direct += __shfl_xor_sync(KC_FULL, n_direct, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (n_kmers) atomicAdd(a.stats + KC_ST_KMERS, (unsigned long long)n_kmers);
		if (n_new) atomicAdd(a.stats + KC_ST_NEW, (unsigned long long)n_new);
		if (n_overflow) atomicAdd(a.stats + KC_ST_OVERFLOW, (unsigned long long)n_overflow);
		if (n_dropped) atomicAdd(a.stats + KC_ST_DROPPED, (unsigned long long)n_dropped);
		if (n_direct) atomicAdd(a.stats + KC_ST_DIRECT, (unsigned long long)n_direct);
	}
}

/* The region lists into the table.  CTA b takes tile b % tiles_per_region of region
 * b / tiles_per_region: CTAs are dispatched in order, so the resident ones work on two or
 * three neighbouring regions and their slices (<= 16 MiB each) stay in L2 while they fill. */
__global__ void __launch_bounds__(KC_THREADS) kc_flush_kernel(uint64_t *base, const uint64_t *lists, const unsigned long long *cursors,
                                                               const uint64_t list_cap, const uint32_t region_bits,
                                                               const uint32_t rslot_bits, const uint32_t tiles_per_region,
                                                               const uint32_t tile_entries, uint32_t *bloom, const InsertCtl ctl,
                                                               unsigned long long *stats)
{
	const uint32_t region = blockIdx.x / tiles_per_region, tile = blockIdx.x % tiles_per_region;
	const uint64_t filled = cursors[(uint64_t)region * KC_CURSOR_STRIDE];
	const uint64_t n = filled < list_cap ? filled : list_cap;
	const uint64_t lo = (uint64_t)tile * tile_entries;
	if (lo >= n) return;
	const uint64_t hi = lo + tile_entries < n ? lo + tile_entries : n;
	const uint64_t *list = lists + (uint64_t)region * list_cap;
	uint64_t *slice = base + ((uint64_t)region << rslot_bits);
	uint32_t *const bloom_slice = kc_bloom_slice(bloom, ctl, region_bits, region); /* stays in L2 beside the table slice */
	uint32_t n_new = 0, n_overflow = 0;
	/* one entry per thread and step: more entries in flight per thread (four home slots
	 * requested at once) measured slower -- the compare-and-swap chain, not the first load, is
	 * what a thread waits for, and the registers halve the resident threads */
	for (uint64_t i = lo + threadIdx.x; i < hi; i += KC_THREADS) {
		const uint64_t q = __ldcs(reinterpret_cast<const unsigned long long *>(list + i));
		kc_insert_region(slice, rslot_bits, q >> region_bits, bloom_slice, ctl, ctl.bloom_bits - region_bits, n_new, n_overflow);
	}
	for (int o = 16; o; o >>= 1) {
		n_new += __shfl_xor_sync(KC_FULL, n_new, o);
		n_overflow += __shfl_xor_sync(KC_FULL, n_overflow, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (n_new) atomicAdd(stats + KC_ST_NEW, (unsigned long long)n_new);
		if (n_overflow) atomicAdd(stats + KC_ST_OVERFLOW, (unsigned long long)n_overflow);
	}
}

__global__ void __launch_bounds__(KC_THREADS) kc_insert_kernel(const InsertArgs a, const int part_shift)
{
	uint32_t n_new = 0, n_overflow = 0, n_kmers = 0;
	const uint64_t stride = (uint64_t)gridDim.x * KC_THREADS;
	for (uint64_t i = (uint64_t)blockIdx.x * KC_THREADS + threadIdx.x; i < a.n; i += stride) {
		uint64_t q = __ldcs(reinterpret_cast<const unsigned long long *>(a.hashed + i));
		uint32_t owner;
		kc_owner(q, a.n_parts, part_shift, owner, q);
		++n_kmers;
		kc_insert(a.table, a.region_bits, a.rslot_bits, q, kc_bloom_of(a.table, a.n_slots, a.list_cap, a.region_bits), a.ctl, n_new, n_overflow);
	}
	for (int o = 16; o; o >>= 1) {
		n_kmers += __shfl_xor_sync(KC_FULL, n_kmers, o);
		n_new += __shfl_xor_sync(KC_FULL, n_new, o);
		n_overflow += __shfl_xor_sync(KC_FULL, n_overflow, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (n_kmers) atomicAdd(a.stats + KC_ST_KMERS, (unsigned long long)n_kmers);
		if (n_new) atomicAdd(a.stats + KC_ST_NEW, (unsigned long long)n_new);
		if (n_overflow) atomicAdd(a.stats + KC_ST_OVERFLOW, (unsigned long long)n_overflow);
	}
}

/* several owners: the inbox into the region lists (c4x_insert_buf, kc-c4.c:64-72, on the owner's side) */
__global__ void __launch_bounds__(KC_THREADS) kc_route_kernel(const RouteArgs a)
{
	uint32_t n_new = 0, n_overflow = 0, n_direct = 0;
	const uint64_t stride = (uint64_t)gridDim.x * KC_THREADS;
	const uint64_t filled = *a.n_ptr;
	const uint64_t n = filled < a.inbox_cap ? filled : a.inbox_cap;
	for (uint64_t i = (uint64_t)blockIdx.x * KC_THREADS + threadIdx.x; i < n; i += stride) {
		const uint64_t q = __ldcs(reinterpret_cast<const unsigned long long *>(a.inbox + i));
		const uint64_t region = q & ((1ull << a.region_bits) - 1ull);
		const uint64_t at = atomicAdd(a.cursors + region * KC_CURSOR_STRIDE, 1ull);
		if (at < a.cap) {
			a.lists[region * a.cap + at] = q;
		} else {
			++n_direct;
			kc_insert(a.table, a.region_bits, a.rslot_bits, q, a.bloom, a.ctl, n_new, n_overflow);
		}
	}
	for (int o = 16; o; o >>= 1) {
		n_new += __shfl_xor_sync(KC_FULL, n_new, o);
		n_overflow += __shfl_xor_sync(KC_FULL, n_overflow, o);
		n_direct += __shfl_xor_sync(KC_FULL, n_direct, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (n_new) atomicAdd(a.stats + KC_ST_NEW, (unsigned long long)n_new);
		if (n_overflow) atomicAdd(a.stats + KC_ST_OVERFLOW, (unsigned long long)n_overflow);
		if (n_direct) atomicAdd(a.stats + KC_ST_DIRECT, (unsigned long long)n_direct);
	}
}

/* kc-c4.c:186-197: bin min(count, 255) of every used slot.  One private histogram per warp in
 * shared memory; lanes that hit the same bin are merged first (most used slots of a read set
 * share a handful of counts). */
__global__ void __launch_bounds__(KC_THREADS) kc_hist_kernel(const uint64_t *table, const uint64_t n_slots,
                                                              unsigned long long *hist)
{
	__shared__ uint32_t sh[KC_THREADS / 32][256];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int i = lane; i < 256; i += 32) sh[warp][i] = 0;
	__syncwarp();
	const ulonglong2 *pairs = reinterpret_cast<const ulonglong2 *>(table);
	const uint64_t n_pairs = n_slots >> 1, stride = (uint64_t)gridDim.x * KC_THREADS;
	for (uint64_t i0 = (uint64_t)blockIdx.x * KC_THREADS; i0 < n_pairs; i0 += stride) {
		const uint64_t i = i0 + threadIdx.x;
		ulonglong2 v = make_ulonglong2(0, 0);
		if (i < n_pairs) v = __ldcs(pairs + i);
#pragma unroll
		for (int j = 0; j < 2; ++j) {
			const uint64_t s = j ? v.y : v.x;
			const uint32_t cnt = (uint32_t)s & KC_COUNT_MAX;
			const uint32_t bin = s ? (cnt < 255u ? cnt : 255u) : 0u; /* bin 0 = free: not reported */
			const uint32_t peers = __match_any_sync(KC_FULL, bin);
			if (bin && lane == __ffs(peers) - 1) {
				sh[warp][bin] += __popc(peers);
				if (sh[warp][bin] >= 0x80000000u) { /* flush before 32 bits run out */
					atomicAdd(hist + bin, (unsigned long long)sh[warp][bin]);
					sh[warp][bin] = 0;
				}
			}
			__syncwarp();
		}
	}
	__syncwarp();
	for (int i = lane; i < 256; i += 32)
		if (sh[warp][i]) atomicAdd(hist + i, (unsigned long long)sh[warp][i]);
	if (n_slots & 1) { /* cannot happen for a power of two above 1; kept for completeness */
		if (blockIdx.x == 0 && threadIdx.x == 0) {
			const uint64_t s = table[n_slots - 1];
			const uint32_t cnt = (uint32_t)s & KC_COUNT_MAX;
			if (s) atomicAdd(hist + (cnt < 255u ? cnt : 255u), 1ull);
		}
	}
}

/* yak-count.c:205-239: one bin per count, 0..1023, over the used slots.  Private histograms
 * per warp would take 32 KB of shared memory: one per CTA, lanes that hit the same bin merged
 * first (most used slots of a read set share a handful of counts). */
__global__ void __launch_bounds__(KC_THREADS) kc_hist1024_kernel(const uint64_t *table, const uint64_t n_slots,
                                                                  unsigned long long *hist)
{
	__shared__ unsigned long long sh[1024];
	for (int i = threadIdx.x; i < 1024; i += KC_THREADS) sh[i] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const uint64_t stride = (uint64_t)gridDim.x * KC_THREADS;
	for (uint64_t i0 = (uint64_t)blockIdx.x * KC_THREADS; i0 < n_slots; i0 += stride) {
		const uint64_t i = i0 + threadIdx.x;
		const uint64_t s = i < n_slots ? __ldcs(reinterpret_cast<const unsigned long long *>(table + i)) : 0ull;
		const uint32_t bin = s ? (uint32_t)s & KC_COUNT_MAX : 0xFFFFFFFFu; /* free slots are not entries */
		const uint32_t peers = __match_any_sync(KC_FULL, bin);
		if (s && lane == __ffs(peers) - 1) atomicAdd(&sh[bin], (unsigned long long)__popc(peers));
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 1024; i += KC_THREADS)
		if (sh[i]) atomicAdd(hist + i, sh[i]);
}

int shift_of(uint32_t n_parts)
{
	if (n_parts & (n_parts - 1)) return -1;
	int s = 0;
	while ((1u << s) < n_parts) ++s;
	return s;
}

} // namespace

template <int MODE>
static cudaError_t launch_scan(const CountArgs &a, cudaStream_t stream)
{
	if (a.end_chunk <= a.first_chunk) return cudaSuccess;
	const uint64_t blocks = (a.end_chunk - a.first_chunk + KC_THREADS - 1) / KC_THREADS;
	if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
	kc_scan_kernel<MODE><<<(unsigned)blocks, KC_THREADS, 0, stream>>>(a, shift_of(a.n_parts));
	return cudaGetLastError();
}

cudaError_t launch_count(const CountArgs &a, cudaStream_t stream) { return launch_scan<KC_DIRECT>(a, stream); }
cudaError_t launch_partition(const CountArgs &a, cudaStream_t stream) { return launch_scan<KC_PARTITION>(a, stream); }
cudaError_t launch_extract(const CountArgs &a, cudaStream_t stream) { return launch_scan<KC_EXTRACT>(a, stream); }
cudaError_t launch_push(const CountArgs &a, cudaStream_t stream) { return launch_scan<KC_PUSH>(a, stream); }

cudaError_t launch_flush(uint64_t *base, const uint64_t *lists, const unsigned long long *cursors, uint64_t list_cap,
                         uint32_t region_bits, uint32_t rslot_bits, uint32_t *bloom, const InsertCtl &ctl,
                         unsigned long long *stats, cudaStream_t stream)
{
	if (!list_cap) return cudaSuccess;
	const uint32_t tile = KC_FLUSH_TILE;
	const uint64_t tiles = (list_cap + tile - 1) / tile;
	const uint64_t blocks = tiles << region_bits;
	if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
	kc_flush_kernel<<<(unsigned)blocks, KC_THREADS, 0, stream>>>(base, lists, cursors, list_cap, region_bits, rslot_bits, (uint32_t)tiles, tile, bloom,
	                                                             ctl, stats);
	return cudaGetLastError();
}

cudaError_t launch_route(const RouteArgs &a, int n_sm, cudaStream_t stream)
{
	if (!a.inbox_cap) return cudaSuccess;
	kc_route_kernel<<<(unsigned)(n_sm > 0 ? n_sm : 1) * 8, KC_THREADS, 0, stream>>>(a); /* 8 CTAs of 256 threads per SM */
	return cudaGetLastError();
}

cudaError_t launch_insert(const InsertArgs &a, int n_sm, cudaStream_t stream)
{
	if (a.n == 0) return cudaSuccess;
	uint64_t blocks = (a.n + KC_THREADS - 1) / KC_THREADS;
	const uint64_t most = (uint64_t)(n_sm > 0 ? n_sm : 1) * 64;
	if (blocks > most) blocks = most;
	kc_insert_kernel<<<(unsigned)blocks, KC_THREADS, 0, stream>>>(a, shift_of(a.n_parts));
	return cudaGetLastError();
}

cudaError_t launch_histogram(const uint64_t *table, uint64_t n_slots, unsigned long long *hist256, int n_sm, cudaStream_t stream)
{
	uint64_t blocks = ((n_slots >> 1) + KC_THREADS - 1) / KC_THREADS;
	const uint64_t resident = (uint64_t)(n_sm > 0 ? n_sm : 1) * 8; /* 8 CTAs of 256 threads per SM */
	if (blocks > resident) blocks = resident;
	if (blocks < 1) blocks = 1;
	kc_hist_kernel<<<(unsigned)blocks, KC_THREADS, 0, stream>>>(table, n_slots, hist256);
	return cudaGetLastError();
}

cudaError_t launch_histogram1024(const uint64_t *table, uint64_t n_slots, unsigned long long *hist1024, int n_sm, cudaStream_t stream)
{
	uint64_t blocks = (n_slots + KC_THREADS - 1) / KC_THREADS;
	const uint64_t resident = (uint64_t)(n_sm > 0 ? n_sm : 1) * 8;
	if (blocks > resident) blocks = resident;
	if (blocks < 1) blocks = 1;
	kc_hist1024_kernel<<<(unsigned)blocks, KC_THREADS, 0, stream>>>(table, n_slots, hist1024);
	return cudaGetLastError();
}

} // namespace kcgpu
