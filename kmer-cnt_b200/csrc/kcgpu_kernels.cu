/*
 * kcgpu_kernels.cu -- sm_100a kernels of the full k-mer counting mode (the kc-c4 path).
 *
 *  Every scanning thread owns the 16 stream positions of one 128-bit chunk and reads the k-mers that end
 *  there (kc-c4.c:80-87) out of one packed 48-byte window -- the chunk and the 32 bytes before it, k - 1 <= 30 --
 *  without a byte loop (kc_extract16, kcgpu_kernels.cuh); canonical word, hash64 (kc-c4.c:40-50).  Then
 *
 *  kc_scan_tile_kernel     one owner: a CTA sorts its tile of 512 chunks (up to 8 192 k-mers) by region in shared
 *                          memory and appends every region's share to its list as ONE run -- one cursor atomic,
 *                          neighbouring stores (count_seq_buf / c4x_insert_buf, kc-c4.c:64-90); a list that is
 *                          full sends the k-mer straight to the table instead.  2.5 x the rate of one cursor
 *                          atomic and one 8-byte store per k-mer (round 1)
 *  kc_push_tile_kernel     several owners: the tile is sorted by owner and every owner's share appended to its
 *                          inbox (the owner's list area used as one list) as one run with one atomic on the
 *                          owner's cursor -- peer memory over NVLink for the other GPUs: the all-to-all of the
 *                          partition step fused into the producer
 *  kc_route_tile_kernel    several owners: what arrived in the inbox, filed under its region through the same tiles
 *  kc_scan_kernel<MODE>    KC_DIRECT: no lists, every k-mer added to its owner's table with 64-bit
 *                          compare-and-swap; KC_EXTRACT: appended to one list per owner (warp-aggregated), for
 *                          an exchange by NCCL all-to-all
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

enum { KC_DIRECT = 1, KC_EXTRACT = 2 };

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
	/* (claiming first and looking afterwards -- one L2 operation for a new k-mer instead of two -- is slower: the
	 * k-mers that are there already then cost two atomics instead of a load and one; profiles/r2_kc_ablation.txt) */
	kc_insert_from(base, rslot_bits, tag, pos, ld_slot(base + pos), ctl.mode, n_new, n_overflow);
}

__device__ __forceinline__ void kc_insert(uint64_t *table, uint32_t region_bits, uint32_t rslot_bits, uint64_t q, uint32_t *bloom,
                                          const InsertCtl &ctl, uint32_t &n_new, uint32_t &n_overflow)
{
	const uint64_t region = q & ((1ull << region_bits) - 1ull);
	kc_insert_region(table + (region << rslot_bits), rslot_bits, q >> region_bits, kc_bloom_slice(bloom, ctl, region_bits, region), ctl,
	                 ctl.bloom_bits - region_bits, n_new, n_overflow);
}

/* kc_insert out of line: the rare way out of a tile (a list that is full), kept away from the tile kernels' registers,
 * and the insert of the list-less form */
__device__ __noinline__ void kc_insert_slow(uint64_t *table, uint32_t region_bits, uint32_t rslot_bits, uint64_t q, uint32_t *bloom, int mode,
                                            uint32_t bloom_bits, uint32_t bloom_hashes, uint32_t *n_new, uint32_t *n_overflow)
{
	const InsertCtl ctl{mode, bloom_bits, bloom_hashes};
	kc_insert(table, region_bits, rslot_bits, q, bloom, ctl, *n_new, *n_overflow);
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
	uint32_t n_new = 0, n_overflow = 0, n_dropped = 0, n_direct = 0;
	uint64_t h[KC_TILE_N];
	const uint32_t ok = kc_extract16(reinterpret_cast<const uint4 *>(a.bytes), c, a.end_chunk, kc_extract_of(a.k), h);
	uint32_t n_kmers = __popc(ok);
	/* four positions at a time: their compare-and-swap chains overlap, and the code of the insert is there four
	 * times, not sixteen (all sixteen inline: 3 x slower; one at a time out of line: 1.3 x slower) */
#pragma unroll 1
	for (int g = 0; g < KC_TILE_N; g += 4) {
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if (!(ok >> (g + j) & 1u)) continue;
			uint32_t owner;
			uint64_t q;
			kc_owner(h[g + j], a.n_parts, part_shift, owner, q);
			if (MODE == KC_DIRECT) {
				uint64_t *base = a.tables[owner];
				kc_insert(base, a.region_bits, a.rslot_bits, q, kc_bloom_of(base, a.n_slots, a.list_cap, a.region_bits), a.ctl, n_new, n_overflow);
			} else {
				/* one atomic per owner and warp: the lanes that have a k-mer for the same owner
				 * reserve consecutive entries of its list */
				cg::coalesced_group active = cg::coalesced_threads();
				cg::coalesced_group same = cg::labeled_partition(active, owner);
				uint32_t at = 0;
				if (same.thread_rank() == 0) at = atomicAdd(a.part_counts + owner, same.size());
				at = same.shfl(at, 0) + same.thread_rank();
				if (at < a.cap_per_part) a.out_keys[(uint64_t)owner * a.cap_per_part + at] = kc_unsplit(q, owner, a.n_parts, part_shift);
				else ++n_dropped;
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

/* ---- filing a tile of k-mers under their regions through shared memory ----
 *
 * 8-byte stores scattered over thousands of region lists are what the first scan waited for
 * (~30 G/s, tools/exp/kc_store_exp.cu: scatter8); whole runs leave at ~200 G/s (part).  So a CTA
 * collects a tile of KC_TILE_THREADS x KC_TILE_N k-mers, ranks them by region with one
 * shared-memory atomic each, sorts the tile by region in shared memory, reserves ONE run per
 * region with one atomic on the region's cursor, and writes the sorted tile out: neighbouring
 * threads write neighbouring entries of the same list.  What a full list cannot take goes
 * straight to the table, as before. */
#define KC_TILE_THREADS 512
#ifndef KC_TILE_MIN_CTAS
#define KC_TILE_MIN_CTAS 2
#endif
enum { KC_TILE_ENTRIES = KC_TILE_THREADS * KC_TILE_N };
static_assert(KC_TILE_ENTRIES == KC_TILE_BIAS, "the bias of kc_tile_word is the size of a tile");

struct TileSmem {
	unsigned long long *stage; /* KC_TILE_ENTRIES: the tile, sorted by region                          */
	unsigned long long *gpos;  /* per region: where its entries go (kc_tile_word, kcgpu_kernels.cuh)               */
	uint32_t *cnt;             /* per region: entries in the tile, then (lbase) where its run starts   */
};

__device__ __forceinline__ TileSmem kc_tile_smem(unsigned long long *smem, uint32_t n_regions)
{
	TileSmem s;
	s.stage = smem;
	s.gpos = smem + KC_TILE_ENTRIES;
	s.cnt = reinterpret_cast<uint32_t *>(s.gpos + n_regions);
	return s;
}

struct TileDest {
	uint64_t *table;             /* the owner's table */
	uint32_t *bloom;             /* its Bloom filter */
	uint64_t *lists;             /* region lists: `stride` entries apart, room for `cap` entries each */
	unsigned long long *cursors;
	uint64_t cap, stride;
	uint32_t region_bits, rslot_bits;
};

/* s.cnt must be zero (and that visible to the CTA) on entry; all threads of the CTA call. */
__device__ __forceinline__ void kc_file_tile(const uint64_t (&q)[KC_TILE_N], const uint32_t ok, const TileSmem &s, const TileDest &d,
                                             const InsertCtl &ctl, uint32_t &n_direct, uint32_t &n_new, uint32_t &n_overflow)
{
	__shared__ uint32_t s_warp[KC_TILE_THREADS / 32];
	__shared__ uint32_t s_total;
	const uint32_t n_regions = 1u << d.region_bits, rmask = n_regions - 1u;
	const uint32_t tid = threadIdx.x, lane = tid & 31u;
	uint32_t rk[KC_TILE_N / 2];
#pragma unroll
	for (int j = 0; j < KC_TILE_N; ++j) {
		uint32_t r = 0;
		if (ok >> j & 1u) r = atomicAdd(s.cnt + ((uint32_t)q[j] & rmask), 1u);
		rk[j >> 1] = (j & 1) ? rk[j >> 1] | r << 16 : r;
	}
	__syncthreads();
	/* exclusive prefix over the regions' counts; every region with entries reserves its run */
	{
		const uint32_t per = (n_regions + KC_TILE_THREADS - 1) / KC_TILE_THREADS, lo = tid * per;
		uint32_t sum = 0;
		for (uint32_t i = 0; i < per; ++i)
			if (lo + i < n_regions) sum += s.cnt[lo + i];
		uint32_t incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t v = __shfl_up_sync(KC_FULL, incl, o);
			if (lane >= (uint32_t)o) incl += v;
		}
		if (lane == 31u) s_warp[tid >> 5] = incl;
		__syncthreads();
		if (tid < 32u) {
			const uint32_t w = tid < KC_TILE_THREADS / 32 ? s_warp[tid] : 0u;
			uint32_t wi = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t v = __shfl_up_sync(KC_FULL, wi, o);
				if (lane >= (uint32_t)o) wi += v;
			}
			if (tid < KC_TILE_THREADS / 32) s_warp[tid] = wi - w;
		}
		__syncthreads();
		uint32_t at = s_warp[tid >> 5] + incl - sum;
		for (uint32_t i = 0; i < per; ++i) {
			const uint32_t r = lo + i;
			if (r >= n_regions) break;
			const uint32_t c = s.cnt[r];
			s.cnt[r] = at;
			if (c) {
				const unsigned long long g = atomicAdd(d.cursors + (uint64_t)r * KC_CURSOR_STRIDE, (unsigned long long)c);
				s.gpos[r] = kc_tile_word(g, c, at, r, d.cap, d.stride);
			}
			at += c;
		}
		if (tid == KC_TILE_THREADS - 1) s_total = at;
	}
	__syncthreads();
#pragma unroll
	for (int j = 0; j < KC_TILE_N; ++j)
		if (ok >> j & 1u) s.stage[s.cnt[(uint32_t)q[j] & rmask] + (rk[j >> 1] >> (16 * (j & 1)) & 0xFFFFu)] = q[j];
	__syncthreads();
	const uint32_t total = s_total;
	for (uint32_t i = tid; i < total; i += KC_TILE_THREADS) {
		const uint64_t w = s.stage[i];
		const uint32_t r = (uint32_t)w & rmask;
		uint64_t pos;
		if (kc_tile_fits(s.gpos[r], i, &pos)) {
			__stcs(reinterpret_cast<unsigned long long *>(d.lists) + pos, (unsigned long long)w); /* read once, by the flush: no reason to stay in L2 */
		} else { /* the run crosses the end of the list: what fits is filed, the rest goes straight to the table */
			if (pos < d.cap) {
				d.lists[(uint64_t)r * d.stride + pos] = w;
			} else {
				++n_direct;
				kc_insert_slow(d.table, d.region_bits, d.rslot_bits, w, d.bloom, ctl.mode, ctl.bloom_bits, ctl.bloom_hashes, &n_new, &n_overflow);
			}
		}
	}
}

/* one owner: extract, hash and file, a tile of KC_TILE_THREADS chunks at a time */
__global__ void __launch_bounds__(KC_TILE_THREADS, KC_TILE_MIN_CTAS) kc_scan_tile_kernel(const CountArgs a)
{
	extern __shared__ unsigned long long kc_dyn_smem[];
	const TileSmem s = kc_tile_smem(kc_dyn_smem, 1u << a.region_bits);
	TileDest d;
	d.table = a.tables[0];
	d.bloom = kc_bloom_of(d.table, a.n_slots, a.list_cap, a.region_bits);
	d.lists = kc_lists_of(d.table, a.n_slots);
	d.cursors = kc_cursors_of(d.table, a.n_slots, a.list_cap, a.region_bits);
	d.cap = a.list_cap;
	d.stride = a.list_cap;
	d.region_bits = a.region_bits;
	d.rslot_bits = a.rslot_bits;
	const uint4 *chunks = reinterpret_cast<const uint4 *>(a.bytes);
	const Extract x = kc_extract_of(a.k);
	uint32_t n_kmers = 0, n_new = 0, n_overflow = 0, n_direct = 0;
	for (uint64_t c0 = a.first_chunk + (uint64_t)blockIdx.x * KC_TILE_THREADS; c0 < a.end_chunk; c0 += (uint64_t)gridDim.x * KC_TILE_THREADS) {
		/* the last tile's entries are on their way out of shared memory; its counts are not needed any more */
		for (uint32_t r = threadIdx.x; r < (1u << a.region_bits); r += KC_TILE_THREADS) s.cnt[r] = 0;
		__syncthreads();
		uint64_t q[KC_TILE_N];
		const uint32_t ok = kc_extract16(chunks, c0 + threadIdx.x, a.end_chunk, x, q);
		n_kmers += __popc(ok);
		kc_file_tile(q, ok, s, d, a.ctl, n_direct, n_new, n_overflow);
	}
	for (int o = 16; o; o >>= 1) {
		n_kmers += __shfl_xor_sync(KC_FULL, n_kmers, o);
		n_new += __shfl_xor_sync(KC_FULL, n_new, o);
		n_overflow += __shfl_xor_sync(KC_FULL, n_overflow, o);
		n_direct += __shfl_xor_sync(KC_FULL, n_direct, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (n_kmers) atomicAdd(a.stats + KC_ST_KMERS, (unsigned long long)n_kmers);
		if (n_new) atomicAdd(a.stats + KC_ST_NEW, (unsigned long long)n_new);
		if (n_overflow) atomicAdd(a.stats + KC_ST_OVERFLOW, (unsigned long long)n_overflow);
		if (n_direct) atomicAdd(a.stats + KC_ST_DIRECT, (unsigned long long)n_direct);
	}
}

/* several owners: extract, hash, sort the tile by owner in shared memory, and push every owner's share
 * into its inbox as ONE run (one atomic on the owner's cursor per tile, 8 KB or so of neighbouring
 * stores) -- into peer memory over NVLink for the other GPUs: the all-to-all of the partition step
 * inside the producer */
__global__ void __launch_bounds__(KC_TILE_THREADS, KC_TILE_MIN_CTAS) kc_push_tile_kernel(const CountArgs a, const int part_shift)
{
	extern __shared__ unsigned long long kc_dyn_smem[];
	unsigned long long *stage = kc_dyn_smem; /* KC_TILE_ENTRIES */
	__shared__ uint32_t s_cnt[KC_MAX_PARTS], s_lbase[KC_MAX_PARTS + 1];
	__shared__ unsigned long long s_gbase[KC_MAX_PARTS];
	const uint4 *chunks = reinterpret_cast<const uint4 *>(a.bytes);
	const Extract x = kc_extract_of(a.k);
	const uint32_t tid = threadIdx.x, lane = tid & 31u;
	const uint64_t cap = kc_inbox_cap(a.list_cap, a.region_bits);
	uint32_t n_kmers = 0, n_new = 0, n_overflow = 0, n_direct = 0;
	for (uint64_t c0 = a.first_chunk + (uint64_t)blockIdx.x * KC_TILE_THREADS; c0 < a.end_chunk; c0 += (uint64_t)gridDim.x * KC_TILE_THREADS) {
		if (tid < KC_MAX_PARTS) s_cnt[tid] = 0;
		__syncthreads(); /* also: the last tile has left shared memory */
		uint64_t q[KC_TILE_N];
		const uint32_t ok = kc_extract16(chunks, c0 + tid, a.end_chunk, x, q);
		n_kmers += __popc(ok);
		/* rank within the owner's share: the lanes of a warp that have a k-mer for the same owner take
		 * neighbouring ranks with one shared-memory atomic between them */
		uint64_t owners = 0; /* 4 bits per position */
		uint32_t rk[KC_TILE_N / 2];
#pragma unroll
		for (int j = 0; j < KC_TILE_N; ++j) {
			uint32_t owner;
			kc_owner(q[j], a.n_parts, part_shift, owner, q[j]);
			const bool mine = ok >> j & 1u;
			const uint32_t peers = __match_any_sync(KC_FULL, mine ? owner : KC_MAX_PARTS);
			uint32_t r = 0;
			if (mine) {
				const int leader = __ffs(peers) - 1;
				if ((int)lane == leader) r = atomicAdd(&s_cnt[owner], (uint32_t)__popc(peers));
				r = __shfl_sync(peers, r, leader) + __popc(peers & ((1u << lane) - 1u));
			}
			owners |= (uint64_t)owner << (4 * j);
			rk[j >> 1] = (j & 1) ? rk[j >> 1] | r << 16 : r;
		}
		__syncthreads();
		if (tid < a.n_parts) {
			uint32_t at = 0;
			for (uint32_t o = 0; o < tid; ++o) at += s_cnt[o];
			s_lbase[tid] = at;
			const uint32_t n = s_cnt[tid];
			if (tid == a.n_parts - 1) s_lbase[a.n_parts] = at + n;
			if (n) s_gbase[tid] = atomicAdd(kc_inbox_cursor(a.tables[tid], a.n_slots, a.list_cap, a.region_bits), (unsigned long long)n);
		}
		__syncthreads();
#pragma unroll
		for (int j = 0; j < KC_TILE_N; ++j)
			if (ok >> j & 1u) stage[s_lbase[(uint32_t)(owners >> (4 * j)) & 15u] + (rk[j >> 1] >> (16 * (j & 1)) & 0xFFFFu)] = q[j];
		__syncthreads();
		for (uint32_t o = 0; o < a.n_parts; ++o) {
			const uint32_t lo = s_lbase[o], n = s_lbase[o + 1] - lo;
			if (!n) continue;
			uint64_t *base = a.tables[o];
			uint64_t *inbox = kc_lists_of(base, a.n_slots);
			const unsigned long long g = s_gbase[o];
			for (uint32_t t = tid; t < n; t += KC_TILE_THREADS) {
				const uint64_t w = stage[lo + t];
				if (g + t < cap) {
					inbox[g + t] = w;
				} else { /* inbox full: straight to the owner's table */
					++n_direct;
					kc_insert_slow(base, a.region_bits, a.rslot_bits, w, kc_bloom_of(base, a.n_slots, a.list_cap, a.region_bits), a.ctl.mode,
					               a.ctl.bloom_bits, a.ctl.bloom_hashes, &n_new, &n_overflow);
				}
			}
		}
	}
	for (int o = 16; o; o >>= 1) {
		n_kmers += __shfl_xor_sync(KC_FULL, n_kmers, o);
		n_new += __shfl_xor_sync(KC_FULL, n_new, o);
		n_overflow += __shfl_xor_sync(KC_FULL, n_overflow, o);
		n_direct += __shfl_xor_sync(KC_FULL, n_direct, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (n_kmers) atomicAdd(a.stats + KC_ST_KMERS, (unsigned long long)n_kmers);
		if (n_new) atomicAdd(a.stats + KC_ST_NEW, (unsigned long long)n_new);
		if (n_overflow) atomicAdd(a.stats + KC_ST_OVERFLOW, (unsigned long long)n_overflow);
		if (n_direct) atomicAdd(a.stats + KC_ST_DIRECT, (unsigned long long)n_direct);
	}
}

/* several owners: the inbox into the region lists, tile by tile (c4x_insert_buf, kc-c4.c:64-72, on the owner's side) */
__global__ void __launch_bounds__(KC_TILE_THREADS, KC_TILE_MIN_CTAS) kc_route_tile_kernel(const RouteArgs a)
{
	extern __shared__ unsigned long long kc_dyn_smem[];
	const uint32_t n_regions = 1u << a.region_bits;
	const TileSmem s = kc_tile_smem(kc_dyn_smem, n_regions);
	const uint64_t filled = *a.n_ptr;
	const uint64_t n = filled < a.inbox_cap ? filled : a.inbox_cap;
	uint32_t n_new = 0, n_overflow = 0, n_direct = 0;
	TileDest d;
	d.table = a.table;
	d.bloom = a.bloom;
	d.lists = a.lists;
	d.cursors = a.cursors;
	d.cap = a.cap;
	d.stride = a.cap;
	d.region_bits = a.region_bits;
	d.rslot_bits = a.rslot_bits;
	for (uint64_t t0 = (uint64_t)blockIdx.x * KC_TILE_ENTRIES; t0 < n; t0 += (uint64_t)gridDim.x * KC_TILE_ENTRIES) {
		for (uint32_t r = threadIdx.x; r < n_regions; r += KC_TILE_THREADS) s.cnt[r] = 0;
		__syncthreads();
		uint64_t q[KC_TILE_N];
		uint32_t ok = 0;
#pragma unroll
		for (int j = 0; j < KC_TILE_N; ++j) {
			const uint64_t i = t0 + (uint64_t)j * KC_TILE_THREADS + threadIdx.x;
			q[j] = 0;
			if (i < n) q[j] = __ldcs(reinterpret_cast<const unsigned long long *>(a.inbox + i)), ok |= 1u << j;
		}
		kc_file_tile(q, ok, s, d, a.ctl, n_direct, n_new, n_overflow);
		__syncthreads();
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

/* The region lists into the table.  CTA b takes tile b % tiles_per_region of region
 * b / tiles_per_region: CTAs are dispatched in order, so the resident ones work on two or
 * three neighbouring regions (one or two when they are as large as they get, 64 MiB) and their slices stay in L2 while they fill. */
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
	/* one entry per thread and step: more entries in flight per thread (four home slots requested
	 * at once in round 1, two in round 2 -- 39 registers, or 32 forced) measured slower every time
	 * (config 5: 427 / 377 ms against 353): the compare-and-swap chain, not the first load, is
	 * what a thread waits for */
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
/* shared memory of the tile kernels: the tile, and 12 bytes per region */
static size_t tile_smem_bytes(uint32_t region_bits) { return (size_t)KC_TILE_ENTRIES * 8 + ((size_t)12 << region_bits); }

cudaError_t launch_partition(const CountArgs &a, cudaStream_t stream)
{
	if (a.region_bits > KC_TILE_REGION_BITS) return cudaErrorInvalidValue; /* kcgpu_create never makes that many regions */
	if (a.end_chunk <= a.first_chunk) return cudaSuccess;
	/* one tile per CTA as long as the grid allows; the kernel walks on from there */
	uint64_t blocks = (a.end_chunk - a.first_chunk + KC_TILE_THREADS - 1) / KC_TILE_THREADS;
	if (blocks > 0x7FFFFFFFull) blocks = 0x7FFFFFFFull;
	const size_t smem = tile_smem_bytes(a.region_bits);
	cudaError_t e = cudaFuncSetAttribute(kc_scan_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	kc_scan_tile_kernel<<<(unsigned)blocks, KC_TILE_THREADS, smem, stream>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_extract(const CountArgs &a, cudaStream_t stream) { return launch_scan<KC_EXTRACT>(a, stream); }
cudaError_t launch_push(const CountArgs &a, cudaStream_t stream)
{
	if (a.end_chunk <= a.first_chunk) return cudaSuccess;
	uint64_t blocks = (a.end_chunk - a.first_chunk + KC_TILE_THREADS - 1) / KC_TILE_THREADS;
	if (blocks > 0x7FFFFFFFull) blocks = 0x7FFFFFFFull;
	const size_t smem = (size_t)KC_TILE_ENTRIES * 8;
	cudaError_t e = cudaFuncSetAttribute(kc_push_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	kc_push_tile_kernel<<<(unsigned)blocks, KC_TILE_THREADS, smem, stream>>>(a, shift_of(a.n_parts));
	return cudaGetLastError();
}

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
	if (a.region_bits > KC_TILE_REGION_BITS) return cudaErrorInvalidValue;
	const size_t smem = tile_smem_bytes(a.region_bits);
	cudaError_t e = cudaFuncSetAttribute(kc_route_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	kc_route_tile_kernel<<<(unsigned)(n_sm > 0 ? n_sm : 1) * KC_TILE_MIN_CTAS, KC_TILE_THREADS, smem, stream>>>(a);
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
