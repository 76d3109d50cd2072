/*
 * kcgpu_kernels.cu -- sm_100a kernels of the full k-mer counting mode (the kc-c4 path).
 *
 *  kc_scan_kernel<false>   fused extract + insert: every thread owns the 16 stream positions
 *                          of one 128-bit chunk, warms its two rolling words up on the 32
 *                          bytes before them (k - 1 <= 30), and for each position where a
 *                          k-mer ends (kc-c4.c:80-87) hashes the canonical word (kc-c4.c:40-50)
 *                          and adds it to its owner's table with 64-bit compare-and-swap.
 *                          The owner's table may be peer memory: the same instruction then
 *                          travels over NVLink, which is the all-to-all of kc-c4's partition
 *                          step (kc-c4.c:64-72) fused into the producer.
 *  kc_scan_kernel<true>    extract only: hashed k-mers are appended to one list per owner
 *                          (warp-aggregated), for an exchange by NCCL all-to-all.
 *  kc_insert_kernel        the insert step for lists that came from an exchange (kc-c4.c:116-128)
 *  kc_hist_kernel          worker_hist (kc-c4.c:186-197) over the slots
 *
 * No tensor cores: shifts, multiplies and atomics.  The bound is random 32-byte sector
 * traffic of the table.
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

/* kc-c4.c:116-128 on one table: find or claim the slot of q, count up to 1023 */
__device__ __forceinline__ void kc_insert(uint64_t *table, uint32_t region_bits, uint32_t rslot_bits, uint64_t q,
                                          uint32_t &n_new, uint32_t &n_overflow)
{
	const uint64_t region = q & ((1ull << region_bits) - 1ull);
	const uint64_t tag = q >> region_bits;
	const uint64_t rmask = (1ull << rslot_bits) - 1ull;
	uint64_t *base = table + (region << rslot_bits);
	uint64_t pos = (tag * 0x9E3779B97F4A7C15ull) >> (64 - rslot_bits);
	for (int tries = 0; tries < KC_MAX_PROBES; ++tries) {
		uint64_t *p = base + pos;
		uint64_t v = ld_slot(p);
		if (v == 0) {
			v = cas_slot(p, 0, (tag << KC_COUNT_BITS) | 1ull);
			if (v == 0) {
				++n_new;
				return;
			}
		}
		while ((v >> KC_COUNT_BITS) == tag) {
			if ((v & KC_COUNT_MAX) == KC_COUNT_MAX) return; /* kc-c4.c:125 */
			const uint64_t old = cas_slot(p, v, v + 1);
			if (old == v) return;
			v = old;
		}
		pos = (pos + 1) & rmask;
	}
	++n_overflow;
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

template <bool EXTRACT>
__global__ void __launch_bounds__(KC_THREADS) kc_scan_kernel(const CountArgs a, const int part_shift)
{
	const uint64_t n_chunks = a.n_bytes >> 4;
	const uint64_t c = (uint64_t)blockIdx.x * KC_THREADS + threadIdx.x;
	uint32_t n_kmers = 0, n_new = 0, n_overflow = 0, n_dropped = 0;
	if (c < n_chunks) {
		const uint4 *chunks = reinterpret_cast<const uint4 *>(a.bytes);
		const uint4 sep = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
		const uint4 w0 = c >= 2 ? __ldg(chunks + c - 2) : sep;
		const uint4 w1 = c >= 1 ? __ldg(chunks + c - 1) : sep;
		uint4 own = __ldg(chunks + c);
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
#pragma unroll 1
		for (int i = 0; i < 16; ++i) {
			const uint32_t b = own.x & 0xFFu;
			own.x = __funnelshift_r(own.x, own.y, 8);
			own.y = __funnelshift_r(own.y, own.z, 8);
			own.z = __funnelshift_r(own.z, own.w, 8);
			own.w >>= 8;
			uint64_t code = (b >> 1) & 3u;
			code ^= code >> 1;
			fw = (fw << 2 | code) & mask;
			rv = rv >> 2 | (3ull - code) << top;
			run = kc_is_base(b) ? run + 1 : 0;
			if (run < k) continue;
			++n_kmers;
			const uint64_t h = kc_hash64(fw < rv ? fw : rv, mask);
			uint32_t owner;
			uint64_t q;
			kc_owner(h, a.n_parts, part_shift, owner, q);
			if (!EXTRACT) {
				kc_insert(a.tables[owner], a.region_bits, a.rslot_bits, q, n_new, n_overflow);
			} else {
				/* one atomic per owner and warp: the lanes that have a k-mer for the same owner
				 * reserve consecutive entries of its list */
				cg::coalesced_group active = cg::coalesced_threads();
				cg::coalesced_group same = cg::labeled_partition(active, owner);
				uint32_t at = 0;
				if (same.thread_rank() == 0) at = atomicAdd(a.part_counts + owner, same.size());
				at = same.shfl(at, 0) + same.thread_rank();
				if (at < a.cap_per_part) a.out_keys[(uint64_t)owner * a.cap_per_part + at] = h;
				else ++n_dropped;
			}
		}
	}
	for (int o = 16; o; o >>= 1) {
		n_kmers += __shfl_xor_sync(KC_FULL, n_kmers, o);
		n_new += __shfl_xor_sync(KC_FULL, n_new, o);
		n_overflow += __shfl_xor_sync(KC_FULL, n_overflow, o);
		n_dropped += __shfl_xor_sync(KC_FULL, n_dropped, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (n_kmers) atomicAdd(a.stats + KC_ST_KMERS, (unsigned long long)n_kmers);
		if (n_new) atomicAdd(a.stats + KC_ST_NEW, (unsigned long long)n_new);
		if (n_overflow) atomicAdd(a.stats + KC_ST_OVERFLOW, (unsigned long long)n_overflow);
		if (n_dropped) atomicAdd(a.stats + KC_ST_DROPPED, (unsigned long long)n_dropped);
	}
}

__global__ void __launch_bounds__(KC_THREADS) kc_insert_kernel(const InsertArgs a, const int part_shift)
{
	uint32_t n_new = 0, n_overflow = 0, n_kmers = 0;
	const uint64_t stride = (uint64_t)gridDim.x * KC_THREADS;
	for (uint64_t i = (uint64_t)blockIdx.x * KC_THREADS + threadIdx.x; i < a.n; i += stride) {
		const uint64_t h = __ldg(a.hashed + i);
		uint32_t owner;
		uint64_t q;
		kc_owner(h, a.n_parts, part_shift, owner, q);
		++n_kmers;
		kc_insert(a.table, a.region_bits, a.rslot_bits, q, n_new, n_overflow);
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

int shift_of(uint32_t n_parts)
{
	if (n_parts & (n_parts - 1)) return -1;
	int s = 0;
	while ((1u << s) < n_parts) ++s;
	return s;
}

} // namespace

cudaError_t launch_count(const CountArgs &a, cudaStream_t stream)
{
	if (a.n_bytes == 0) return cudaSuccess;
	const uint64_t chunks = a.n_bytes >> 4;
	const uint64_t blocks = (chunks + KC_THREADS - 1) / KC_THREADS;
	if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
	kc_scan_kernel<false><<<(unsigned)blocks, KC_THREADS, 0, stream>>>(a, shift_of(a.n_parts));
	return cudaGetLastError();
}

cudaError_t launch_extract(const CountArgs &a, cudaStream_t stream)
{
	if (a.n_bytes == 0) return cudaSuccess;
	const uint64_t chunks = a.n_bytes >> 4;
	const uint64_t blocks = (chunks + KC_THREADS - 1) / KC_THREADS;
	if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
	kc_scan_kernel<true><<<(unsigned)blocks, KC_THREADS, 0, stream>>>(a, shift_of(a.n_parts));
	return cudaGetLastError();
}

cudaError_t launch_insert(const InsertArgs &a, cudaStream_t stream)
{
	if (a.n == 0) return cudaSuccess;
	uint64_t blocks = (a.n + KC_THREADS - 1) / KC_THREADS;
	if (blocks > 148ull * 64) blocks = 148ull * 64;
	kc_insert_kernel<<<(unsigned)blocks, KC_THREADS, 0, stream>>>(a, shift_of(a.n_parts));
	return cudaGetLastError();
}

cudaError_t launch_histogram(const uint64_t *table, uint64_t n_slots, unsigned long long *hist256, int n_sm, cudaStream_t stream)
{
	uint64_t blocks = ((n_slots >> 1) + KC_THREADS - 1) / KC_THREADS;
	const uint64_t resident = (uint64_t)(n_sm > 0 ? n_sm : 148) * 8; /* 8 CTAs of 256 threads per SM */
	if (blocks > resident) blocks = resident;
	if (blocks < 1) blocks = 1;
	kc_hist_kernel<<<(unsigned)blocks, KC_THREADS, 0, stream>>>(table, n_slots, hist256);
	return cudaGetLastError();
}

} // namespace kcgpu
