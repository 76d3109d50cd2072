/*
 * vafgpu_kernels.cu -- sm_100a kernels of the vaf-counter hot path.
 *
 *  anchor_scan_kernel<S>   the product path: one fused pass that reads the ASCII stream
 *                          with 128-bit loads, packs 16 bases into one 32-bit word per
 *                          thread, forms one anchor every S bases, tests it against a
 *                          Bloom filter held in shared memory and, for the few survivors,
 *                          walks the L2-resident exact table, re-reads the k raw bytes and
 *                          bumps the ref/alt counter with a warp-aggregated atomic.
 *                          Replaces extract_kmers_to_buf + worker_lookup
 *                          (vaf-counter.c:349-427, 449-479) and the SSSE3 encoder
 *                          (vaf-counter.c:261-291).
 *  recipe_scan_kernel      the literal recipe (rolling forward / reverse-complement words,
 *                          canonical minimum, khashl hash and probe) kept as the on-device
 *                          verification mode.
 *
 * HBM-bound integer work: no tensor cores, no TMEM; what matters is coalesced 16-byte
 * loads, few issue slots per base and keeping the random accesses on chip.
 */
#include "vafgpu_kernels.cuh"

namespace vafgpu {

#define FULL 0xFFFFFFFFu

/* ------------------------------------------------------------------------------------ */
/* shared device helpers                                                                  */

/* warp-aggregated counter bump: lanes that hit the same counter elect one leader that adds
 * the group size (hits are rare, but a deep-coverage SNP makes many lanes hit one word) */
__device__ __forceinline__ void bump(uint32_t *counts, uint32_t val)
{
	unsigned active = __activemask();
	unsigned peers = __match_any_sync(active, val);
	if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&counts[val], (uint32_t)__popc(peers));
}

__device__ __forceinline__ bool is_base(uint32_t b)
{
	uint32_t u = b & 0xDFu; /* fold case */
	return u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'U';
}

/* 16 ASCII bases -> 16 two-bit codes, first base in the low bits.  (b >> 1) & 3 maps
 * A,C,T/U,G (either case) to 0,1,2,3; one AND and one multiply gather the four codes of a
 * 32-bit word into its top byte, three byte permutes gather the four top bytes. */
__device__ __forceinline__ uint32_t pack16(uint4 w)
{
	const uint32_t M = 0x00820820u; /* 2^23 + 2^17 + 2^11 + 2^5 */
	uint32_t p0 = (w.x & 0x06060606u) * M;
	uint32_t p1 = (w.y & 0x06060606u) * M;
	uint32_t p2 = (w.z & 0x06060606u) * M;
	uint32_t p3 = (w.w & 0x06060606u) * M;
	uint32_t lo = __byte_perm(p0, p1, 0x0073);
	uint32_t hi = __byte_perm(p2, p3, 0x0073);
	return __byte_perm(lo, hi, 0x5410);
}

/* ------------------------------------------------------------------------------------ */
/* anchor-filter kernel                                                                   */

struct AnchorParams {
	const uint4 *chunks;   /* stream as 16-byte chunks */
	const uint8_t *bytes;
	uint64_t n_bytes;
	uint32_t n_chunks;
	uint32_t n_tiles;      /* 32 chunks each */
	uint32_t tiles_per_span;
	uint32_t n_spans;
	uint32_t *counts;
	unsigned long long *stats;
	const uint32_t *filter;
	uint32_t filter_words;
	const vg_slot_t *slots;
	uint32_t slot_bits;
	int k, len;
};

__device__ __forceinline__ uint4 load_chunk(const AnchorParams &p, uint32_t chunk, bool want)
{
	if (want && chunk < p.n_chunks) return __ldcs(p.chunks + chunk); /* streaming: evict first */
	return make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
}

/* survivors of the filter: walk the exact table from the anchor's home slot; every slot
 * filed under this anchor names an oriented pattern k-mer and where the anchor sits in it */
__device__ __noinline__ void resolve_candidate(const AnchorParams &p, uint32_t anchor, uint32_t amask,
                                               uint64_t q, uint32_t &n_hits)
{
	const uint32_t smask = (1u << p.slot_bits) - 1u;
	uint32_t s = vg_slot_home(anchor, p.slot_bits);
	for (;;) {
		uint4 raw = __ldg(reinterpret_cast<const uint4 *>(p.slots) + s);
		uint64_t okey = (uint64_t)raw.y << 32 | raw.x;
		if (okey == VG_EMPTY_KEY) return;
		uint32_t val = raw.z, off = raw.w;
		if (((uint32_t)(okey >> 2 * off) & amask) == anchor && q >= off && q - off + p.k <= p.n_bytes) {
			/* compare the k raw bytes at q - off with the oriented key */
			const uint8_t *b = p.bytes + (q - off);
			uint64_t km = 0;
			bool ok = true;
			for (int i = 0; i < p.k; ++i) {
				uint32_t c = b[i];
				ok &= is_base(c);
				km |= (uint64_t)((c >> 1) & 3u) << 2 * i;
			}
			if (ok && km == okey) {
				bump(p.counts, val);
				++n_hits;
			}
		}
		s = (s + 1) & smask;
	}
}

template <int S>
__global__ void __launch_bounds__(1024, 1) anchor_scan_kernel(const __grid_constant__ AnchorParams p)
{
	extern __shared__ uint32_t s_filter[];
	{ /* stage the filter */
		const uint4 *src = reinterpret_cast<const uint4 *>(p.filter);
		uint4 *dst = reinterpret_cast<uint4 *>(s_filter);
		for (uint32_t i = threadIdx.x; i < p.filter_words / 4; i += blockDim.x) dst[i] = __ldg(src + i);
	}
	__syncthreads();

	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps_per_cta = blockDim.x >> 5;
	const uint32_t warp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
	const uint32_t n_warps = gridDim.x * warps_per_cta;
	const uint32_t amask = vg_mask32(p.len);
	const uint32_t nw = p.filter_words;
	uint32_t n_cand = 0, n_hits = 0;

	for (uint32_t span = warp; span < p.n_spans; span += n_warps) {
		const uint32_t t0 = span * p.tiles_per_span;
		const uint32_t t1 = min(t0 + p.tiles_per_span, p.n_tiles);
		/* software pipeline: tile t is processed while t+1 is decoded and t+2 is in flight.
		 * The tile after the span is only needed for its first chunk (lane 0). */
		uint4 w_next = load_chunk(p, (t0 + 1) * 32 + lane, t0 + 1 < t1 || lane == 0);
		uint32_t cur = pack16(load_chunk(p, t0 * 32 + lane, true));
		for (uint32_t t = t0; t < t1; ++t) {
			uint4 w_after = load_chunk(p, (t + 2) * 32 + lane, t + 2 < t1 || (t + 2 == t1 && lane == 0));
			uint32_t nxt_tile = pack16(w_next);
			uint32_t nxt = 0;
			if (S < 16) { /* the anchor at offset 16 - S may run into the next chunk */
				nxt = __shfl_down_sync(FULL, cur, 1);
				uint32_t head = __shfl_sync(FULL, nxt_tile, 0);
				if (lane == 31) nxt = head;
			}
			const uint64_t q0 = (uint64_t)(t * 32 + lane) * 16;
#pragma unroll
			for (int j = 0; j < 16 / S; ++j) {
				uint32_t a = (j == 0 ? cur : __funnelshift_r(cur, nxt, 2 * j * S)) & amask;
				uint32_t h = vg_filter_hash(vg_canon32(a, p.len));
				uint32_t word = s_filter[vg_filter_word(h, nw)];
				uint32_t m = vg_filter_mask(h);
				if ((word & m) == m) {
					++n_cand;
					resolve_candidate(p, a, amask, q0 + j * S, n_hits);
				}
			}
			cur = nxt_tile;
			w_next = w_after;
		}
	}
	/* statistics: one atomic per warp */
	for (int o = 16; o; o >>= 1) {
		n_cand += __shfl_xor_sync(FULL, n_cand, o);
		n_hits += __shfl_xor_sync(FULL, n_hits, o);
	}
	if (lane == 0 && (n_cand | n_hits)) {
		atomicAdd(&p.stats[ST_CANDIDATES], (unsigned long long)n_cand);
		atomicAdd(&p.stats[ST_HITS], (unsigned long long)n_hits);
	}
}

/* ------------------------------------------------------------------------------------ */
/* recipe kernel: vaf-counter.c:349-427 + 449-479 as written, one segment per thread      */

#define RECIPE_SEG 32

__global__ void __launch_bounds__(256) recipe_scan_kernel(const ScanArgs a)
{
	const uint64_t seg = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t first = seg * RECIPE_SEG; /* first position this thread may emit at */
	const int k = a.k;
	const uint64_t mask = (1ULL << 2 * k) - 1, one = 1;
	const int shift = 2 * (k - 1);
	const uint32_t bmask = (1u << a.rbits) - 1u;
	uint32_t n_kmers = 0, n_hits = 0;
	if (first < a.n_bytes) {
		/* warm up on the k-1 bytes before the segment: the run length and both words after
		 * them equal those of a scan from the read start whenever a k-mer can be emitted */
		uint64_t i = first >= (uint64_t)(k - 1) ? first - (k - 1) : 0;
		const uint64_t end = min(first + RECIPE_SEG, a.n_bytes);
		uint64_t fw = 0, rc = 0;
		int run = 0;
		for (; i < end; ++i) {
			uint32_t b = a.bytes[i];
			if (!is_base(b)) {
				run = 0;
				fw = rc = 0;
				continue;
			}
			uint64_t c = (b >> 1) & 3u;
			c ^= c >> 1; /* A0 C1 T2 G3 -> the reference's A0 C1 G2 T3 */
			fw = (fw << 2 | c) & mask;
			rc = rc >> 2 | (3 - c) << shift;
			if (++run < k || i < first) continue;
			uint64_t y = fw < rc ? fw : rc;
			++n_kmers;
			uint32_t s = vg_h2b(vg_kmer_hash(y), a.rbits);
			for (;;) { /* khashl.h:137-150 */
				uint64_t key = __ldg(a.rkeys + s);
				if (key == VG_EMPTY_KEY) break;
				if (key == y) {
					bump(a.counts, __ldg(a.rvals + s));
					++n_hits;
					break;
				}
				s = (s + 1) & bmask;
			}
		}
		(void)one;
	}
	for (int o = 16; o; o >>= 1) {
		n_kmers += __shfl_xor_sync(FULL, n_kmers, o);
		n_hits += __shfl_xor_sync(FULL, n_hits, o);
	}
	if ((threadIdx.x & 31) == 0 && (n_kmers | n_hits)) {
		atomicAdd(&a.stats[ST_KMERS], (unsigned long long)n_kmers);
		atomicAdd(&a.stats[ST_HITS], (unsigned long long)n_hits);
	}
}

/* ------------------------------------------------------------------------------------ */
/* launchers                                                                              */

static const int kMaxDynSmem = VG_MAX_FILTER_WORDS * 4;

cudaError_t kernels_init_device(int)
{
	cudaError_t e;
#define OPT_IN(S)                                                                                   \
	e = cudaFuncSetAttribute(anchor_scan_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
	                         kMaxDynSmem);                                                         \
	if (e != cudaSuccess) return e;
	OPT_IN(1) OPT_IN(2) OPT_IN(4) OPT_IN(8) OPT_IN(16)
#undef OPT_IN
	return cudaSuccess;
}

cudaError_t launch_anchor_scan(const ScanArgs &a, int n_sm, cudaStream_t stream)
{
	if (a.n_bytes == 0) return cudaSuccess;
	if (a.n_bytes / 16 > 0xFFFFFF00ull) return cudaErrorInvalidValue; /* chunk index is 32-bit */
	AnchorParams p;
	p.chunks = reinterpret_cast<const uint4 *>(a.bytes);
	p.bytes = a.bytes;
	p.n_bytes = a.n_bytes;
	p.n_chunks = (uint32_t)(a.n_bytes / 16);
	p.n_tiles = (p.n_chunks + 31) / 32;
	const int threads = 1024;
	const uint32_t resident_warps = (uint32_t)n_sm * (threads / 32);
	uint32_t tps = p.n_tiles / (resident_warps * 4u);
	p.tiles_per_span = tps < 1 ? 1 : (tps > 64 ? 64 : tps);
	p.n_spans = (p.n_tiles + p.tiles_per_span - 1) / p.tiles_per_span;
	p.counts = a.counts;
	p.stats = a.stats;
	p.filter = a.filter;
	p.filter_words = a.filter_words;
	p.slots = a.slots;
	p.slot_bits = a.slot_bits;
	p.k = a.k;
	p.len = a.len;
	uint32_t ctas = (p.n_spans + (threads / 32) - 1) / (threads / 32);
	if (ctas > (uint32_t)n_sm) ctas = (uint32_t)n_sm;
	const size_t smem = (size_t)a.filter_words * 4;
	switch (a.stride) {
	case 1: anchor_scan_kernel<1><<<ctas, threads, smem, stream>>>(p); break;
	case 2: anchor_scan_kernel<2><<<ctas, threads, smem, stream>>>(p); break;
	case 4: anchor_scan_kernel<4><<<ctas, threads, smem, stream>>>(p); break;
	case 8: anchor_scan_kernel<8><<<ctas, threads, smem, stream>>>(p); break;
	case 16: anchor_scan_kernel<16><<<ctas, threads, smem, stream>>>(p); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

cudaError_t launch_recipe_scan(const ScanArgs &a, int, cudaStream_t stream)
{
	if (a.n_bytes == 0) return cudaSuccess;
	const uint64_t segs = (a.n_bytes + RECIPE_SEG - 1) / RECIPE_SEG;
	const unsigned blocks = (unsigned)((segs + 255) / 256);
	recipe_scan_kernel<<<blocks, 256, 0, stream>>>(a);
	return cudaGetLastError();
}

} // namespace vafgpu
