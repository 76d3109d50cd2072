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

#include <cstdlib>

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

/* Per-warp candidate queue in shared memory: survivors of the filter are compacted into it
 * and resolved 32 at a time, one per lane, so that the L2 round trips of the exact table
 * are taken by full warps and off the streaming loop.  It must absorb everything one tile
 * can produce on top of the < 32 entries left by the last drain. */
template <int S> struct Launch {
	static constexpr int kThreads = VG_THREADS(S);
	static constexpr int kQueue = VG_QUEUE_ENTRIES(S); /* candidates + 32 verify entries */
	static constexpr int kQueueBytes = VG_QUEUE_BYTES(S);
};

struct AnchorParams {
	const uint4 *chunks;   /* the launch's range of the stream as 16-byte chunks        */
	const uint8_t *bytes;  /* the whole stream (candidate verification may look outside
	                          the range: a k-mer may start up to S-1 bases before it)    */
	uint64_t n_bytes;      /* length of the whole stream                                */
	uint64_t range_lo;     /* byte offset of chunks[0] in the stream                    */
	uint32_t n_chunks;     /* chunks in the range, < 2^28                               */
	uint32_t n_tiles;      /* 32 chunks each */
	uint32_t tiles_per_span;
	uint32_t n_spans;
	uint32_t *counts;
	unsigned long long *stats;
	const uint32_t *filter;
	uint32_t filter_words;
	const uint4 *tags;     /* buckets of four tags */
	const vg_slot_t *slots;
	uint32_t bucket_bits;
	int k, len;
	uint32_t pf_bytes;     /* L2 prefetch distance ahead of the register pipeline, 0 = off */
	uint32_t c4, c32;      /* the constants 4 and 32, passed as data so that index scaling and
	                          top-5-bit extraction compile to IMAD / IMAD.HI (FMA pipe) instead of
	                          LEA / SHF (integer ALU pipe) */
};

__device__ __forceinline__ void l2_prefetch(const void *ptr)
{
	asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

/* the stream load: read once, 16 bytes per lane.  VG_STREAM_LOAD picks the cache operator
 * (tuning knob kept while the kernel is being profiled). */
#ifndef VG_STREAM_LOAD
#define VG_STREAM_LOAD 0
#endif
__device__ __forceinline__ uint4 ld_stream(const uint4 *ptr)
{
#if VG_STREAM_LOAD == 0
	return __ldcs(ptr); /* ld.global.cs: evict-first in L1 and L2 */
#elif VG_STREAM_LOAD == 1
	return __ldg(ptr); /* ld.global.nc */
#elif VG_STREAM_LOAD == 2
	return __ldcg(ptr); /* ld.global.cg: L2 only */
#elif VG_STREAM_LOAD == 3
	uint4 v;
	asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
	return v;
#else
	return *ptr;
#endif
}

/* The exact table is hit at random while gigabytes stream past it: its lines are loaded
 * with an evict-last L2 policy so the stream (loaded evict-first) does not push them out. */
__device__ __forceinline__ uint64_t l2_keep_policy()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ uint4 ldg_keep(const uint4 *ptr, uint64_t pol)
{
	uint4 v;
	asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
	             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
	             : "l"(ptr), "l"(pol));
	return v;
}

/* Second stage: up to 32 (slot, anchor position) pairs whose tag matched, one per lane.  The
 * payload names an oriented pattern k-mer and how far before its end the anchor ends; the k
 * raw bytes of that place in the stream decide.  Batching them makes the two dependent L2
 * round trips (payload, then bytes) happen once per 32 verifications instead of once each,
 * and lets the counter update aggregate over the warp. */
template <int S>
__device__ __forceinline__ uint32_t verify_batch(const AnchorParams &p, const uint2 *vq, uint32_t n, uint32_t lane)
{
	bool ok = false;
	uint32_t val = 0;
	if (lane < n) {
		const uint2 e = vq[lane];
		const uint4 raw = ldg_keep(reinterpret_cast<const uint4 *>(p.slots) + e.x, l2_keep_policy());
		const uint64_t okey = (uint64_t)raw.y << 32 | raw.x;
		const uint64_t end = p.range_lo + (uint64_t)e.y * (uint32_t)S + raw.w; /* the k-mer would occupy [end - k, end) */
		val = raw.z;
		if (end >= (uint64_t)p.k && end <= p.n_bytes) {
			/* if all k bases agree, the anchor (whose tag may lack a bit) agrees as well */
			const uint8_t *b = p.bytes + (end - p.k);
			ok = true;
			for (int i = 0; i < p.k; ++i) {
				const uint32_t c = b[i];
				ok = ok && is_base(c) && ((c >> 1) & 3u) == ((uint32_t)(okey >> 2 * i) & 3u);
			}
		}
	}
	/* warp-aggregated counter update: lanes that found the same counter elect one leader */
	const uint32_t have = __ballot_sync(FULL, ok);
	if (ok) {
		const uint32_t peers = __match_any_sync(have, val);
		if (lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&p.counts[val], (uint32_t)__popc(peers));
	}
	return __popc(have);
}

/* First stage: resolve n <= 32 queued survivors of the filter, one per lane.  Fetch the
 * anchor's home bucket (four tags, one 16-byte load from L2).  Slots fill in scan order and
 * are never freed, so the occupied slots of a bucket are a prefix of it: a tag (never 0) can
 * only match an occupied slot, and the chain ends in this bucket iff its last slot is free.
 * A matching tag (rare: the anchor really is one a pattern carries) goes to the verify queue. */
template <int S>
__device__ __forceinline__ uint32_t drain_queue(const AnchorParams &p, const uint2 *wq, uint32_t first, uint32_t n,
                                                uint2 *vq, uint32_t &vn, uint32_t lane, uint32_t lt_mask)
{
	const uint32_t bmask = (1u << p.bucket_bits) - 1u;
	const uint64_t keep = l2_keep_policy();
	bool active = lane < n;
	uint2 e = make_uint2(0u, 0u);
	if (active) e = wq[first + lane];
	const uint32_t tag = vg_tag(e.x);
	uint32_t b = vg_bucket_home(e.x, p.bucket_bits), hits = 0;
	while (__any_sync(FULL, active)) {
		uint4 t = make_uint4(0u, 0u, 0u, 0u);
		if (active) t = ldg_keep(p.tags + b, keep);
		uint32_t mm = (t.x == tag ? 1u : 0u) | (t.y == tag ? 2u : 0u) | (t.z == tag ? 4u : 0u) | (t.w == tag ? 8u : 0u);
		if (!active) mm = 0;
		while (__any_sync(FULL, mm != 0)) { /* rare: queue one matching slot per lane and round */
			const bool m = mm != 0;
			const uint32_t votes = __ballot_sync(FULL, m);
			if (vn + __popc(votes) > 32) { /* make room: run a full batch first */
				__syncwarp();
				hits += verify_batch<S>(p, vq, vn, lane);
				vn = 0;
				__syncwarp();
			}
			if (m) {
				vq[vn + __popc(votes & lt_mask)] = make_uint2(b * 4 + (uint32_t)(__ffs(mm) - 1), e.y);
				mm &= mm - 1;
			}
			vn += __popc(votes);
		}
		if (t.w == 0) active = false;
		else b = (b + 1) & bmask;
	}
	return hits;
}

/* state a warp carries through the stream */
struct Pipe {
	uint32_t t;        /* next tile to scan                                        */
	uint32_t c;        /* this lane's chunk in it                                  */
	const uint4 *ptr;  /* its address                                              */
	uint32_t cur;      /* that chunk, packed                                       */
	uint32_t carry;    /* lane 31's packed chunk of the previous tile: lane 0's left neighbour */
	uint4 w0, w1, w2;  /* raw chunks in flight; buffer `phase` holds tile t+1      */
	uint32_t phase;
	uint32_t qn;       /* entries in the candidate queue                           */
};

/* shared-memory word `idx` of the table at shared address `base`; idx * four is an IMAD */
__device__ __forceinline__ uint32_t lds_word(uint32_t base, uint32_t idx, uint32_t four)
{
	uint32_t v;
	asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(idx * four + base));
	return v;
}

/* One tile.  `use` holds the raw chunk of tile t+1 (loaded three tiles ago); `fill` is the
 * buffer consumed by the previous tile and is refilled first thing with tile t+3.  The
 * anchors of a chunk end at its aligned offsets and reach back into the LEFT neighbour only
 * (lane-1's chunk, or lane 31's of the previous tile), so nothing in the probes waits for a
 * load; the only consumer of loaded data is the pack at the very end.
 *   LS  anchor length fixed at compile time (0 = take it from the parameters) */
template <int S, bool CANON, int LS, bool INTERIOR>
__device__ __forceinline__ void scan_tile(const AnchorParams &p, Pipe &s, uint4 &fill, const uint4 &use,
                                          uint32_t filter, uint32_t bits, uint2 *wq, uint32_t lane, uint32_t lt_mask)
{
	const uint32_t nw = p.filter_words, last = p.n_chunks - 1;
	const int L = LS ? LS : p.len;
	const uint32_t amask = vg_mask32(L);
	fill = INTERIOR ? ld_stream(s.ptr + 96) : ld_stream(p.chunks + min(s.c + 96, last));
	if (INTERIOR && p.pf_bytes) l2_prefetch(reinterpret_cast<const uint8_t *>(s.ptr) + p.pf_bytes); /* this lane's chunk, some tiles on */
	/* one rotate serves both needs: lanes 1..31 get their left neighbour, lane 0 gets lane 31's
	 * chunk, which is its left neighbour in the NEXT tile */
	const uint32_t rot = __shfl_sync(FULL, s.cur, (lane + 31) & 31);
	const uint32_t left = lane == 0 ? s.carry : rot;
#pragma unroll
	for (int j = 0; j < 16 / S; ++j) {
		/* bases [s0, s0 + L) relative to the chunk start, s0 = (j+1) S - L, possibly < 0 */
		const int sh = 2 * ((j + 1) * S - L);
		uint32_t a = sh >= 0 ? s.cur >> (sh & 31) : __funnelshift_r(left, s.cur, (sh + 32) & 31);
		if (L < 16) a &= amask;
		uint32_t key = a;
		if (CANON) { /* a * rc(a), see vg_rc32 */
			uint32_t r = __brev(a);
			r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
			key = a * ((r ^ 0xAAAAAAAAu) >> (32 - 2 * L));
		}
		const uint32_t word = lds_word(filter, vg_filter_word(key, nw), p.c4);
		const uint32_t h2 = vg_hash2(key);
		const uint32_t m1 = lds_word(bits, __umulhi(h2, p.c32), p.c4);              /* 1 << (h2 >> 27) */
		const uint32_t m2 = lds_word(bits, __umulhi(vg_hash3(h2), p.c32), p.c4);
		bool hit = (~word & (m1 | m2)) == 0;
		if (!INTERIOR) hit = hit && s.c <= last;
		/* queue the survivors, compacted.  Large panels: nearly every vote has a survivor, so
		 * the branch around the push is not worth its two instructions */
		if (CANON || __any_sync(FULL, hit)) {
			const uint32_t votes = __ballot_sync(FULL, hit);
			if (hit) wq[s.qn + __popc(votes & lt_mask)] = make_uint2(a, s.c * (16 / S) + j + 1);
			s.qn += __popc(votes);
		}
	}
	s.carry = rot; /* only lane 0's copy is ever used */
	s.cur = pack16(use);
	++s.t;
	s.c += 32;
	s.ptr += 32;
}

/* The hot loop: scan tiles until the span ends or 32 candidates are queued.  No calls, no
 * table walks.  Three raw buffers rotate by a phase counter instead of by register copies.
 * INTERIOR: every address touched (loads up to tile t1+2, prefetch some tiles further) is
 * inside the range, so nothing is clamped or predicated. */
template <int S, bool CANON, int LS, bool INTERIOR>
__device__ __forceinline__ void scan_tiles(const AnchorParams &p, Pipe &s, uint32_t t1, uint32_t filter, uint32_t bits,
                                           uint2 *wq, uint32_t lane, uint32_t lt_mask)
{
	for (;;) {
		if (s.phase == 0) {
			scan_tile<S, CANON, LS, INTERIOR>(p, s, s.w2, s.w0, filter, bits, wq, lane, lt_mask);
			s.phase = 1;
			if (s.t >= t1 || s.qn >= 32) break;
		}
		if (s.phase == 1) {
			scan_tile<S, CANON, LS, INTERIOR>(p, s, s.w0, s.w1, filter, bits, wq, lane, lt_mask);
			s.phase = 2;
			if (s.t >= t1 || s.qn >= 32) break;
		}
		scan_tile<S, CANON, LS, INTERIOR>(p, s, s.w1, s.w2, filter, bits, wq, lane, lt_mask);
		s.phase = 0;
		if (s.t >= t1 || s.qn >= 32) break;
	}
}

/* The streaming kernel.  One CTA per SM, persistent over spans of tiles (a tile = 32 chunks
 * of 16 bytes = one 128-bit load per lane).
 *   CANON  the filter holds strand-symmetric keys, which halves its load for large panels;
 *          small panels file both orientations and skip the reverse complement. */
template <int S, bool CANON, int LS>
__global__ void __launch_bounds__(Launch<S>::kThreads, 1) anchor_scan_kernel(const __grid_constant__ AnchorParams p)
{
	extern __shared__ uint32_t s_filter[]; /* filter words | 32-word bit table | candidate queues */
	const uint32_t nw = p.filter_words;
	{ /* stage the filter */
		const uint4 *src = reinterpret_cast<const uint4 *>(p.filter);
		uint4 *dst = reinterpret_cast<uint4 *>(s_filter);
		for (uint32_t i = threadIdx.x; i < nw / 4; i += blockDim.x) dst[i] = __ldg(src + i);
		if (threadIdx.x < 32) s_filter[nw + threadIdx.x] = 1u << threadIdx.x;
	}
	__syncthreads();
	const uint32_t filter_sa = (uint32_t)__cvta_generic_to_shared(s_filter);
	const uint32_t bits_sa = filter_sa + nw * 4;

	const uint32_t lane = threadIdx.x & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint2 *const wq = reinterpret_cast<uint2 *>(s_filter + nw + 32) + (threadIdx.x >> 5) * Launch<S>::kQueue;
	uint2 *const vq = wq + Launch<S>::kQueue - 32; /* the last 32 entries: tag matches awaiting verification */
	uint32_t vn = 0;
	const uint32_t warps_per_cta = blockDim.x >> 5;
	const uint32_t warp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
	const uint32_t n_warps = gridDim.x * warps_per_cta;
	const uint32_t last = p.n_chunks - 1;
	uint32_t n_cand = 0, n_hits = 0;
	Pipe s;
	s.qn = 0;
	s.t = 0;
	s.phase = 0;
	uint32_t t1 = 0, span = warp;
	bool interior = false;

	for (;;) {
		if (s.t >= t1 && span < p.n_spans) { /* open the next span */
			s.t = span * p.tiles_per_span;
			t1 = min(s.t + p.tiles_per_span, p.n_tiles);
			span += n_warps;
			/* Interior span: everything the pipeline touches (loads up to tile t1+1, prefetch up
			 * to tile t1+7) lies inside the range.  Otherwise loads are clamped to the last
			 * chunk: a chunk past the end is never scanned, and as a right neighbour it can only
			 * create a false candidate, which the bounds check of the verification rejects. */
			interior = (uint64_t)(t1 + 34) * 32 <= p.n_chunks;
			s.c = s.t * 32 + lane;
			s.ptr = p.chunks + s.c;
			if (interior && p.pf_bytes) l2_prefetch(reinterpret_cast<const uint8_t *>(p.chunks + s.t * 32) + lane * (p.pf_bytes / 32));
			/* register pipeline: tile t is scanned while t+1 is being packed and t+2 is in
			 * flight; the L2 prefetch runs 8 tiles ahead of that */
			s.cur = pack16(ld_stream(p.chunks + min(s.c, last)));
			s.w0 = ld_stream(p.chunks + min(s.c + 32, last));
			s.w1 = ld_stream(p.chunks + min(s.c + 64, last));
			s.phase = 0;
			/* the chunk before the span: the last one of the previous span, or of the previous
			 * launch range; nothing ('\n's) at the very start of the stream */
			const uint32_t c0 = s.t * 32;
			uint4 before = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
			if (c0 > 0 || p.range_lo > 0) before = __ldg(p.chunks + c0 - 1); /* chunks[-1] exists when range_lo > 0 */
			s.carry = pack16(before);
		}
		const bool finished = s.t >= t1; /* no span left */
		if (!finished) {
			if (interior) scan_tiles<S, CANON, LS, true>(p, s, t1, filter_sa, bits_sa, wq, lane, lt_mask);
			else scan_tiles<S, CANON, LS, false>(p, s, t1, filter_sa, bits_sa, wq, lane, lt_mask);
		}
		if (s.qn >= 32 || (finished && s.qn)) { /* the one place candidates are resolved */
			const uint32_t n = min(s.qn, 32u);
			s.qn -= n;
			n_cand += n;
			__syncwarp();
			n_hits += drain_queue<S>(p, wq, s.qn, n, vq, vn, lane, lt_mask);
			__syncwarp();
		} else if (finished) break;
	}
	if (vn) n_hits += verify_batch<S>(p, vq, vn, lane);
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

template <int S, bool CANON, int LS>
static cudaError_t launch_one(const AnchorParams &p0, uint32_t filter_words, int n_sm, cudaStream_t stream)
{
	AnchorParams p = p0;
	const int threads = Launch<S>::kThreads;
	const size_t smem = (size_t)filter_words * 4 + 128 + Launch<S>::kQueueBytes;
	static bool opted_in[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 64 && !opted_in[dev]) {
		cudaError_t e = cudaFuncSetAttribute(anchor_scan_kernel<S, CANON, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                                     VG_SMEM_BUDGET);
		if (e != cudaSuccess) return e;
		opted_in[dev] = true;
	}
	const uint32_t resident_warps = (uint32_t)n_sm * (threads / 32);
	uint32_t tps = p.n_tiles / (resident_warps * 4u);
	p.tiles_per_span = tps < 1 ? 1 : (tps > 64 ? 64 : tps);
	p.n_spans = (p.n_tiles + p.tiles_per_span - 1) / p.tiles_per_span;
	uint32_t ctas = (p.n_spans + (threads / 32) - 1) / (threads / 32);
	if (ctas > (uint32_t)n_sm) ctas = (uint32_t)n_sm;
	anchor_scan_kernel<S, CANON, LS><<<ctas, threads, smem, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t kernels_init_device(int) { return cudaSuccess; }

cudaError_t launch_anchor_scan(const ScanArgs &a, int n_sm, cudaStream_t stream)
{
	/* queue entries carry a 32-bit anchor index within the launch: cut the stream into
	 * ranges of at most 2 GiB; verification still sees the whole stream */
	const uint64_t kRange = 1ull << 31;
	for (uint64_t lo = 0; lo < a.n_bytes; lo += kRange) {
		const uint64_t n = a.n_bytes - lo < kRange ? a.n_bytes - lo : kRange;
		AnchorParams p;
		p.chunks = reinterpret_cast<const uint4 *>(a.bytes + lo);
		p.bytes = a.bytes;
		p.n_bytes = a.n_bytes;
		p.range_lo = lo;
		p.n_chunks = (uint32_t)(n / 16);
		p.n_tiles = (p.n_chunks + 31) / 32;
		p.tiles_per_span = p.n_spans = 0;
		p.counts = a.counts;
		p.stats = a.stats;
		p.filter = a.filter;
		p.filter_words = a.filter_words;
		p.tags = reinterpret_cast<const uint4 *>(a.tags);
		p.slots = a.slots;
		p.bucket_bits = a.bucket_bits;
		p.k = a.k;
		p.len = a.len;
		{
			static const char *env = getenv("VAFGPU_PF_TILES"); /* tuning knob: tiles of L2 prefetch lead, <= 32 */
			int tiles = env ? atoi(env) : 8;
			p.pf_bytes = (uint32_t)(tiles < 0 ? 0 : tiles > 32 ? 32 : tiles) * 512u;
		}
		cudaError_t e;
		p.c4 = 4;
		p.c32 = 32;
		/* the headline plans (k = 15, 21, 31) get the anchor length as a compile-time constant */
#define GO(S, LS) e = a.canon ? launch_one<S, true, LS>(p, a.filter_words, n_sm, stream) : launch_one<S, false, LS>(p, a.filter_words, n_sm, stream)
		switch (a.stride) {
		case 1: GO(1, 0); break;
		case 2: GO(2, 0); break;
		case 4: if (a.len == 12) GO(4, 12); else GO(4, 0); break;
		case 8: if (a.len == 14) GO(8, 14); else GO(8, 0); break;
		case 16: if (a.len == 16) GO(16, 16); else GO(16, 0); break;
		default: return cudaErrorInvalidValue;
		}
#undef GO
		if (e != cudaSuccess) return e;
	}
	return cudaSuccess;
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
