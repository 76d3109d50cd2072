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

/* Per-warp candidate queue in shared memory: anchors that need the exact table walked are
 * compacted into it and resolved 32 at a time, one per lane, off the streaming loop.  It
 * must absorb everything one tile can produce on top of the < 32 entries left by the last
 * drain. */
template <int S, bool DEFER> struct Launch {
	static constexpr int kThreads = VG_THREADS(S, DEFER);
	static constexpr int kQueue = VG_QUEUE_ENTRIES(S); /* candidates + 32 verify entries */
	static constexpr int kQueueBytes = VG_QUEUE_BYTES(S, DEFER);
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
	uint32_t n_spans;      /* n_full_spans of tiles_per_span tiles, then shorter ones of tail_tiles */
	uint32_t n_full_spans, tail_tiles;
	uint32_t *counts;
	unsigned long long *stats;
	const uint32_t *filter;
	uint32_t filter_words;
	const uint32_t *filter2; /* second filter level, L2-resident (deferred form) */
	uint32_t filter2_words;
	const uint4 *buckets;  /* three tags + control word each */
	uint32_t n_buckets;
	const vg_slot_t *slots;
	int k, len;
	uint64_t keep;         /* L2 evict-last access policy (createpolicy), made once per device */
	uint32_t c4;           /* the constant 4, passed as data so that index scaling compiles to
	                          IMAD (FMA pipe) instead of LEA (integer ALU pipe) */
};

/* Distance of the L2 prefetch ahead of the register pipeline, in bytes, and the cache policy
 * of every global load, per form of the kernel -- measured, not derived (profiles/r2_anchor_knobs.txt):
 *   queue form (small panels)     8 tiles ahead; the stream is read once and evict-first in L1
 *                                 and L2 (ld.global.cs), the exact table evict-last in L2;
 *   deferred form (large panels)  2 tiles ahead of a raw chunk that is itself loaded two steps
 *                                 before it is packed (three register buffers); stream, second
 *                                 filter level and exact table all through L2 only with the
 *                                 default priority (ld.global.cg).  With evict-first stream loads
 *                                 and evict-last table loads the same kernel runs 9 % slower:
 *                                 half of every GPU's L2 traffic crosses the die-to-die link and
 *                                 the priorities cost more there than they save. */
#ifndef VG_PF_BYTES
#define VG_PF_BYTES 4096
#endif
#ifndef VG_PF_BYTES_DEFER
#define VG_PF_BYTES_DEFER 1024
#endif
/* how many steps before it is packed a raw chunk of the deferred form is loaded (1: two
 * register buffers, 2: three) */
#ifndef VG_AHEAD
#define VG_AHEAD 2
#endif
template <bool DEFER> struct Knobs {
	static constexpr int kPrefetch = DEFER ? VG_PF_BYTES_DEFER : VG_PF_BYTES;
};

__device__ __forceinline__ void l2_prefetch(const void *ptr)
{
	asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}
template <int BYTES> __device__ __forceinline__ void l2_prefetch_ahead(const void *ptr)
{
	if (BYTES > 0) asm volatile("prefetch.global.L2 [%0+%1];" ::"l"(ptr), "n"(BYTES));
}

/* the stream load: 16 bytes per lane, read once */
template <bool DEFER> __device__ __forceinline__ uint4 ld_stream(const uint4 *ptr) { return DEFER ? __ldcg(ptr) : __ldcs(ptr); }

/* The exact table is hit at random while gigabytes stream past it.  Queue form: its lines are
 * loaded with an evict-last L2 policy so the stream (loaded evict-first) does not push them out. */
__device__ __forceinline__ uint64_t l2_keep_policy()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
template <bool DEFER> __device__ __forceinline__ uint4 ldg_table(const uint4 *ptr, uint64_t pol)
{
	if (DEFER) return __ldcg(ptr);
	uint4 v;
	asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
	             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
	             : "l"(ptr), "l"(pol));
	return v;
}

/* reverse complement of a packed word of 16 bases (stream encoding, complement = code ^ 2):
 * bit reversal reverses the bases and swaps the two bits of each; swap them back and flip
 * the high one */
__device__ __forceinline__ uint32_t rc16(uint32_t x)
{
	const uint32_t r = __brev(x);
	return (((r >> 1) & 0x55555555u) | ((r << 1) & 0xAAAAAAAAu)) ^ 0xAAAAAAAAu;
}

/* Second stage: up to 32 (payload index, anchor position) pairs whose tag matched, one per
 * lane.  The payload names an oriented pattern k-mer and how far before its end the anchor
 * ends; the k raw bytes of that place in the stream decide.  Batching them makes the two
 * dependent L2 round trips (payload, then bytes) happen once per 32 verifications instead of
 * once each, and lets the counter update aggregate over the warp. */
template <int S, bool DEFER>
__device__ __forceinline__ uint32_t verify_batch(const AnchorParams &p, const uint2 *vq, uint32_t n, uint32_t lane)
{
	bool ok = false;
	uint32_t val = 0;
	if (lane < n) {
		const uint2 e = vq[lane];
		const uint4 raw = ldg_table<DEFER>(reinterpret_cast<const uint4 *>(p.slots) + e.x, p.keep);
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

/* filter key of an anchor as the kernel's generic (slow) paths compute it */
template <bool CANON> __device__ __forceinline__ uint32_t anchor_key(uint32_t a, int L)
{
	return CANON ? a * (rc16(a) >> (32 - 2 * L)) : a;
}

/* First stage: resolve n <= 32 queued anchors, one per lane.  Fetch the anchor's home bucket
 * (three tags + control word, one 16-byte load from L2); a matching tag (rare: the anchor
 * really is one a pattern carries) goes to the verify queue; the control word says whether
 * entries of this chain live further on. */
template <int S, bool CANON, bool DEFER>
__device__ __forceinline__ uint32_t drain_queue(const AnchorParams &p, const uint2 *wq, uint32_t first, uint32_t n,
                                                uint2 *vq, uint32_t &vn, uint32_t lane, uint32_t lt_mask)
{
	const uint64_t keep = p.keep;
	bool active = lane < n;
	uint2 e = make_uint2(0u, 0u);
	if (active) e = wq[first + lane];
	const uint32_t tag = vg_tag(e.x, p.len);
	uint32_t b = vg_bucket_home(vg_hash_lo(anchor_key<CANON>(e.x, p.len), p.filter_words), p.n_buckets), hits = 0;
	while (__any_sync(FULL, active)) {
		uint4 t = make_uint4(VG_FREE_TAG, VG_FREE_TAG, VG_FREE_TAG, 0u);
		if (active) t = ldg_table<DEFER>(p.buckets + b, keep);
		uint32_t mm = (t.x == tag ? 1u : 0u) | (t.y == tag ? 2u : 0u) | (t.z == tag ? 4u : 0u);
		if (!active) mm = 0;
		while (__any_sync(FULL, mm != 0)) { /* rare: queue one matching slot per lane and round */
			const bool m = mm != 0;
			const uint32_t votes = __ballot_sync(FULL, m);
			if (vn + __popc(votes) > 32) { /* make room: run a full batch first */
				__syncwarp();
				hits += verify_batch<S, DEFER>(p, vq, vn, lane);
				vn = 0;
				__syncwarp();
			}
			if (m) {
				vq[vn + __popc(votes & lt_mask)] = make_uint2((t.w & ~VG_CTRL_MORE) + (uint32_t)(__ffs(mm) - 1), e.y);
				mm &= mm - 1;
			}
			vn += __popc(votes);
		}
		if (!(t.w & VG_CTRL_MORE)) active = false;
		else b = b + 1 == p.n_buckets ? 0 : b + 1;
	}
	return hits;
}

/* deferred path: what a lane remembers of the survivors of the previous tile */
template <int NA> struct Pending {
	uint32_t w[NA];   /* their words of the second filter level, requested a tile ago     */
	uint32_t pm[NA];  /* the two bits to find there                                       */
	uint32_t a[NA];   /* the anchors                                                      */
};

/* state a warp carries through the stream */
struct Pipe {
	uint32_t t;        /* next tile to scan                                        */
	uint32_t c;        /* this lane's chunk in it                                  */
	uint32_t cur;      /* that chunk, packed                                       */
	uint32_t carry;    /* lane 31's packed chunk of the previous tile: lane 0's left neighbour */
	uint32_t rcur, rcarry; /* reverse complements of the two (strand-symmetric keys only)  */
	uint4 w0, w1, w2;  /* raw chunks in flight                                     */
	uint32_t phase;    /* queue path: buffer `phase` holds tile t+1                */
	uint32_t qn;       /* entries in the candidate queue                           */
};

/* shared-memory word `idx` of the table at shared address `base`; idx * four is an IMAD */
__device__ __forceinline__ uint32_t lds_word(uint32_t base, uint32_t idx, uint32_t four)
{
	uint32_t v;
	asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(idx * four + base));
	return v;
}

/* the two filter bits of an anchor whose hash tail is `lo` */
__device__ __forceinline__ uint32_t pair_mask(uint32_t pairs, uint32_t lo, uint32_t four)
{
#if VG_PAIR_ALU
	/* shf.l.wrap takes the shift amount modulo 32, so the second field needs no masking */
	return (1u << (lo >> 27)) | __funnelshift_l(0u, 1u, lo >> 22);
#else
	return lds_word(pairs, vg_pair_index(lo), four);
#endif
}

/* The anchors of one chunk and their fate in the filter.  They end at the chunk's aligned
 * offsets and reach back into the LEFT neighbour only (lane-1's chunk, or lane 31's of the
 * previous tile), so nothing here waits for a load.
 *   LS  anchor length fixed at compile time (0 = take it from the parameters) */
template <int S, bool CANON, int LS>
__device__ __forceinline__ void probe_anchors(const AnchorParams &p, uint32_t cur, uint32_t left, uint32_t rcur, uint32_t rleft,
                                              uint32_t filter, uint32_t pairs, uint32_t (&a)[16 / S], bool (&hit)[16 / S],
                                              uint32_t (&lo)[16 / S])
{
	constexpr int NA = 16 / S;
	const uint32_t nw = p.filter_words;
	const int L = LS ? LS : p.len;
	const uint32_t amask = vg_mask32(L);
#pragma unroll
	for (int j = 0; j < NA; ++j) {
		/* bases [s0, s0 + L) relative to the chunk start, s0 = (j+1) S - L, possibly < 0 */
		const int sh = 2 * ((j + 1) * S - L);
		a[j] = sh >= 0 ? cur >> (sh & 31) : __funnelshift_r(left, cur, (sh + 32) & 31);
		if (L < 16 && !(sh >= 0 && sh + 2 * L == 32)) a[j] &= amask;
		uint32_t key = a[j];
		if (CANON) {
			/* a * rc(a): the same for an anchor and its reverse complement.  rc(a) is a window of
			 * the reverse-complemented words: base i of the chunk sits at 15 - i of rcur */
			const int rsh = 2 * (16 - (j + 1) * S);
			uint32_t r = rsh == 0 ? rcur : __funnelshift_r(rcur, rleft, rsh & 31);
			if (L < 16) r &= amask;
			key = a[j] * r;
		}
		const uint64_t prod = (uint64_t)vg_hash1(key) * nw;
		lo[j] = (uint32_t)prod;
		const uint32_t word = lds_word(filter, (uint32_t)(prod >> 32), p.c4);
		const uint32_t pm = pair_mask(pairs, lo[j], p.c4);
		hit[j] = (~word & pm) == 0;
	}
}

/* the check of a pending survivor: did it pass the second filter level as well? */
__device__ __forceinline__ bool passed2(uint32_t w, uint32_t pm) { return (~w & pm) == 0; }

/* ---- queue path ---- */

/* One tile.  `use` holds the raw chunk of tile t+1 (loaded three tiles ago); `fill` is the
 * buffer consumed by the previous tile and is refilled first thing with tile t+3; the only
 * consumer of loaded data is the pack at the very end. */
template <int S, bool CANON, int LS, bool INTERIOR>
__device__ __forceinline__ void scan_tile(const AnchorParams &p, Pipe &s, uint32_t t1, uint4 &fill, const uint4 &use, uint32_t filter,
                                          uint32_t pairs, uint2 *wq, uint32_t lane, uint32_t lt_mask)
{
	constexpr int NA = 16 / S;
	const uint32_t last = p.n_chunks - 1;
	fill = ld_stream<false>(p.chunks + (INTERIOR ? s.c + 96 : min(s.c + 96, last)));
	/* prefetch inside the span only: the next span's owner prefetches its own start */
	if (INTERIOR && s.t + VG_PF_BYTES / 512 < t1) l2_prefetch_ahead<VG_PF_BYTES>(p.chunks + s.c); /* this lane's chunk, some tiles on */
	/* one rotate serves both needs: lanes 1..31 get their left neighbour, lane 0 gets lane 31's
	 * chunk, which is its left neighbour in the NEXT tile */
	const uint32_t rot = __shfl_sync(FULL, s.cur, (lane + 31) & 31);
	const uint32_t left = lane == 0 ? s.carry : rot;
	uint32_t rrot = 0, rleft = 0;
	if (CANON && NA > 1) {
		rrot = __shfl_sync(FULL, s.rcur, (lane + 31) & 31);
		rleft = lane == 0 ? s.rcarry : rrot;
	}
	uint32_t a[NA], lo[NA];
	bool hit[NA];
	probe_anchors<S, CANON, LS>(p, s.cur, left, s.rcur, rleft, filter, pairs, a, hit, lo);
#pragma unroll
	for (int j = 0; j < NA; ++j) {
		if (!INTERIOR) hit[j] = hit[j] && s.c <= last;
		/* queue the survivors, compacted */
		if (__any_sync(FULL, hit[j])) {
			const uint32_t votes = __ballot_sync(FULL, hit[j]);
			if (hit[j]) wq[s.qn + __popc(votes & lt_mask)] = make_uint2(a[j], s.c * NA + j + 1);
			s.qn += __popc(votes);
		}
	}
	s.carry = rot; /* only lane 0's copy is ever used */
	s.cur = pack16(use);
	if (CANON) {
		s.rcarry = rrot;
		s.rcur = rc16(s.cur);
	}
	++s.t;
	s.c += 32;
}

/* The hot loop of the queue path: scan tiles until the span ends or 32 candidates are queued.
 * No calls, no table walks.  Three raw buffers rotate by a phase counter instead of by
 * register copies.  INTERIOR: every address touched (loads up to tile t1+2, prefetch some
 * tiles further) is inside the range, so nothing is clamped or predicated. */
template <int S, bool CANON, int LS, bool INTERIOR>
__device__ __forceinline__ void scan_tiles(const AnchorParams &p, Pipe &s, uint32_t t1, uint32_t filter, uint32_t pairs,
                                           uint2 *wq, uint32_t lane, uint32_t lt_mask)
{
	for (;;) {
		if (s.phase == 0) {
			scan_tile<S, CANON, LS, INTERIOR>(p, s, t1, s.w2, s.w0, filter, pairs, wq, lane, lt_mask);
			s.phase = 1;
			if (s.t >= t1 || s.qn >= 32) break;
		}
		if (s.phase == 1) {
			scan_tile<S, CANON, LS, INTERIOR>(p, s, t1, s.w0, s.w1, filter, pairs, wq, lane, lt_mask);
			s.phase = 2;
			if (s.t >= t1 || s.qn >= 32) break;
		}
		scan_tile<S, CANON, LS, INTERIOR>(p, s, t1, s.w1, s.w2, filter, pairs, wq, lane, lt_mask);
		s.phase = 0;
		if (s.t >= t1 || s.qn >= 32) break;
	}
}

/* ---- deferred path ---- */

/* what the tiles of a span share; the members a tile needs every time are pinned to
 * registers (opaque to ptxas, which would otherwise re-read them from the constant bank, or
 * recompute them, once per tile) */
struct DeferCtx {
	uint32_t filter, pairs; /* shared-memory addresses */
	uint32_t nw, four, nw2, rot_lane;
	const uint4 *chunks;
	const uint32_t *filter2;
	uint64_t keep;
	uint2 *wq, *vq;
	uint32_t lane, lt_mask;
	uint32_t vn, n_cand, n_hits;
};
/* a value ptxas must keep in a register: it went through shared memory and came back by a
 * volatile load, which cannot be repeated */
__device__ __forceinline__ uint32_t pin(uint32_t v, uint32_t scratch_sa)
{
	asm volatile("st.volatile.shared.u32 [%1], %0;\n\tld.volatile.shared.u32 %0, [%1];" : "+r"(v) : "r"(scratch_sa) : "memory");
	return v;
}
template <typename T> __device__ __forceinline__ const T *pin(const T *v, uint32_t scratch_sa)
{
	unsigned long long u = reinterpret_cast<unsigned long long>(v);
	asm volatile("st.volatile.shared.u64 [%1], %0;\n\tld.volatile.shared.u64 %0, [%1];" : "+l"(u) : "r"(scratch_sa) : "memory");
	return reinterpret_cast<const T *>(u);
}

/* the rare branch of the deferred path: queue the flagged anchors of the tile whose lane
 * chunk is `c`; the caller leaves the streaming loop when 32 are waiting */
template <int S, int LS>
__device__ __forceinline__ void defer_slow(const AnchorParams &p, Pipe &s, Pending<16 / S> &pd, uint32_t c, DeferCtx &x)
{
	constexpr int NA = 16 / S;
#pragma unroll
	for (int j = 0; j < NA; ++j) {
		const bool flag = passed2(pd.w[j], pd.pm[j]);
		const uint32_t votes = __ballot_sync(FULL, flag);
		if (flag) x.wq[s.qn + __popc(votes & x.lt_mask)] = make_uint2(pd.a[j], c * NA + j + 1);
		s.qn += __popc(votes);
	}
}

template <int S, int LS>
__device__ __forceinline__ void defer_check(const AnchorParams &p, Pipe &s, Pending<16 / S> &pd, uint32_t c, DeferCtx &x)
{
	constexpr int NA = 16 / S;
	bool any = false;
#pragma unroll
	for (int j = 0; j < NA; ++j) any = any || passed2(pd.w[j], pd.pm[j]);
	if (__any_sync(FULL, any)) defer_slow<S, LS>(p, s, pd, c, x);
}

/* the probes of one tile, waiting for their turn at the second filter level: the anchors,
 * their bit pairs, and the word to fetch for those that passed the first level
 * (VG_NO_WORD otherwise) */
#define VG_NO_WORD 0xFFFFFFFFu
template <int NA> struct Probed {
	uint32_t a[NA], pm[NA], word2[NA];
};

/* probe the anchors of the tile in s.cur */
template <int S, int LS, bool INTERIOR>
__device__ __forceinline__ void defer_probe(const AnchorParams &p, Pipe &s, Probed<16 / S> &nx, uint32_t rot, uint32_t rrot, DeferCtx &x)
{
	constexpr int NA = 16 / S;
	const int L = LS ? LS : p.len;
	const uint32_t amask = vg_mask32(L);
	/* rot / rrot: s.cur / s.rcur rotated by one lane.  One rotate serves both needs: lanes 1..31
	 * get their left neighbour, lane 0 gets lane 31's chunk, which is its left neighbour in the
	 * NEXT tile */
	const uint32_t left = x.lane == 0 ? s.carry : rot;
	const uint32_t rleft = x.lane == 0 ? s.rcarry : rrot;
#pragma unroll
	for (int j = 0; j < NA; ++j) {
		/* bases [s0, s0 + L) relative to the chunk start, s0 = (j+1) S - L, possibly < 0 */
		const int sh = 2 * ((j + 1) * S - L);
		uint32_t a = sh >= 0 ? s.cur >> (sh & 31) : __funnelshift_r(left, s.cur, (sh + 32) & 31);
		if (L < 16 && !(sh >= 0 && sh + 2 * L == 32)) a &= amask;
		/* a * rc(a): the same for an anchor and its reverse complement.  rc(a) is a window of
		 * the reverse-complemented words: base i of the chunk sits at 15 - i of rcur */
		const int rsh = 2 * (16 - (j + 1) * S);
		uint32_t r = rsh == 0 ? s.rcur : __funnelshift_r(s.rcur, rleft, rsh & 31);
		if (L < 16) r &= amask;
		const uint64_t prod = (uint64_t)vg_hash1(a * r) * x.nw;
		const uint32_t lo = (uint32_t)prod;
		const uint32_t word = lds_word(x.filter, (uint32_t)(prod >> 32), x.four);
		const uint32_t pm = pair_mask(x.pairs, lo, x.four);
		bool hit = (~word & pm) == 0;
		if (!INTERIOR) hit = hit && s.c <= p.n_chunks - 1;
		nx.a[j] = a;
nx.pm[j] = pm;
		nx.word2[j] = hit ? vg_mulhi(vg_pair_frac(lo), x.nw2) : VG_NO_WORD;
	}
	s.carry = rot; /* only lane 0's copy is ever used */
	s.rcarry = rrot;
}

/* One step of the deferred path's software pipeline.  On entry the probes of tile k-1 are in
 * `nx`, the second-level words its predecessor's survivors asked for are in flight in `pd`,
 * s.c is the lane's chunk of tile k (raw in `use`, loaded a step ago).  ptxas tracks every
 * global load of this kernel with ONE scoreboard, so whoever waits for a load waits for all
 * loads issued so far.  The step is therefore arranged around a single wait:
 *   first everything that consumes a load: the words of tile k-2's survivors are looked at,
 *     tile k is packed;
 *   then every new load: the words for tile k-1's survivors, and the refill of the raw buffer
 *     the previous step packed (tile k+1; the L2 prefetch runs further ahead);
 *   then the probes of tile k, which carry over to the next step.
 * Each load so gets one whole step to arrive.  (The refill sits behind the previous step's
 * closing branch, not next to the pack that emptied its buffer: there ptxas hoists it above
 * the pack's last reads and then needs a temporary -- and a copy that waits for the load.) */
template <int S, int LS, bool INTERIOR>
__device__ __forceinline__ void defer_step(const AnchorParams &p, Pipe &s, Pending<16 / S> &pd, Probed<16 / S> &nx, uint4 &fill,
                                           const uint4 &use, uint32_t t1, DeferCtx &x)
{
	constexpr int NA = 16 / S;
	const uint32_t last = p.n_chunks - 1;
	/* everything that consumes a load first ... */
	defer_check<S, LS>(p, s, pd, s.c - 64, x);
	s.cur = pack16(use);
	s.rcur = rc16(s.cur);
	/* ... then every new load.  A lane without a survivor keeps the word of an older one.  That
	 * can only cause a false alarm (the stale word happens to hold the new pair), never a miss:
	 * an anchor some pattern carries passes the first level, so its word is fetched afresh; and
	 * whatever is flagged is resolved exactly by drain_queue. */
#pragma unroll
	for (int j = 0; j < NA; ++j) {
		pd.a[j] = nx.a[j];
		pd.pm[j] = nx.pm[j];
		if (nx.word2[j] != VG_NO_WORD) pd.w[j] = __ldcg(x.filter2 + nx.word2[j]);
	}
	fill = ld_stream<true>(x.chunks + (INTERIOR ? s.c + 32 * VG_AHEAD : min(s.c + 32 * VG_AHEAD, last)));
	/* prefetch inside the span only: the next span's owner prefetches its own start */
	if (INTERIOR && s.t + VG_AHEAD + VG_PF_BYTES_DEFER / 512 < t1) l2_prefetch_ahead<VG_PF_BYTES_DEFER>(x.chunks + s.c + 32 * VG_AHEAD);
	{
		const uint32_t rot = __shfl_sync(FULL, s.cur, x.rot_lane);
		const uint32_t rrot = NA > 1 ? __shfl_sync(FULL, s.rcur, x.rot_lane) : 0u;
		defer_probe<S, LS, INTERIOR>(p, s, nx, rot, rrot, x);
	}
	++s.t;
	s.c += 32;
}

/* The hot loop of the deferred path: tiles from s.t up to t1, or until 32 anchors are queued
 * for the resolver (the caller runs it and comes back).  s.cur and s.w0 hold tiles s.t and
 * s.t + 1 on entry, s.c is the lane's chunk of tile s.t, pd is empty.  The pipeline probes one
 * tile past the last one whose buckets it requests (that tile is probed again by whoever scans
 * it).  On return tiles up to s.t - 1 are done except for the buckets in pd (tile s.t - 1,
 * lane chunk s.c - 32), which the caller looks at. */
template <int S, int LS, bool INTERIOR>
__device__ __forceinline__ void defer_span(const AnchorParams &p, Pipe &s, Pending<16 / S> &pd, uint32_t t1, DeferCtx &x)
{
	constexpr int NA = 16 / S;
	Probed<NA> nx;
	{
		const uint32_t rot = __shfl_sync(FULL, s.cur, x.rot_lane);
		const uint32_t rrot = NA > 1 ? __shfl_sync(FULL, s.rcur, x.rot_lane) : 0u;
		defer_probe<S, LS, INTERIOR>(p, s, nx, rot, rrot, x); /* tile s.t */
	}
	++s.t;
	s.c += 32;
	/* from here s.t counts probed tiles; a step probes tile s.t and requests the buckets of tile s.t - 1 */
#if VG_AHEAD == 2
	for (;;) { /* three raw buffers: a chunk is loaded two steps before it is packed */
		defer_step<S, LS, INTERIOR>(p, s, pd, nx, s.w2, s.w0, t1, x);
		if (s.t > t1 || s.qn >= 32) break;
		defer_step<S, LS, INTERIOR>(p, s, pd, nx, s.w0, s.w1, t1, x);
		if (s.t > t1 || s.qn >= 32) break;
		defer_step<S, LS, INTERIOR>(p, s, pd, nx, s.w1, s.w2, t1, x);
		if (s.t > t1 || s.qn >= 32) break;
	}
#else
	for (;;) {
		defer_step<S, LS, INTERIOR>(p, s, pd, nx, s.w1, s.w0, t1, x);
		if (s.t > t1 || s.qn >= 32) break;
		defer_step<S, LS, INTERIOR>(p, s, pd, nx, s.w0, s.w1, t1, x);
		if (s.t > t1 || s.qn >= 32) break;
	}
#endif
	/* back to "s.t = next tile to scan": the last probed tile (s.t - 1) has no buckets requested */
	--s.t;
	s.c -= 32;
}

/* Spans: warps take spans warp, warp + n_warps, ... so that at any time the resident warps read
 * one window of the stream that moves through it (kind to DRAM pages and the TLB).  All spans
 * but those of the last round have tiles_per_span tiles; what is left for the last round is cut
 * into one equal share per warp, so that nobody idles while others work off a full span. */
__device__ __forceinline__ uint32_t span_begin(const AnchorParams &p, uint32_t span)
{
	return span < p.n_full_spans ? span * p.tiles_per_span : p.n_full_spans * p.tiles_per_span + (span - p.n_full_spans) * p.tail_tiles;
}
__device__ __forceinline__ uint32_t span_end(const AnchorParams &p, uint32_t span)
{
	return min(span_begin(p, span) + (span < p.n_full_spans ? p.tiles_per_span : p.tail_tiles), p.n_tiles);
}

/* The streaming kernel.  One CTA per SM, persistent over spans of tiles (a tile = 32 chunks
 * of 16 bytes = one 128-bit load per lane).
 *   CANON  the filter holds strand-symmetric keys, which halves its load for large panels;
 *          small panels file both orientations and skip the reverse complement
 *   DEFER  survivors request their home bucket at once and inspect it during the next tile;
 *          only tag matches and chained buckets (both rare) go to the queue */
template <int S, bool CANON, bool DEFER, int LS>
__global__ void __launch_bounds__(Launch<S, DEFER>::kThreads, 1) anchor_scan_kernel(const __grid_constant__ AnchorParams p)
{
	extern __shared__ uint32_t s_filter[]; /* filter words | bit-pair table | candidate queues */
	using LC = Launch<S, DEFER>;
	constexpr int NA = 16 / S;
	const uint32_t nw = p.filter_words;        /* odd: what the hash is taken modulo */
	const uint32_t nwp = (nw + 3u) & ~3u;      /* the array is padded to whole 16-byte units */
	{ /* stage the filter */
		const uint4 *src = reinterpret_cast<const uint4 *>(p.filter);
		uint4 *dst = reinterpret_cast<uint4 *>(s_filter);
		for (uint32_t i = threadIdx.x; i < nwp / 4; i += blockDim.x) dst[i] = __ldg(src + i);
#if !VG_PAIR_ALU
		for (uint32_t i = threadIdx.x; i < VG_PAIRS; i += blockDim.x) s_filter[nwp + i] = vg_pair_mask(i);
#endif
	}
	__syncthreads();
	const uint32_t filter_sa = (uint32_t)__cvta_generic_to_shared(s_filter);
	const uint32_t pairs_sa = filter_sa + nwp * 4;

	const uint32_t lane = threadIdx.x & 31;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint2 *const wq = reinterpret_cast<uint2 *>(s_filter + nwp + VG_PAIR_TABLE_BYTES / 4) + (threadIdx.x >> 5) * LC::kQueue;
	uint2 *const vq = wq + LC::kQueue - 32; /* the last 32 entries: tag matches awaiting verification */
	const uint32_t warps_per_cta = blockDim.x >> 5;
	const uint32_t warp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
	const uint32_t n_warps = gridDim.x * warps_per_cta;
	const uint32_t last = p.n_chunks - 1;
	Pipe s;
	s.qn = 0;
	s.t = 0;
	s.phase = 0;
	s.rcur = s.rcarry = 0;
	uint32_t t1 = 0;
	bool interior = false;

	/* open span number `span`: prime the register pipeline (tile t is scanned while t+1 is being
	 * packed and t+2 is in flight; the L2 prefetch runs 8 tiles ahead of that) */
	auto open_span = [&](uint32_t t0, uint32_t t_end) {
		s.t = t0;
		t1 = t_end;
		/* Interior span: everything the pipeline touches (loads up to tile t1+2, prefetch 8 tiles
		 * beyond) lies inside the range.  Otherwise loads are clamped to the last chunk: a
		 * chunk past the end is never scanned, and as a right neighbour it can only create a
		 * false candidate, which the bounds check of the verification rejects. */
		constexpr int kPf = Knobs<DEFER>::kPrefetch;
		interior = (uint64_t)(t1 + 5 + kPf / 512) * 32 <= p.n_chunks;
		s.c = s.t * 32 + lane;
		if (interior && kPf) l2_prefetch(reinterpret_cast<const uint8_t *>(p.chunks + s.t * 32) + lane * (kPf / 32));
		s.cur = pack16(ld_stream<DEFER>(p.chunks + min(s.c, last)));
		s.w0 = ld_stream<DEFER>(p.chunks + min(s.c + 32, last));
		if (!DEFER || VG_AHEAD == 2) s.w1 = ld_stream<DEFER>(p.chunks + min(s.c + 64, last));
		s.phase = 0;
		/* the chunk before the span: the last one of the previous span, or of the previous
		 * launch range; nothing ('\n's) at the very start of the stream */
		const uint32_t c0 = s.t * 32;
		uint4 before = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
		if (c0 > 0 || p.range_lo > 0) before = __ldg(p.chunks + c0 - 1); /* chunks[-1] exists when range_lo > 0 */
		s.carry = pack16(before);
		if (CANON) {
			s.rcur = rc16(s.cur);
			s.rcarry = rc16(s.carry);
		}
	};

	if (DEFER) {
		DeferCtx x;
		const uint32_t z = (uint32_t)__cvta_generic_to_shared(vq + lane); /* scratch: this lane's verify-queue entry, not in use yet */
		x.filter = pin(filter_sa, z), x.pairs = pin(pairs_sa, z), x.keep = p.keep, x.wq = wq, x.vq = vq, x.lane = lane, x.lt_mask = lt_mask;
		x.nw = pin(nw, z), x.four = pin(p.c4, z), x.nw2 = pin(p.filter2_words, z), x.rot_lane = pin((lane + 31) & 31, z);
		x.chunks = pin(p.chunks, z), x.filter2 = pin(p.filter2, z);
		x.vn = x.n_cand = x.n_hits = 0;
		Pending<NA> pd;
#pragma unroll
		for (int j = 0; j < NA; ++j) {
			pd.w[j] = pd.a[j] = 0;
			pd.pm[j] = 1;
		}
		auto drain_full = [&]() {
			while (s.qn >= 32) {
				s.qn -= 32;
				x.n_cand += 32;
				__syncwarp();
				x.n_hits += drain_queue<S, true, true>(p, wq, s.qn, 32, vq, x.vn, lane, lt_mask);
				__syncwarp();
			}
		};
		for (uint32_t span = warp; span < p.n_spans; span += n_warps) {
			open_span(span_begin(p, span), span_end(p, span));
			for (;;) {
				if (interior) defer_span<S, LS, true>(p, s, pd, t1, x);
				else defer_span<S, LS, false>(p, s, pd, t1, x);
				/* the survivors of the last tile scanned; the queue takes one tile's worth above 31 entries */
				drain_full();
				defer_check<S, LS>(p, s, pd, s.c - 32, x);
				drain_full();
#pragma unroll
				for (int j = 0; j < NA; ++j) pd.w[j] = 0;
				if (s.t >= t1) break;
				open_span(s.t, t1); /* back into the span where the resolver interrupted it */
			}
		}
		if (s.qn) {
			x.n_cand += s.qn;
			__syncwarp();
			x.n_hits += drain_queue<S, true, true>(p, wq, 0, s.qn, vq, x.vn, lane, lt_mask);
			__syncwarp();
		}
		if (x.vn) x.n_hits += verify_batch<S, true>(p, vq, x.vn, lane);
		if (lane == 0 && (x.n_cand | x.n_hits)) {
			atomicAdd(&p.stats[ST_CANDIDATES], (unsigned long long)x.n_cand);
			atomicAdd(&p.stats[ST_HITS], (unsigned long long)x.n_hits);
		}
		return;
	}

	uint32_t vn = 0, n_cand = 0, n_hits = 0, span = warp;
	for (;;) {
		if (s.t >= t1 && span < p.n_spans) {
			open_span(span_begin(p, span), span_end(p, span));
			span += n_warps;
		}
		const bool finished = s.t >= t1; /* no span left */
		if (!finished) {
			if (interior) scan_tiles<S, CANON, LS, true>(p, s, t1, filter_sa, pairs_sa, wq, lane, lt_mask);
			else scan_tiles<S, CANON, LS, false>(p, s, t1, filter_sa, pairs_sa, wq, lane, lt_mask);
		}
		if (s.qn >= 32 || (finished && s.qn)) { /* the one place candidates are resolved */
			const uint32_t n = min(s.qn, 32u);
			s.qn -= n;
			n_cand += n;
			__syncwarp();
			n_hits += drain_queue<S, CANON, false>(p, wq, s.qn, n, vq, vn, lane, lt_mask);
			__syncwarp();
		} else if (finished) break;
	}
	if (vn) n_hits += verify_batch<S, DEFER>(p, vq, vn, lane);
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

/* span length of the persistent kernel, in tiles (128 KiB): long enough to amortise opening a
 * span, short enough that the resident warps read one compact, moving window of the stream */
#ifndef VG_SPAN_TILES
#define VG_SPAN_TILES 256u
#endif

/* one instantiation of the anchor kernel as the launcher sees it */
struct KernelForm {
	void (*fn)(const AnchorParams);
	int threads;
	size_t queue_bytes;
};

template <int S, bool CANON, bool DEFER, int LS> static KernelForm form_of()
{
	return KernelForm{anchor_scan_kernel<S, CANON, DEFER, LS>, Launch<S, DEFER>::kThreads, (size_t)Launch<S, DEFER>::kQueueBytes};
}

template <int S, int LS> static KernelForm form_of_plan(bool canon, bool defer)
{
	if (canon && defer) {
		if constexpr (VG_DEFER_OK(S)) return form_of<S, true, true, LS>();
		else return KernelForm{nullptr, 0, 0};
	}
	if (canon) return form_of<S, true, false, LS>();
	return form_of<S, false, false, LS>();
}

/* the headline plans (k = 15, 21, 31) get the anchor length as a compile-time constant */
static KernelForm select_form(int stride, int len, bool canon, bool defer)
{
	switch (stride) {
	case 1: return form_of_plan<1, 0>(canon, defer);
	case 2: return form_of_plan<2, 0>(canon, defer);
	case 4: return len == 12 ? form_of_plan<4, 12>(canon, defer) : form_of_plan<4, 0>(canon, defer);
	case 8: return len == 14 ? form_of_plan<8, 14>(canon, defer) : form_of_plan<8, 0>(canon, defer);
	case 16: return len == 16 ? form_of_plan<16, 16>(canon, defer) : form_of_plan<16, 0>(canon, defer);
	default: return KernelForm{nullptr, 0, 0};
	}
}

cudaError_t kernels_prepare(const ScanArgs &a)
{
	const KernelForm f = select_form(a.stride, a.len, a.canon != 0, a.defer != 0);
	if (!f.fn) return cudaErrorInvalidValue;
	return cudaFuncSetAttribute(reinterpret_cast<const void *>(f.fn), cudaFuncAttributeMaxDynamicSharedMemorySize, VG_SMEM_BUDGET);
}

int kernels_threads(const ScanArgs &a) { return select_form(a.stride, a.len, a.canon != 0, a.defer != 0).threads; }

static cudaError_t launch_form(const KernelForm &f, const AnchorParams &p0, int n_sm, cudaStream_t stream)
{
	AnchorParams p = p0;
	const uint32_t warps = (uint32_t)f.threads / 32;
	const size_t smem = (size_t)((p.filter_words + 3u) & ~3u) * 4 + VG_PAIR_TABLE_BYTES + f.queue_bytes;
	const uint32_t resident_warps = (uint32_t)n_sm * warps;
	const uint32_t cap = VG_SPAN_TILES;
	uint32_t tps = (p.n_tiles + resident_warps * 4u - 1) / (resident_warps * 4u);
	if (tps < 1) tps = 1;
	if (tps <= cap) { /* small input: at most four spans per warp, the last one possibly short */
		p.tiles_per_span = p.tail_tiles = tps;
		p.n_spans = p.n_full_spans = (p.n_tiles + tps - 1) / tps;
	} else { /* whole rounds of full spans, then one round of equal shares of the rest */
		p.tiles_per_span = cap;
		p.n_full_spans = p.n_tiles / (resident_warps * cap) * resident_warps;
		const uint32_t rest = p.n_tiles - p.n_full_spans * cap;
		p.tail_tiles = rest ? (rest + resident_warps - 1) / resident_warps : 1;
		p.n_spans = p.n_full_spans + (rest + p.tail_tiles - 1) / p.tail_tiles;
	}
	uint32_t ctas = (p.n_spans + warps - 1) / warps;
	if (ctas > (uint32_t)n_sm) ctas = (uint32_t)n_sm;
	void *args[] = {&p};
	return cudaLaunchKernel(reinterpret_cast<const void *>(f.fn), dim3(ctas), dim3((unsigned)f.threads), args, smem, stream);
}

__global__ void make_policy_kernel(uint64_t *out) { *out = l2_keep_policy(); }

cudaError_t kernels_make_policy(uint64_t *policy)
{
	uint64_t *d = nullptr;
	cudaError_t e = cudaMalloc(&d, sizeof *d);
	if (e != cudaSuccess) return e;
	make_policy_kernel<<<1, 1>>>(d);
	e = cudaMemcpy(policy, d, sizeof *d, cudaMemcpyDeviceToHost);
	cudaFree(d);
	return e;
}

cudaError_t launch_anchor_scan(const ScanArgs &a, int n_sm, cudaStream_t stream)
{
	const KernelForm f = select_form(a.stride, a.len, a.canon != 0, a.defer != 0);
	if (!f.fn) return cudaErrorInvalidValue;
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
		p.tiles_per_span = p.n_spans = p.n_full_spans = p.tail_tiles = 0;
		p.counts = a.counts;
		p.stats = a.stats;
		p.filter = a.filter;
		p.filter_words = a.filter_words;
		p.filter2 = a.filter2;
		p.filter2_words = a.filter2_words;
		p.buckets = reinterpret_cast<const uint4 *>(a.buckets);
		p.n_buckets = a.n_buckets;
		p.slots = a.slots;
		p.k = a.k;
		p.len = a.len;
		p.keep = a.keep_policy;
		p.c4 = 4;
		const cudaError_t e = launch_form(f, p, n_sm, stream);
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
