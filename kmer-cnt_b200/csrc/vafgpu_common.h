/*
 * vafgpu_common.h -- encodings, hashes and table layouts shared by the host table builder
 * and the device kernels.
 *
 * Two encodings are in play:
 *   reference encoding (the C ABI, the recipe kernel): A0 C1 G2 T3, first base most
 *     significant (vaf-counter.c:117-127);
 *   stream encoding (the anchor-filter kernel): code = (ascii >> 1) & 3, i.e. A0 C1 T2 G3,
 *     first base LEAST significant, because that is what one AND and one multiply per four
 *     ASCII bytes produce.  Complement is code ^ 2.
 */
#ifndef VAFGPU_COMMON_H
#define VAFGPU_COMMON_H

#include <stdint.h>

#if defined(__CUDACC__)
#define VG_HD __host__ __device__ __forceinline__
#else
#define VG_HD static inline
#endif

#define VG_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define VG_SMEM_BUDGET (227 * 1024) /* dynamic shared memory per CTA on sm_100: filter + candidate queues */
#define VG_MIN_FILTER_WORDS 1024u

/* Launch geometry of the anchor kernel, shared with the table builder because the candidate
 * queues and the filter split one shared-memory budget.
 *   queue path    (small panels, or tiny k): survivors of the filter are rare; they are
 *                 compacted into a per-warp queue and resolved 32 at a time;
 *   deferred path (large panels, S >= 4): every tenth anchor survives, so each lane requests
 *                 its survivor's home bucket from L2 straight away and looks at it one tile
 *                 later; that needs registers, hence fewer threads. */
#ifndef VG_THREADS_DEFER
#define VG_THREADS_DEFER(S) ((S) >= 8 ? 1024 : 768)
#endif
#define VG_DEFER_OK(S) ((S) >= 4)
#define VG_THREADS(S, DEFER) ((DEFER) ? VG_THREADS_DEFER(S) : (S) >= 4 ? 1024 : 256) /* tiny k: fewer warps, deeper queues */
#define VG_QUEUE_ENTRIES(S) (32 + 32 * (16 / (S)) + 32)  /* per warp: a drain's leftovers + one tile, + 32 tag matches awaiting verification */
#define VG_QUEUE_BYTES(S, DEFER) ((VG_THREADS(S, DEFER) / 32) * VG_QUEUE_ENTRIES(S) * 8)
#define VG_FILTER_BUDGET_WORDS(S, DEFER) ((uint32_t)(((VG_SMEM_BUDGET - VG_QUEUE_BYTES(S, DEFER)) / 4) & ~3))

/* exact-table payload: an oriented pattern k-mer (stream encoding) with the offset of the
 * anchor it is filed under.  16 bytes so one LDG.128 fetches it. */
typedef struct {
	uint64_t okey; /* oriented k-mer, VG_EMPTY_KEY if the slot is free */
	uint32_t val;  /* (pattern index << 1) | is_alt                     */
	uint32_t off;  /* t: the anchor ends t bases before the k-mer does:
	                  anchor = (okey >> 2*(k-t-L)) & mask(L)            */
} vg_slot_t;

/* ---- the reference's hash (recipe kernel and its table) ---- */
VG_HD uint32_t vg_kmer_hash(uint64_t key) /* vaf-counter.c:56-63 */
{
	key ^= key >> 33;
	key *= 0xff51afd7ed558ccdULL;
	key ^= key >> 33;
	return (uint32_t)key;
}
VG_HD uint32_t vg_h2b(uint32_t hash, uint32_t bits) /* khashl.h:98 */
{
	return hash * 2654435769u >> (32 - bits);
}

/* ---- stream encoding helpers ---- */
VG_HD uint32_t vg_mask32(int nbases) { return nbases >= 16 ? 0xFFFFFFFFu : (1u << 2 * nbases) - 1u; }

/* reverse complement of an L-base word in stream encoding (L <= 16) */
VG_HD uint32_t vg_rc32(uint32_t x, int L)
{
#if defined(__CUDA_ARCH__)
	uint32_t r = __brev(x);
#else
	uint32_t r = x;
	r = (r >> 16) | (r << 16);
	r = ((r & 0xFF00FF00u) >> 8) | ((r & 0x00FF00FFu) << 8);
	r = ((r & 0xF0F0F0F0u) >> 4) | ((r & 0x0F0F0F0Fu) << 4);
	r = ((r & 0xCCCCCCCCu) >> 2) | ((r & 0x33333333u) << 2);
	r = ((r & 0xAAAAAAAAu) >> 1) | ((r & 0x55555555u) << 1);
#endif
	/* bit reversal also swapped the two bits of every base: swap them back */
	r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
	r ^= 0xAAAAAAAAu;       /* complement every base */
	return r >> (32 - 2 * L); /* the L bases sit at the top after reversal */
}

VG_HD uint32_t vg_mulhi(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
	return __umulhi(a, b);
#else
	return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

/* Filter key of an anchor.  Large panels (canon) use a function that is the same for an
 * anchor and its reverse complement, so one filter entry serves both strands: the upper half
 * of the product of the two, each moved to the top of its word -- symmetric like
 * min(a, rc(a)) but multiplies (FMA pipe) instead of a compare-select, and the shifts to the
 * top drop whatever a 32-bit window of the stream holds beyond the anchor, so the kernel never
 * masks (the integer ALU pipe is the scarce resource of the kernel). */
VG_HD uint32_t vg_filter_key(uint32_t a, int L, int canon)
{
	const int up = 32 - 2 * L;
	return canon ? vg_mulhi(a << up, vg_rc32(a, L) << up) : a;
}

/* One multiply chain places an anchor everywhere.  h = key * odd constant; the 64-bit product
 * h * n_words gives the filter word (high half) and a second, well mixed 32-bit value `lo`
 * (low half; n_words is odd, so key -> lo is a bijection); a third, g = high half of
 * h * another constant.  On the device: IMAD, IMAD.WIDE, IMAD.HI -- all on the FMA pipe.
 *   filter (shared memory)  word = hi(h * n_words), bits lo & 31 and g & 31: the kernel tests
 *                           them by shifting the word (the shifter takes the amount modulo 32)
 *   second level (L2)       word = 1 + hi(lo * (n_words2 - 2)), bit lo & 31; the first and the
 *                           last word of the array stay zero: anchors that failed the first
 *                           level read one of them
 *   exact table             home bucket = hi(lo * n_buckets)                                   */
#define VG_HASH_C1 0x9E3779B1u
#define VG_HASH_C2 0x85EBCA6Bu
VG_HD uint32_t vg_hash1(uint32_t key) { return key * VG_HASH_C1; }
VG_HD uint32_t vg_hash_g(uint32_t h) { return vg_mulhi(h, VG_HASH_C2); }
VG_HD uint32_t vg_filter_word(uint32_t key, uint32_t n_words) { return vg_mulhi(vg_hash1(key), n_words); }
VG_HD uint32_t vg_hash_lo(uint32_t key, uint32_t n_words) { return vg_hash1(key) * n_words; }
VG_HD uint32_t vg_filter_mask(uint32_t key, uint32_t n_words)
{
	return (1u << (vg_hash_lo(key, n_words) & 31u)) | (1u << (vg_hash_g(vg_hash1(key)) & 31u));
}
VG_HD uint32_t vg_filter2_word(uint32_t key, uint32_t n_words, uint32_t n_words2)
{
	return 1u + vg_mulhi(vg_hash_lo(key, n_words), n_words2 - 2u);
}
VG_HD uint32_t vg_filter2_mask(uint32_t key, uint32_t n_words) { return 1u << (vg_hash_lo(key, n_words) & 31u); }

/* Exact table: buckets of 16 bytes (one LDG.128) = three 32-bit tags + one control word.
 *   tag   the forward anchor (low 31 bits | bit 31 when L = 16, so that it never equals
 *         VG_FREE_TAG); occupied slots are a prefix of the bucket
 *   ctrl  bits 0..30: index in the payload array of the bucket's first entry (the entries of
 *         a bucket are consecutive there); bit 31: an entry that hashes to this bucket, or
 *         passed through it, lives further on -- keep looking in the next bucket
 * Entries are placed by linear probing over buckets from the home bucket, which both
 * strands of an anchor share when the filter key is strand-symmetric.  The payload decides. */
#define VG_FREE_TAG 0x7FFFFFFFu
#define VG_CTRL_MORE 0x80000000u
VG_HD uint32_t vg_tag(uint32_t anchor, int L) { return L >= 16 ? anchor | 0x80000000u : anchor; }
VG_HD uint32_t vg_bucket_home(uint32_t lo, uint32_t n_buckets) { return vg_mulhi(lo, n_buckets); }

#endif
