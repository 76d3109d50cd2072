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
#define VG_SMEM_BUDGET (227 * 1024) /* dynamic shared memory per CTA on sm_100: filter + bit-pair table + candidate queues */
#define VG_MIN_FILTER_WORDS 1024u
#ifndef VG_PAIR_ALU
#define VG_PAIR_ALU 1               /* bit pair of a key: made with shifts (1) or read from a shared-memory table (0) */
#endif
#define VG_PAIRS 992u               /* ordered pairs of distinct bit positions in a 32-bit word */
#define VG_PAIR_TABLE_BYTES (VG_PAIR_ALU ? 0 : 4096)

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
#define VG_FILTER_BUDGET_WORDS(S, DEFER) \
	((uint32_t)(((VG_SMEM_BUDGET - VG_QUEUE_BYTES(S, DEFER) - VG_PAIR_TABLE_BYTES) / 4) & ~3))

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

/* Filter key of an anchor.  Large panels (canon) use a function that is the same for an
 * anchor and its reverse complement, so one filter entry serves both strands: the product
 * a * rc(a) mod 2^32 -- symmetric like min(a, rc(a)) but a multiply (FMA pipe) instead of a
 * compare-select (the integer ALU pipe is the scarce resource of the kernel). */
VG_HD uint32_t vg_filter_key(uint32_t a, int L, int canon) { return canon ? a * vg_rc32(a, L) : a; }

/* One multiply chain places an anchor everywhere.  h = key * odd constant; the 64-bit product
 * h * n_words gives the filter word (high half) and a second, well mixed 32-bit value (low
 * half) whose top bits pick the bit pair and the home bucket of the exact table.  On the
 * device this is IMAD, IMAD.WIDE, IMAD.HI, IMAD.HI: all on the FMA pipe. */
VG_HD uint32_t vg_mulhi(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
	return __umulhi(a, b);
#else
	return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
VG_HD uint32_t vg_hash1(uint32_t key) { return key * 0x9E3779B1u; }
VG_HD uint32_t vg_filter_word(uint32_t key, uint32_t n_words) { return vg_mulhi(vg_hash1(key), n_words); }
VG_HD uint32_t vg_hash_lo(uint32_t key, uint32_t n_words) { return vg_hash1(key) * n_words; }

/* Blocked Bloom filter: one 32-bit word per key, two bits in it (the same bit twice for one key
 * in 32).  VG_PAIR_ALU: the bit positions are the top two 5-bit fields of `lo` and the mask is
 * made with four shifts; otherwise the pair comes from a 992-entry table (shared memory on
 * the device): entry i * 31 + j names bits i and (j < i ? j : j + 1). */
VG_HD uint32_t vg_pair_index(uint32_t lo) { return vg_mulhi(lo, VG_PAIRS); }
VG_HD uint32_t vg_pair_mask(uint32_t idx)
{
	const uint32_t i = idx / 31u, j = idx % 31u;
	return (1u << i) | (1u << (j < i ? j : j + 1u));
}
#if VG_PAIR_ALU
VG_HD uint32_t vg_lo_mask(uint32_t lo) { return (1u << (lo >> 27)) | (1u << ((lo >> 22) & 31u)); }
VG_HD uint32_t vg_pair_frac(uint32_t lo) { return lo << 10; } /* the bits of lo the pair did not use */
#else
VG_HD uint32_t vg_lo_mask(uint32_t lo) { return vg_pair_mask(vg_pair_index(lo)); }
VG_HD uint32_t vg_pair_frac(uint32_t lo) { return lo * VG_PAIRS; } /* the low half of the same product: where in the pair's cell lo falls */
#endif
VG_HD uint32_t vg_filter_mask(uint32_t key, uint32_t n_words) { return vg_lo_mask(vg_hash_lo(key, n_words)); }

/* Second filter level (large panels): the same two bits in a word of a much bigger array that
 * lives in L2 (a few MB: >= 128 bits per key).  Only anchors that passed the on-chip filter
 * look at it, one 4-byte load each; what passes both goes to the exact table.  The word comes
 * from the part of the hash the pair did not use. */
VG_HD uint32_t vg_filter2_word(uint32_t key, uint32_t n_words, uint32_t n_words2)
{
	return vg_mulhi(vg_pair_frac(vg_hash_lo(key, n_words)), n_words2);
}

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
