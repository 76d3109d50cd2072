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

/* launch geometry of the anchor kernel for stride S, shared with the table builder because
 * the candidate queues and the filter split one shared-memory budget */
#define VG_THREADS(S) ((S) >= 4 ? 1024 : 256)              /* tiny k: fewer warps, deeper queues   */
#define VG_QUEUE_ENTRIES(S) (32 + 32 * (16 / (S)) + 32)    /* per warp: a drain's leftovers + one tile, + 32 tag matches awaiting verification */
#define VG_QUEUE_BYTES(S) ((VG_THREADS(S) / 32) * VG_QUEUE_ENTRIES(S) * 8)
#define VG_FILTER_BUDGET_WORDS(S) ((uint32_t)(((VG_SMEM_BUDGET - VG_QUEUE_BYTES(S) - 128) / 4) & ~3))

/* exact-table slot: an oriented pattern k-mer (stream encoding) with the offset of the
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

/* Blocked Bloom filter: one 32-bit word per key, two bits in it.  The word comes from the
 * top of one multiplicative hash, the two bit positions from the top ten bits of a second
 * and a third one.  On the device 1 << b comes from a 32-word shared-memory table: bank b
 * holds entry b, so the look-up never conflicts, and unlike a shift it costs no slot of the
 * integer ALU pipe, the kernel's scarcest resource. */
VG_HD uint32_t vg_hash1(uint32_t key) { return key * 0x9E3779B1u; }
VG_HD uint32_t vg_hash2(uint32_t key) { return key * 0x85EBCA6Bu; }
VG_HD uint32_t vg_mulhi(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
	return __umulhi(a, b);
#else
	return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
VG_HD uint32_t vg_filter_word(uint32_t key, uint32_t n_words) { return vg_mulhi(vg_hash1(key), n_words); }
VG_HD uint32_t vg_hash3(uint32_t h2) { return h2 * 0xC2B2AE35u; }
VG_HD uint32_t vg_filter_mask(uint32_t key)
{
	const uint32_t h2 = vg_hash2(key);
	return (1u << (h2 >> 27)) | (1u << (vg_hash3(h2) >> 27));
}

/* exact table: buckets of four 32-bit tags (one LDG.128) with the 16-byte payloads in a
 * parallel array that is only touched when a tag matches.  Tag 0 marks a free slot; a slot
 * is filled at the first free position scanning on from the home bucket, so a lookup stops
 * at the first free slot it meets.  The tag keeps the low 31 bits of the forward anchor
 * (all of it for L <= 15); the payload decides. */
VG_HD uint32_t vg_tag(uint32_t anchor) { return anchor | 0x80000000u; }
VG_HD uint32_t vg_bucket_home(uint32_t anchor, uint32_t bucket_bits)
{
	return (anchor * 0xCC9E2D51u) >> (32 - bucket_bits);
}

#endif
