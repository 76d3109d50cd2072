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
#define VG_MAX_FILTER_WORDS 57344u /* 224 KiB of shared memory */
#define VG_MIN_FILTER_WORDS 1024u

/* exact-table slot: an oriented pattern k-mer (stream encoding) with the offset of the
 * anchor it is filed under.  16 bytes so one LDG.128 fetches it. */
typedef struct {
	uint64_t okey; /* oriented k-mer, VG_EMPTY_KEY if the slot is free */
	uint32_t val;  /* (pattern index << 1) | is_alt                     */
	uint32_t off;  /* anchor = (okey >> 2*off) & mask(L)                */
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

VG_HD uint32_t vg_canon32(uint32_t x, int L)
{
	uint32_t r = vg_rc32(x, L);
	return x < r ? x : r;
}

/* filter: word index and the two probe bits from one multiplicative hash */
VG_HD uint32_t vg_filter_hash(uint32_t canon_anchor) { return canon_anchor * 0x9E3779B1u; }
VG_HD uint32_t vg_filter_word(uint32_t h, uint32_t n_words)
{
#if defined(__CUDA_ARCH__)
	return __umulhi(h, n_words);
#else
	return (uint32_t)(((uint64_t)h * n_words) >> 32);
#endif
}
VG_HD uint32_t vg_filter_mask(uint32_t h) { return (1u << ((h >> 5) & 31)) | (1u << ((h >> 10) & 31)); }

/* exact table: home slot of a forward anchor */
VG_HD uint32_t vg_slot_home(uint32_t anchor, uint32_t slot_bits)
{
	return (anchor * 0xCC9E2D51u) >> (32 - slot_bits);
}

#endif
