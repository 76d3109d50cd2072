/*
 * kcgpu_kernels.cuh -- launch interface between the C ABI of include/kcgpu.h and the sm_100a
 * kernels of the full k-mer counting mode, plus the table geometry both sides share.
 *
 * Table: n_slots 64-bit words, cut into 2^region_bits regions of 2^rslot_bits slots.  A k-mer
 * with hash h = hash64(canonical k-mer) (kc-c4.c:40-50) belongs to owner h mod n_parts; with
 * q = h div n_parts its region is the low region_bits of q (the reference's partition by hash
 * suffix, kc-c4.c:66) and the slot holds ((q >> region_bits) + 1) << 10 | count, the reference's
 * "key << KC_BITS | count" word (kc-c4.c:11-15,124-125) with the key moved up by one, so that
 * 0 is a free slot and nothing else: an entry may carry a count of 0 (yak-count's second pass
 * starts from entries without counts, yak-count.c:451-452).  Linear probing inside the region
 * from a multiplicative hash of the tag.  Nothing is lost: (owner, region, tag) give h back,
 * and hash64 is invertible.
 *
 * Region lists: like the reference (count_seq_buf files k-mers per partition, worker_for then
 * fills one partition's table at a time, kc-c4.c:64-72,116-128), the scan does not touch the
 * table: it appends q to the list of its region, and a second kernel walks the lists region by
 * region, so that the slice of the table it fills (at most 64 MiB) stays in L2 while it is
 * filled.  Random 8-byte updates of a table far larger than L2 cost ~130 bytes of DRAM traffic
 * each (profiles/r1_kc_scan_v0); through the lists a k-mer costs 16 bytes of streaming plus its
 * share of one pass over the table.
 *
 * One allocation per owner, so that one pointer (one CUDA IPC handle) names it all:
 *   [ table: n_slots x 8 B ][ lists: n_regions x list_cap x 8 B ][ cursors: n_regions x 256 B, one 64-bit count each ][ Bloom filter: 2^bloom_bits / 8 B ]
 * With several owners the first half of the list area is the owner's inbox (its cursor follows
 * the regions'): whoever finds a k-mer appends it there in sector-sized runs; at a flush the
 * owner files what arrived under its region in the second half and empties that region by region.
 */
#ifndef KCGPU_KERNELS_CUH
#define KCGPU_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KC_HD __host__ __device__ __forceinline__
#else
#define KC_HD static inline
#endif

namespace kcgpu {

enum { KC_COUNT_BITS = 10, KC_COUNT_MAX = 1023 }; /* kc-c4.c:11-12 */
enum { KC_MAX_PARTS = 16 };
enum { KC_MAX_PROBES = 8192 }; /* a region this crowded is reported as overflow, not walked for ever */
#ifndef KC_REGION_SLOT_BITS_DEFAULT
#define KC_REGION_SLOT_BITS_DEFAULT 23
#endif
/* slots per region at most: 64 MiB of table, which stays in the 126 MB L2 while the region's list is emptied into it
 * (compare-and-swap into slices of 4 / 16 / 64 / 256 MiB: 47 / 50 / 45 / 27 G per second, tools/exp/kc_store_exp.cu);
 * the fewer regions, the longer the runs a tile of k-mers leaves in each list */
enum { KC_REGION_SLOT_BITS = KC_REGION_SLOT_BITS_DEFAULT };
enum { KC_TILE_REGION_BITS = 12 }; /* the most regions a CTA sorts a tile of k-mers by in shared memory */
enum { KC_CURSOR_STRIDE = 32 };    /* 64-bit words between two regions' cursors: one 256-byte line each */
enum { KC_FLUSH_TILE = 2048 };     /* list entries per CTA of the list-insert kernel */
enum { KC_TAG_BITS = 64 - KC_COUNT_BITS - 1 }; /* the stored key is tag + 1 */
enum { KC_BLOOM_BLOCK_BITS = 9 };                /* a Bloom block is 512 bits = 64 bytes (yak-count.c:14-15) */

/* what the insert step does with a k-mer (yak-count.c:150-177) */
enum {
	KC_INS_COUNT = 0,  /* make an entry if there is none, count up to 1023: kc-c4, yak-count without -b      */
	KC_INS_CLAIM = 1,  /* yak-count's first pass with -b: an entry with a count of 0 for every k-mer the
	                      Bloom filter has seen before; the counts of this pass are thrown away anyway   */
	KC_INS_LOOKUP = 2, /* yak-count's second pass: count the k-mers that have an entry, skip the others  */
};

/* the insert step's mode and the owner-independent part of the Bloom geometry; the filter
 * itself lies behind the cursors of every owner's allocation */
struct InsertCtl {
	int mode;
	uint32_t bloom_bits;   /* log2 of the bits of one owner's filter; 0 = no filter */
	uint32_t bloom_hashes;
};

/* per-context counters the kernels add to (unsigned long long each) */
enum { KC_ST_KMERS = 0, KC_ST_NEW = 1, KC_ST_OVERFLOW = 2, KC_ST_DROPPED = 3, KC_ST_DIRECT = 4, KC_ST_N = 8 };

/* regions needed so that the tag of a 2k-bit hash fits beside the count */
KC_HD uint32_t kc_region_bits(int k) { return 2 * k > KC_TAG_BITS ? (uint32_t)(2 * k - KC_TAG_BITS) : 0u; }

KC_HD uint64_t kc_hash64(uint64_t key, uint64_t mask) /* kc-c4.c:40-50 */
{
	key = (~key + (key << 21)) & mask;
	key ^= key >> 24;
	key = (key * 265) & mask;
	key ^= key >> 14;
	key = (key * 21) & mask;
	key ^= key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

/* ---- extraction without a byte loop (the tile kernels; host + device so that tests/cpu_sim can run it) ----
 * The rolling words of count_seq_buf (kc-c4.c:74-90) for the 16 positions of a chunk are windows of
 * one packed word: 48 bytes -- the chunk and the 32 before it -- are packed to 2 bits per base once
 * (codes A0 C1 G2 T3, kc-c4.c:21-38), the forward word of the k-mer that ends at byte e is a
 * window of the pair-reversed packing, the reverse word a window of the complemented packing, and
 * "a run of k bases ends here" is a test on a 48-bit mask of the bytes that are not bases. */
#define KC_TILE_N 16 /* positions per thread: one 16-byte chunk */

KC_HD uint32_t kc_byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
	return __byte_perm(a, b, sel);
#else
	const uint64_t both = (uint64_t)b << 32 | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) r |= (uint32_t)(both >> (8 * (sel >> (4 * i) & 7)) & 0xFF) << (8 * i);
	return r;
#endif
}

KC_HD uint32_t kc_brev(uint32_t x)
{
#if defined(__CUDA_ARCH__)
	return __brev(x);
#else
	x = (x >> 1 & 0x55555555u) | (x & 0x55555555u) << 1;
	x = (x >> 2 & 0x33333333u) | (x & 0x33333333u) << 2;
	x = (x >> 4 & 0x0F0F0F0Fu) | (x & 0x0F0F0F0Fu) << 4;
	x = (x >> 8 & 0x00FF00FFu) | (x & 0x00FF00FFu) << 8;
	return x >> 16 | x << 16;
#endif
}

/* bits s .. s+31 of hi:lo, 0 <= s <= 31 */
KC_HD uint32_t kc_funnel_r(uint32_t lo, uint32_t hi, int s)
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(lo, hi, s);
#else
	return (uint32_t)(((uint64_t)hi << 32 | lo) >> s);
#endif
}

/* 16 bytes -> 16 two-bit codes, byte i in bits 2i..2i+1; (b >> 1) & 3 gives A0 C1 T2 G3, the fix-up
 * x ^ (x >> 1 & 0x5555...) turns that into A0 C1 G2 T3 */
KC_HD uint32_t kc_pack16(uint4 w)
{
	const uint32_t M = 0x00820820u; /* 2^23 + 2^17 + 2^11 + 2^5: gathers the four codes of a word in its top byte */
	const uint32_t p0 = (w.x & 0x06060606u) * M, p1 = (w.y & 0x06060606u) * M;
	const uint32_t p2 = (w.z & 0x06060606u) * M, p3 = (w.w & 0x06060606u) * M;
	const uint32_t x = kc_byte_perm(kc_byte_perm(p0, p1, 0x0073), kc_byte_perm(p2, p3, 0x0073), 0x5410);
	return x ^ (x >> 1 & 0x55555555u);
}

/* the order of the 16 two-bit fields reversed */
KC_HD uint32_t kc_rev16(uint32_t x)
{
	const uint32_t y = kc_brev(x);
	return (y >> 1 & 0x55555555u) | (y & 0x55555555u) << 1;
}

/* four bytes -> four bits: bit i set when byte i is none of A C G T U a c g t u (the strict
 * table, kc-c4.c:21-38).  With c = bits 2..1 of the byte: bit 7 clear, bit 6 set, bit 3 clear,
 * bit 4 set exactly for T / U (c = 2), bit 0 set unless c = 2; bit 5 is the case. */
KC_HD uint32_t kc_not_base4(uint32_t w)
{
	const uint32_t is2 = (w >> 2) & ~(w >> 1) & 0x01010101u;
	uint32_t bad = (w ^ 0x40404040u) & 0xC8C8C8C8u;       /* bits 7, 6, 3 */
	bad |= ((w >> 4) ^ is2) & 0x01010101u;                 /* bit 4 against c == 2 */
	bad |= ~(w | is2) & 0x01010101u;                       /* bit 0 */
	bad |= bad >> 3;                                       /* bits 3, 6, 7 -> bits 0, 3, 4 */
	bad |= bad >> 4;                                       /* bits 4, 7 (and what they took in) -> bits 0, 3 */
	bad = (bad | bad >> 3) & 0x01010101u;                  /* everything in bit 0 of its byte */
	return (bad * 0x00204081u) >> 21 & 0xFu;               /* bits 0, 8, 16, 24 -> 21, 22, 23, 24 */
}

KC_HD uint32_t kc_not_base16(uint4 w)
{
	return kc_not_base4(w.x) | kc_not_base4(w.y) << 4 | kc_not_base4(w.z) << 8 | kc_not_base4(w.w) << 12;
}

/* what the extraction needs of k */
struct Extract {
	uint64_t mask, lim; /* the k bytes that end at byte e of the window are bases: (inv << (63 - e)) < lim */
	int down;
};
KC_HD Extract kc_extract_of(int k)
{
	Extract x;
	x.mask = (1ull << 2 * k) - 1ull;
	x.lim = 1ull << (64 - k);
	x.down = 64 - 2 * k;
	return x;
}

/* hash64 of the canonical k-mer that ends at each of the 16 bytes of chunk c (kc-c4.c:74-90); returns
 * the positions where one does (a run of k bases ends there).  A chunk at or behind `end` gives none. */
KC_HD uint32_t kc_extract16(const uint4 *chunks, const uint64_t c, const uint64_t end, const Extract &x, uint64_t (&h)[KC_TILE_N])
{
	const bool live = c < end;
	uint4 sep;
	sep.x = sep.y = sep.z = sep.w = 0x0A0A0A0Au;
#if defined(__CUDA_ARCH__)
	const uint4 w0 = live && c >= 2 ? __ldg(chunks + c - 2) : sep;
	const uint4 w1 = live && c >= 1 ? __ldg(chunks + c - 1) : sep;
	const uint4 own = live ? __ldg(chunks + c) : sep;
#else
	const uint4 w0 = live && c >= 2 ? chunks[c - 2] : sep;
	const uint4 w1 = live && c >= 1 ? chunks[c - 1] : sep;
	const uint4 own = live ? chunks[c] : sep;
#endif
	/* byte i of the 48-byte window (i = 32 + j for position j of the chunk): code in bits 2i of P,
	 * in bits 2 (47 - i) of R; bit i of `inv` set when it is not a base */
	const uint32_t P0 = kc_pack16(w0), P1 = kc_pack16(w1), P2 = kc_pack16(own);
	const uint32_t R0 = kc_rev16(P2), R1 = kc_rev16(P1), R2 = kc_rev16(P0);
	const uint64_t inv = (uint64_t)(kc_not_base16(w0) | kc_not_base16(w1) << 16) | (uint64_t)kc_not_base16(own) << 32;
	uint32_t ok = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int j = 0; j < KC_TILE_N; ++j) {
		/* forward word: byte 32 + j in bits 0..1, older bases above it (kc-c4.c:83) */
		const int sf = 2 * (15 - j);
		const uint64_t fw = ((uint64_t)kc_funnel_r(R1, R2, sf) << 32 | kc_funnel_r(R0, R1, sf)) & x.mask;
		/* reverse word: the complement of the 32 bases that end at byte 32 + j, the newest on top (kc-c4.c:84) */
		const int sr = 2 * (j + 1);
		const uint64_t y = sr < 32 ? (uint64_t)kc_funnel_r(P1, P2, sr) << 32 | kc_funnel_r(P0, P1, sr) : (uint64_t)P2 << 32 | P1;
		const uint64_t rv = ~y >> x.down;
		h[j] = kc_hash64(fw < rv ? fw : rv, x.mask);
		if ((inv << (31 - j)) < x.lim) ok |= 1u << j;
	}
	return ok;
}

/* ---- where a tile's entries go (the tile kernels; host + device so that tests/cpu_sim can check the arithmetic) ----
 * A region whose c entries stand at places lbase .. lbase + c - 1 of the sorted tile has reserved positions
 * g .. g + c - 1 of its list (g = what its cursor held).  One 64-bit word per region tells entry i of the tile
 * where to go: word + i is, if the whole run fits the list, KC_TILE_BIAS + its index in the list area
 * (region * stride + g + i - lbase), else -- bit 63 set -- KC_TILE_BIAS + its position in the region's list
 * (g + i - lbase), which the writer compares with the capacity itself.  The bias keeps word + i free of
 * borrows into bit 63 (lbase <= KC_TILE_BIAS). */
#define KC_TILE_BIAS 8192ull /* entries of a tile: 512 threads x 16 positions */

KC_HD unsigned long long kc_tile_word(unsigned long long g, uint32_t c, uint32_t lbase, uint32_t region, uint64_t cap, uint64_t stride)
{
	const unsigned long long rel = g + KC_TILE_BIAS - lbase;
	return g + c <= cap ? (unsigned long long)region * stride + rel : rel | 1ull << 63;
}
/* entry i: true and its index in the list area, or false and its position in its region's list */
KC_HD bool kc_tile_fits(unsigned long long word, uint32_t i, uint64_t *where)
{
	const unsigned long long at = word + i;
	*where = (at & ~(1ull << 63)) - KC_TILE_BIAS;
	return !(at >> 63);
}

struct CountArgs {
	const uint8_t *bytes; /* stream: reads separated by '\n', 16-byte aligned */
	uint64_t n_bytes;     /* multiple of 16                                    */
	uint64_t first_chunk, end_chunk; /* the 16-byte chunks this launch owns (it reads up to two before them) */
	uint64_t n_slots, list_cap;      /* geometry of every owner's allocation; list_cap 0 = no lists */
	int k;
	uint32_t n_parts;     /* owners of the hash space                          */
	uint32_t region_bits, rslot_bits;
	uint64_t *tables[KC_MAX_PARTS]; /* table of every owner (peer memory over NVLink for the others) */
	unsigned long long *stats;
	InsertCtl ctl;
	/* extract-only form */
	uint64_t *out_keys;   /* [part * cap_per_part + i] */
	uint64_t cap_per_part;
	uint32_t *part_counts;
};

struct InsertArgs {
	const uint64_t *hashed; /* hash64 values owned by this table */
	uint64_t n;
	uint32_t n_parts;
	uint32_t region_bits, rslot_bits;
	uint64_t *table;
	uint64_t n_slots, list_cap; /* where the Bloom filter lies */
	InsertCtl ctl;
	unsigned long long *stats;
};

/* asynchronous launches on `stream` */
cudaError_t launch_count(const CountArgs &a, cudaStream_t stream);   /* extract + insert straight into the tables */
cudaError_t launch_partition(const CountArgs &a, cudaStream_t stream); /* extract + append to the region lists (one owner) */
cudaError_t launch_push(const CountArgs &a, cudaStream_t stream);      /* extract + append to the owners' inboxes (several owners) */
/* insert what the region lists hold (`cap` entries per region at most, one cursor per region)
 * into the table; the cursors are left as they are */
cudaError_t launch_flush(uint64_t *table, const uint64_t *lists, const unsigned long long *cursors, uint64_t cap,
                         uint32_t region_bits, uint32_t rslot_bits, uint32_t *bloom, const InsertCtl &ctl,
                         unsigned long long *stats, cudaStream_t stream);

/* several owners: what arrived in the inbox, filed under its region (straight to the table
 * where a region list is full) */
struct RouteArgs {
	const uint64_t *inbox;
	const unsigned long long *n_ptr; /* the inbox's cursor */
	uint64_t inbox_cap;
	uint64_t *lists;                 /* region lists, `cap` entries each */
	unsigned long long *cursors;
	uint64_t cap;
	uint64_t *table;
	uint32_t region_bits, rslot_bits;
	uint32_t *bloom;
	InsertCtl ctl;
	unsigned long long *stats;
};
cudaError_t launch_route(const RouteArgs &a, int n_sm, cudaStream_t stream);
cudaError_t launch_extract(const CountArgs &a, cudaStream_t stream); /* extract into per-owner lists */
cudaError_t launch_insert(const InsertArgs &a, int n_sm, cudaStream_t stream);
cudaError_t launch_histogram(const uint64_t *table, uint64_t n_slots, unsigned long long *hist256, int n_sm,
                             cudaStream_t stream);
/* yak-count's histogram (yak-count.c:205-239): 1024 bins, entries with a count of 0 in bin 0 */
cudaError_t launch_histogram1024(const uint64_t *table, uint64_t n_slots, unsigned long long *hist1024, int n_sm,
                                 cudaStream_t stream);

KC_HD uint64_t *kc_lists_of(uint64_t *base, uint64_t n_slots) { return base + n_slots; }
KC_HD unsigned long long *kc_cursors_of(uint64_t *base, uint64_t n_slots, uint64_t list_cap, uint32_t region_bits)
{
	return reinterpret_cast<unsigned long long *>(base + n_slots + (list_cap << region_bits));
}
/* one cursor per region and one more, the inbox's */
KC_HD uint64_t kc_cursor_bytes(uint32_t region_bits) { return (((uint64_t)1 << region_bits) + 1) * KC_CURSOR_STRIDE * 8; }
KC_HD uint64_t kc_bloom_bytes(uint32_t bloom_bits) { return bloom_bits ? (uint64_t)1 << (bloom_bits - 3) : 0; }
KC_HD uint64_t kc_alloc_bytes(uint64_t n_slots, uint64_t list_cap, uint32_t region_bits, uint32_t bloom_bits = 0)
{
	return (n_slots + (list_cap << region_bits)) * 8 + kc_cursor_bytes(region_bits) + kc_bloom_bytes(bloom_bits);
}
/* the Bloom filter of an allocation: behind the cursors */
KC_HD uint32_t *kc_bloom_of(uint64_t *base, uint64_t n_slots, uint64_t list_cap, uint32_t region_bits)
{
	return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(base + n_slots + (list_cap << region_bits)) + kc_cursor_bytes(region_bits));
}
/* several owners: the first half of the list area is the inbox, the second half the region lists */
KC_HD uint64_t kc_inbox_cap(uint64_t list_cap, uint32_t region_bits) { return (list_cap << region_bits) / 2; }
KC_HD unsigned long long *kc_inbox_cursor(uint64_t *base, uint64_t n_slots, uint64_t list_cap, uint32_t region_bits)
{
	return kc_cursors_of(base, n_slots, list_cap, region_bits) + ((uint64_t)KC_CURSOR_STRIDE << region_bits);
}

} // namespace kcgpu
#endif
