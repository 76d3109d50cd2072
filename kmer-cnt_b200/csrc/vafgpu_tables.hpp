/*
 * vafgpu_tables.hpp -- host-side construction of the device lookup structures from the
 * flat (canonical key, value) list the C ABI receives.  Pure C++, no CUDA.
 */
#ifndef VAFGPU_TABLES_HPP
#define VAFGPU_TABLES_HPP

#include <cstdint>
#include <vector>

#include "vafgpu_common.h"

namespace vafgpu {

struct Plan {
	int k = 0;
	int stride = 1; /* S: anchors start at stream offsets that are multiples of S */
	int len = 0;    /* L: anchor length in bases, L <= 16 and L <= k - S + 1     */
};

Plan make_plan(int k);

/* Table of the literal recipe kernel: the reference's geometry (power of two >= 3 *
 * n_patterns buckets, vaf-counter.c:216, khashl.h:152-160), hash and probe order. */
struct RecipeTable {
	uint32_t bits = 2;
	std::vector<uint64_t> keys; /* VG_EMPTY_KEY when free */
	std::vector<uint32_t> vals;
};

/* Structures of the anchor-filter kernel. */
struct AnchorTables {
	Plan plan;
	std::vector<uint32_t> filter; /* blocked Bloom filter over the anchors (padded to x4)   */
	uint32_t filter_words = 0;    /* words it hashes into: odd                              */
	bool canon = false;           /* filter keys are canonical anchors (large panels) rather
	                                 than both orientations of every anchor (small panels)  */
	bool defer = false;           /* scanned by the deferred-lookup form of the kernel       */
	int threads = 0;              /* CTA size the shared-memory split was computed for      */
	std::vector<uint32_t> filter2;/* second level of the filter, L2-resident (deferred form) */
	uint32_t n_buckets = 0;       /* buckets of the exact table                             */
	std::vector<uint32_t> buckets;/* four words per bucket: three tags + control word        */
	std::vector<vg_slot_t> slots; /* payloads, a bucket's entries consecutive               */
	uint32_t n_entries = 0;       /* (oriented key, offset) pairs filed                  */
	uint32_t n_filter_keys = 0;   /* distinct keys in the filter                         */
};

/* keys are canonical k-mers in the reference encoding; duplicates keep the first value */
void build_recipe_table(int k, const uint64_t *keys, const uint32_t *vals, uint32_t n,
                        uint32_t n_patterns, RecipeTable &out);
void build_anchor_tables(int k, const uint64_t *keys, const uint32_t *vals, uint32_t n,
                         AnchorTables &out);

/* reference-encoded k-mer -> stream-encoded oriented k-mer (forward) and its reverse
 * complement, both as they would appear in a read */
uint64_t ref_to_stream(uint64_t ref_key, int k);
uint64_t stream_revcomp(uint64_t okey, int k);

} // namespace vafgpu
#endif
