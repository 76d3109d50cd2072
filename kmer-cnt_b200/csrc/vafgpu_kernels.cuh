/*
 * vafgpu_kernels.cuh -- launch interface between the C-ABI layer and the sm_100a kernels.
 */
#ifndef VAFGPU_KERNELS_CUH
#define VAFGPU_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "vafgpu_common.h"

namespace vafgpu {

/* per-device counters the kernels add to (unsigned long long each) */
enum { ST_CANDIDATES = 0, ST_HITS = 1, ST_KMERS = 2, ST_N = 4 };

struct ScanArgs {
	const uint8_t *bytes;  /* stream, 16-byte aligned                         */
	uint64_t n_bytes;      /* multiple of 16                                  */
	uint32_t *counts;      /* 2 * n_patterns                                  */
	unsigned long long *stats;
	int k;
	/* anchor-filter kernel */
	int stride, len;
	int canon;             /* the filter holds strand-symmetric keys */
	int defer;             /* deferred-lookup form (shared-memory split computed for it) */
	const uint32_t *filter;
	uint32_t filter_words;
	const uint32_t *filter2; /* second filter level (deferred form), L2-resident */
	uint32_t filter2_words;
	const uint32_t *buckets; /* exact table: four words per bucket */
	uint32_t n_buckets;    /* one more, empty, bucket follows the table               */
	const vg_slot_t *slots;
	uint64_t keep_policy;  /* from kernels_make_policy() on this device               */
	/* recipe kernel */
	const uint64_t *rkeys;
	const uint32_t *rvals;
	uint32_t rbits;
};

/* one-time per-device set-up: the L2 evict-last access policy the table loads carry */
cudaError_t kernels_make_policy(uint64_t *policy);

/* one-time per-device set-up of the anchor kernel instantiation these arguments select (its
 * dynamic shared-memory opt-in); call with the device current, before the first launch */
cudaError_t kernels_prepare(const ScanArgs &a);
int kernels_threads(const ScanArgs &a); /* CTA size of that instantiation */

/* asynchronous launches on `stream` */
cudaError_t launch_anchor_scan(const ScanArgs &a, int n_sm, cudaStream_t stream);
cudaError_t launch_recipe_scan(const ScanArgs &a, int n_sm, cudaStream_t stream);

} // namespace vafgpu
#endif
