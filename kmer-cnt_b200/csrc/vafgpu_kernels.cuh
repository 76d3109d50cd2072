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
	int canon;             /* the filter holds canonical anchors */
	const uint32_t *filter;
	uint32_t filter_words;
	const uint32_t *tags;
	const vg_slot_t *slots;
	uint32_t bucket_bits;
	/* recipe kernel */
	const uint64_t *rkeys;
	const uint32_t *rvals;
	uint32_t rbits;
};

/* one-time per-device set-up (shared-memory opt-in); returns cudaSuccess or the error */
cudaError_t kernels_init_device(int n_sm);

/* asynchronous launches on `stream` */
cudaError_t launch_anchor_scan(const ScanArgs &a, int n_sm, cudaStream_t stream);
cudaError_t launch_recipe_scan(const ScanArgs &a, int n_sm, cudaStream_t stream);

} // namespace vafgpu
#endif
