// Counting-mode microbenchmarks for the next step of the design: how
// fast are the memory operations the scan and the flush are made of, in isolation?
//   scatter8   one 8-byte store per key into R lists (what kc_scan_kernel<KC_PARTITION> does)
//   scatter32  one 32-byte sector (two 16-byte stores) per four keys into R lists
//   cursor     one 64-bit atomicAdd with return on R cursors, 256 bytes apart
//   l2cas      load + compare-and-swap per key into a table slice of S MiB (L2-resident for S <= 64)
//   smemcas    the same into a shared-memory table of 16 K slots per CTA
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/exp/kc_store_exp tools/exp/kc_store_exp.cu
// Run:   tools/exp/kc_store_exp [keys in millions = 512]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
	x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
	return x;
}

__global__ void scatter8(uint64_t *lists, uint32_t rbits, uint64_t cap, uint64_t n)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint64_t h = mix(i), r = h & ((1ull << rbits) - 1);
		lists[r * cap + (i >> rbits) % cap] = h; /* position without an atomic: i / R is unique per region on average */
	}
}

__global__ void scatter32(uint64_t *lists, uint32_t rbits, uint64_t cap, uint64_t n)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride) {
		const uint64_t h = mix(i), r = h & ((1ull << rbits) - 1);
		ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(lists + r * cap + ((i >> rbits) * 4) % cap);
		dst[0] = make_ulonglong2(h, h + 1);
		dst[1] = make_ulonglong2(h + 2, h + 3);
	}
}

__global__ void cursor(unsigned long long *cur, uint32_t rbits, uint64_t n, unsigned long long *sink)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	unsigned long long acc = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		acc += atomicAdd(cur + (mix(i) & ((1ull << rbits) - 1)) * 32, 1ull);
	if (acc == 1) *sink = acc;
}

__device__ __forceinline__ void insert(unsigned long long *slice, uint64_t mask, uint64_t tag)
{
	uint64_t pos = (tag * 0x9E3779B97F4A7C15ull) >> 20 & mask;
	for (int t = 0; t < 64; ++t, pos = (pos + 1) & mask) {
		unsigned long long v = slice[pos];
		if (v == 0) {
			v = atomicCAS(slice + pos, 0ull, tag << 10 | 1);
			if (v == 0) return;
		}
		while (v >> 10 == tag) {
			if ((v & 1023) == 1023) return;
			const unsigned long long old = atomicCAS(slice + pos, v, v + 1);
			if (old == v) return;
			v = old;
		}
	}
}

// region-major like kc_flush_kernel: CTA b works on slice b / ctas_per_slice
__global__ void l2cas(unsigned long long *table, uint64_t slice_slots, uint32_t n_slices, uint32_t ctas_per_slice, uint64_t per_cta)
{
	const uint32_t s = blockIdx.x / ctas_per_slice;
	if (s >= n_slices) return;
	unsigned long long *slice = table + (uint64_t)s * slice_slots;
	const uint64_t base = (uint64_t)blockIdx.x * per_cta;
	for (uint64_t i = threadIdx.x; i < per_cta; i += blockDim.x) insert(slice, slice_slots - 1, mix(base + i) >> 24 | 1);
}

__global__ void __launch_bounds__(1024, 1) smemcas(uint64_t per_cta, unsigned long long *sink)
{
	extern __shared__ unsigned long long tab[];
	for (uint32_t i = threadIdx.x; i < 16384; i += blockDim.x) tab[i] = 0;
	__syncthreads();
	const uint64_t base = (uint64_t)blockIdx.x * per_cta;
	/* 8 K distinct keys per CTA: the table stays half full, most operations are increments */
	for (uint64_t i = threadIdx.x; i < per_cta; i += blockDim.x) insert(tab, 16383, (mix((base + i) & 8191) >> 24) | 1);
	__syncthreads();
	if (tab[threadIdx.x] == 12345) *sink = 1;
}

template <typename F> static float timed(F f)
{
	cudaEvent_t a, b;
	cudaEventCreate(&a), cudaEventCreate(&b);
	f();
	cudaDeviceSynchronize();
	cudaEventRecord(a);
	f();
	cudaEventRecord(b);
	cudaEventSynchronize(b);
	float ms = 0;
	cudaEventElapsedTime(&ms, a, b);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
	return ms;
}

int main(int argc, char **argv)
{
	const uint64_t n = (uint64_t)(argc > 1 ? atoll(argv[1]) : 512) << 20;
	const int grid = 148 * 8, block = 256;
	unsigned long long *sink;
	cudaMalloc(&sink, 8);
	for (uint32_t rbits : {8u, 12u, 16u}) {
		const uint64_t cap = ((n >> rbits) * 2 + 31) & ~31ull;
		uint64_t *lists;
		unsigned long long *cur;
		cudaMalloc(&lists, (cap << rbits) * 8);
		cudaMalloc(&cur, (256ull << rbits));
		cudaMemset(cur, 0, 256ull << rbits);
		float ms = timed([&] { scatter8<<<grid, block>>>(lists, rbits, cap, n); });
		printf("scatter8   R=2^%-2u  %7.2f ms  %6.1f G keys/s\n", rbits, ms, n / ms / 1e6);
		ms = timed([&] { scatter32<<<grid, block>>>(lists, rbits, cap, n); });
		printf("scatter32  R=2^%-2u  %7.2f ms  %6.1f G keys/s\n", rbits, ms, n / ms / 1e6);
		ms = timed([&] { cursor<<<grid, block>>>(cur, rbits, n, sink); });
		printf("cursor     R=2^%-2u  %7.2f ms  %6.1f G atomics/s\n", rbits, ms, n / ms / 1e6);
		cudaFree(lists), cudaFree(cur);
	}
	for (uint32_t mib : {4u, 16u, 64u, 256u}) {
		const uint64_t slice_slots = (uint64_t)mib << 17; /* MiB / 8 bytes */
		const uint32_t n_slices = 64, ctas_per_slice = 1024;
		const uint64_t per_cta = slice_slots / 4 / ctas_per_slice; /* slices end a quarter full */
		unsigned long long *table;
		cudaMalloc(&table, slice_slots * n_slices * 8);
		cudaMemset(table, 0, slice_slots * n_slices * 8);
		const uint64_t total = per_cta * ctas_per_slice * n_slices;
		cudaEvent_t a, b;
		cudaEventCreate(&a), cudaEventCreate(&b);
		cudaEventRecord(a);
		l2cas<<<n_slices * ctas_per_slice, block>>>(table, slice_slots, n_slices, ctas_per_slice, per_cta);
		cudaEventRecord(b);
		cudaEventSynchronize(b);
		float ms = 0;
		cudaEventElapsedTime(&ms, a, b);
		printf("l2cas      slice %3u MiB  %7.2f ms  %6.1f G inserts/s\n", mib, ms, total / ms / 1e6);
		cudaFree(table);
	}
	{
		cudaFuncSetAttribute(smemcas, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
		const uint64_t per_cta = n / 148;
		float ms = timed([&] { smemcas<<<148, 1024, 16384 * 8>>>(per_cta, sink); });
		printf("smemcas    16 K slots per CTA  %7.2f ms  %6.1f G inserts/s\n", ms, per_cta * 148 / ms / 1e6);
	}
	return 0;
}
