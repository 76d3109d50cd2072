// Counting-mode microbenchmarks for the next step of the design: how
// fast are the memory operations the scan and the flush are made of, in isolation?
//   scatter8   one 8-byte store per key into R lists (what kc_scan_kernel<KC_PARTITION> does)
//   scatter32  one 32-byte sector (two 16-byte stores) per four keys into R lists
//   cursor     one 64-bit atomicAdd with return on R cursors, 256 bytes apart
//   l2cas      load + compare-and-swap per key into a table slice of S MiB (L2-resident for S <= 64)
//   smemcas    the same into a shared-memory table of 16 K slots per CTA
//   l2red      load + reduction (no return value) per key into an L2-resident slice: the cost of an increment
//              when the entry is known to exist
//   part       one radix-partition pass with shared-memory staging: a CTA ranks a tile of keys by a digit of
//              RB bits (shared-memory atomics), sorts the tile in shared memory, reserves one run per digit
//              with one cursor atomic and writes the runs out -- whole sectors instead of 8-byte stores
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

__global__ void l2red(unsigned long long *table, uint64_t slice_slots, uint32_t n_slices, uint32_t ctas_per_slice, uint64_t per_cta,
                      unsigned long long *sink)
{
	const uint32_t s = blockIdx.x / ctas_per_slice;
	if (s >= n_slices) return;
	unsigned long long *slice = table + (uint64_t)s * slice_slots;
	const uint64_t base = (uint64_t)blockIdx.x * per_cta, mask = slice_slots - 1;
	unsigned long long acc = 0;
	for (uint64_t i = threadIdx.x; i < per_cta; i += blockDim.x) {
		const uint64_t pos = mix(base + i) >> 20 & mask;
		acc += slice[pos];
		atomicAdd(slice + pos, 1ull); /* result unused: RED */
	}
	if (acc == 12345) *sink = acc;
}

// tags and counts in two arrays: a k-mer that has its entry costs one load of the tag and one 32-bit reduction on the
// count (no return value, nothing to wait for); a new one a compare-and-swap on the tag as well.  `dup` of every 8 keys
// repeat an earlier key of the thread's slice (entry exists), the rest are new.
__global__ void l2split(unsigned long long *tags, uint32_t *counts, uint64_t slice_slots, uint32_t n_slices, uint32_t ctas_per_slice, uint64_t per_cta,
                        uint32_t dup)
{
	const uint32_t s = blockIdx.x / ctas_per_slice;
	if (s >= n_slices) return;
	unsigned long long *slice = tags + (uint64_t)s * slice_slots;
	uint32_t *cnt = counts + (uint64_t)s * slice_slots;
	const uint64_t base = (uint64_t)blockIdx.x * per_cta, mask = slice_slots - 1;
	for (uint64_t i = threadIdx.x; i < per_cta; i += blockDim.x) {
		const uint64_t n = base + i;
		const uint64_t key = (n & 7) < dup ? mix(base + (i & 1023)) : mix(n); /* a repeat of one of the CTA's first keys, or a new key */
		const uint64_t tag = key >> 24 | 1;
		uint64_t pos = (tag * 0x9E3779B97F4A7C15ull) >> 20 & mask;
		for (int t = 0; t < 64; ++t, pos = (pos + 1) & mask) {
			unsigned long long v = slice[pos];
			if (v == 0) v = atomicCAS(slice + pos, 0ull, tag), v = v ? v : tag;
			if (v == tag) {
				atomicAdd(cnt + pos, 1u); /* RED */
				break;
			}
		}
	}
}

// the same key stream into the one-word slots the library uses (tag << 10 | count, compare-and-swap loop)
__global__ void l2word(unsigned long long *table, uint64_t slice_slots, uint32_t n_slices, uint32_t ctas_per_slice, uint64_t per_cta, uint32_t dup)
{
	const uint32_t s = blockIdx.x / ctas_per_slice;
	if (s >= n_slices) return;
	unsigned long long *slice = table + (uint64_t)s * slice_slots;
	const uint64_t base = (uint64_t)blockIdx.x * per_cta;
	for (uint64_t i = threadIdx.x; i < per_cta; i += blockDim.x) {
		const uint64_t n = base + i;
		const uint64_t key = (n & 7) < dup ? mix(base + (i & 1023)) : mix(n);
		insert(slice, slice_slots - 1, key >> 24 | 1);
	}
}

// T threads, IPT keys per thread; lists[r * cap + ...], cursors 256 bytes apart as in the library
template <int RB, int T, int IPT>
__global__ void __launch_bounds__(T, 1) part(const uint64_t *__restrict__ in, uint64_t n, uint64_t *lists, unsigned long long *cur, uint64_t cap, int shift)
{
	constexpr int R = 1 << RB, TILE = T * IPT;
	extern __shared__ unsigned long long smem[];
	unsigned long long *stage = smem;
	uint32_t *cnt = reinterpret_cast<uint32_t *>(stage + TILE); /* R counts, then R local bases */
	uint32_t *lbase = cnt + R;
	unsigned long long *gbase = reinterpret_cast<unsigned long long *>(lbase + R);
	const int tid = threadIdx.x;
	const uint64_t n_tiles = n / TILE;
	for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		for (int r = tid; r < R; r += T) cnt[r] = 0;
		__syncthreads();
		uint64_t key[IPT];
		uint16_t rank[IPT];
		const uint64_t *src = in + tile * TILE;
#pragma unroll
		for (int j = 0; j < IPT; ++j) key[j] = __ldcs(src + j * T + tid);
#pragma unroll
		for (int j = 0; j < IPT; ++j) rank[j] = (uint16_t)atomicAdd(cnt + (key[j] >> shift & (R - 1)), 1u);
		__syncthreads();
		/* exclusive prefix over the R counts: PER per thread, warp scan, one scan over the warp totals */
		{
			constexpr int PER = R / T > 0 ? R / T : 1;
			__shared__ uint32_t wsum[32];
			uint32_t sum = 0, c[PER];
#pragma unroll
			for (int i = 0; i < PER; ++i) c[i] = tid * PER + i < R ? cnt[tid * PER + i] : 0, sum += c[i];
			uint32_t incl = sum;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
				if ((tid & 31) >= d) incl += v;
			}
			if ((tid & 31) == 31) wsum[tid >> 5] = incl;
			__syncthreads();
			if (tid < 32) {
				const uint32_t w = tid < T / 32 ? wsum[tid] : 0;
				uint32_t wi = w;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
					if (tid >= d) wi += v;
				}
				wsum[tid] = wi - w;
			}
			__syncthreads();
			uint32_t at = wsum[tid >> 5] + incl - sum;
#pragma unroll
			for (int i = 0; i < PER; ++i)
				if (tid * PER + i < R) lbase[tid * PER + i] = at, at += c[i];
		}
		__syncthreads();
		for (int r = tid; r < R; r += T) {
			const uint32_t c = cnt[r];
			gbase[r] = (c ? atomicAdd(cur + r * 32, (unsigned long long)c) : 0ull) + (unsigned long long)r * cap - lbase[r];
		}
#pragma unroll
		for (int j = 0; j < IPT; ++j) stage[lbase[key[j] >> shift & (R - 1)] + rank[j]] = key[j];
		__syncthreads();
#pragma unroll 4
		for (int i = tid; i < TILE; i += T) {
			const unsigned long long k = stage[i];
			lists[(gbase[k >> shift & (R - 1)] + i) % (cap << RB)] = k; /* the modulo only keeps a repeated run of the benchmark in bounds */
		}
		__syncthreads();
	}
}

__global__ void fill_keys(uint64_t *keys, uint64_t n)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) keys[i] = mix(i);
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
	for (uint32_t mib : {16u, 64u}) {
		const uint64_t slice_slots = (uint64_t)mib << 17;
		const uint32_t n_slices = 64, ctas_per_slice = 1024;
		const uint64_t per_cta = slice_slots / 4 / ctas_per_slice;
		unsigned long long *table;
		cudaMalloc(&table, slice_slots * n_slices * 8);
		cudaMemset(table, 0, slice_slots * n_slices * 8);
		const uint64_t total = per_cta * ctas_per_slice * n_slices;
		float ms = timed([&] { l2red<<<n_slices * ctas_per_slice, block>>>(table, slice_slots, n_slices, ctas_per_slice, per_cta, sink); });
		printf("l2red      slice %3u MiB  %7.2f ms  %6.1f G updates/s\n", mib, ms, total / ms / 1e6);
		cudaFree(table);
	}
	for (uint32_t dup : {0u, 5u, 7u}) { /* 0, 5 or 7 of 8 keys have their entry already (config 5: 5 of 8) */
		const uint32_t mib = 64;
		const uint64_t slice_slots = (uint64_t)mib << 17;
		const uint32_t n_slices = 32, ctas_per_slice = 1024;
		const uint64_t per_cta = slice_slots / 4 / ctas_per_slice * 4; /* as many keys as a quarter-full slice of new keys would take, times 4 */
		unsigned long long *table;
		uint32_t *counts;
		cudaMalloc(&table, slice_slots * n_slices * 8);
		cudaMalloc(&counts, slice_slots * n_slices * 4);
		const uint64_t total = per_cta * ctas_per_slice * n_slices;
		cudaMemset(table, 0, slice_slots * n_slices * 8);
		cudaDeviceSynchronize();
		cudaEvent_t a, b;
		cudaEventCreate(&a), cudaEventCreate(&b);
		float ms = 0;
		cudaEventRecord(a);
		l2word<<<n_slices * ctas_per_slice, block>>>(table, slice_slots, n_slices, ctas_per_slice, per_cta, dup);
		cudaEventRecord(b);
		cudaEventSynchronize(b);
		cudaEventElapsedTime(&ms, a, b);
		printf("l2word     %u of 8 keys present, slice %u MiB  %7.2f ms  %6.1f G inserts/s\n", dup, mib, ms, total / ms / 1e6);
		cudaMemset(table, 0, slice_slots * n_slices * 8);
		cudaMemset(counts, 0, slice_slots * n_slices * 4);
		cudaDeviceSynchronize();
		cudaEventRecord(a);
		l2split<<<n_slices * ctas_per_slice, block>>>(table, counts, slice_slots, n_slices, ctas_per_slice, per_cta, dup);
		cudaEventRecord(b);
		cudaEventSynchronize(b);
		cudaEventElapsedTime(&ms, a, b);
		printf("l2split    %u of 8 keys present, slice %u MiB  %7.2f ms  %6.1f G inserts/s  (tags + 32-bit counts apart)\n", dup, mib, ms, total / ms / 1e6);
		cudaFree(table), cudaFree(counts);
	}
	{
		uint64_t *keys, *lists;
		unsigned long long *cur;
		cudaMalloc(&keys, n * 8);
		fill_keys<<<grid, block>>>(keys, n);
		auto run = [&](auto kernel, int rb, int threads, int ipt, const char *name) {
			const uint64_t cap = ((n >> rb) * 2 + 31) & ~31ull;
			cudaMalloc(&lists, (cap << rb) * 8);
			cudaMalloc(&cur, 256ull << rb);
			cudaMemset(cur, 0, 256ull << rb);
			const size_t smem = (size_t)threads * ipt * 8 + ((size_t)8 << rb) + ((size_t)8 << rb);
			cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			float ms = timed([&] { kernel<<<148, threads, smem>>>(keys, n, lists, cur, cap, 20); });
			printf("part       %-22s smem %3zu KB  %7.2f ms  %6.1f G keys/s  (%.0f GB/s read + written)\n", name, smem >> 10, ms, n / ms / 1e6, n * 16 / ms / 1e6);
			cudaFree(lists), cudaFree(cur);
		};
		run(part<6, 512, 16>, 6, 512, 16, "R=64 512x16");
		run(part<8, 512, 16>, 8, 512, 16, "R=256 512x16");
		run(part<8, 512, 32>, 8, 512, 32, "R=256 512x32");
		run(part<10, 512, 32>, 10, 512, 32, "R=1024 512x32");
		run(part<10, 1024, 16>, 10, 1024, 16, "R=1024 1024x16");
		run(part<12, 512, 40>, 12, 512, 40, "R=4096 512x40");
		cudaFree(keys);
	}
	{
		cudaFuncSetAttribute(smemcas, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
		const uint64_t per_cta = n / 148;
		float ms = timed([&] { smemcas<<<148, 1024, 16384 * 8>>>(per_cta, sink); });
		printf("smemcas    16 K slots per CTA  %7.2f ms  %6.1f G inserts/s\n", ms, per_cta * 148 / ms / 1e6);
	}
	return 0;
}
