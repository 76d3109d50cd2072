/*
 * vafgpu_api.cu -- the C ABI of include/vafgpu.h: contexts, pinned staging blocks, one
 * stream per block for copy/compute overlap, round-robin over the devices.  With several
 * devices every kernel adds its hits straight into ONE counter vector (device 0's, or one
 * attached from another process) over NVLink peer memory, so there is no merge step and no
 * collective.  Host logic only; the kernels are in vafgpu_kernels.cu.
 */
#include "../../include/vafgpu.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <string>
#include <vector>

#include "vafgpu_kernels.cuh"
#include "vafgpu_tables.hpp"

using namespace vafgpu;

namespace {

thread_local std::string g_create_error;

struct Block {
	char *h = nullptr;        /* pinned host staging */
	uint8_t *d = nullptr;     /* device copy         */
	cudaStream_t stream = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr; /* copy start, copy end, kernel end */
	bool in_flight = false;
	bool owned = false; /* a producer is filling it */
	size_t used = 0;
};

struct Device {
	int ordinal = 0;
	int n_sm = 0;
	uint64_t keep_policy = 0;
	uint32_t *d_filter = nullptr;
	uint32_t *d_buckets = nullptr;
	uint32_t *d_filter2 = nullptr;
	vg_slot_t *d_slots = nullptr;
	uint64_t *d_rkeys = nullptr;
	uint32_t *d_rvals = nullptr;
	uint32_t *d_counts = nullptr;  /* this device's own counter vector                          */
	uint32_t *counts_to = nullptr; /* where its kernels add: d_counts, device 0's vector mapped
	                                  as peer memory, or a vector attached from another process */
	void *attached = nullptr;      /* the mapping cudaIpcOpenMemHandle made on this device       */
	unsigned long long *d_stats = nullptr;
	cudaStream_t main_stream = nullptr;
	std::vector<Block> blocks;
	char *h_slab = nullptr;    /* the blocks' page-locked host memory and their device copies: one allocation */
	uint8_t *d_slab = nullptr; /* each (a driver call per block is start-up time of every run)                */
};

} // namespace

struct vafgpu_ctx {
	int k = 0;
	unsigned flags = 0;
	uint32_t n_patterns = 0;
	size_t n_counts = 0; /* 2 * n_patterns, at least 2 */
	size_t block_bytes = 0;
	Plan plan;
	uint32_t filter_words = 0, filter2_words = 0, n_buckets = 0, rbits = 0;
	bool canon = false, defer = false;
	std::vector<Device> devs;
	bool host_merge = false; /* counters stay per device and are summed on the host */
	bool attached = false;   /* the counters live in another context (vafgpu_attach_counters) */
	std::mutex mu;    /* guards seq, block ownership, st and err: several producers may run */
	std::condition_variable block_free; /* a producer handed a staging block back */
	uint64_t seq = 0; /* blocks handed out so far: round-robin over devices, then buffers */
	vafgpu_stats st{};
	std::string err;
	vafgpu_producer *def = nullptr; /* the producer behind vafgpu_add_read / vafgpu_submit_stream */
};

/* One stream of reads being packed into staging blocks; one per reader thread. */
struct vafgpu_producer {
	vafgpu_ctx *c = nullptr;
	Block *cur = nullptr;
	int cur_dev = 0;
	uint64_t n_reads = 0, n_bases = 0; /* added to the context's statistics at flush */
	std::vector<char> scratch;
};

namespace {

int fail(vafgpu_ctx *c, int code, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (c) {
		std::lock_guard<std::mutex> lk(c->mu);
		c->err = buf;
	} else g_create_error = buf;
	return code;
}

#define CU(c, call)                                                                              \
	do {                                                                                         \
		cudaError_t e_ = (call);                                                                 \
		if (e_ != cudaSuccess)                                                                   \
			return fail(c, VAFGPU_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

int wait_block(vafgpu_ctx *c, Block &b)
{
	if (!b.in_flight) return VAFGPU_OK;
	CU(c, cudaEventSynchronize(b.e2));
	float h2d = 0, ker = 0;
	cudaEventElapsedTime(&h2d, b.e0, b.e1);
	cudaEventElapsedTime(&ker, b.e1, b.e2);
	std::lock_guard<std::mutex> lk(c->mu);
	c->st.h2d_ms += h2d;
	c->st.kernel_ms += ker;
	b.in_flight = false;
	return VAFGPU_OK;
}

/* next free staging block, round-robin over devices then buffers; blocks another producer
 * is filling are skipped, a block still in flight is waited for (that is the back-pressure
 * which replaces kt_pipeline's "at most three blocks") */
int acquire(vafgpu_producer *p)
{
	vafgpu_ctx *c = p->c;
	const size_t nd = c->devs.size();
	Block *b = nullptr;
	int di = 0;
	{
		std::unique_lock<std::mutex> lk(c->mu);
		const size_t total = nd * c->devs[0].blocks.size();
		for (;;) {
			for (size_t tries = 0; tries < total && !b; ++tries) {
				di = (int)(c->seq % nd);
				Device &d = c->devs[di];
				Block &cand = d.blocks[(c->seq / nd) % d.blocks.size()];
				++c->seq;
				if (!cand.owned) {
					cand.owned = true;
					b = &cand;
				}
			}
			if (b) break;
			c->block_free.wait(lk); /* more producers than blocks: wait for one to be submitted */
		}
	}
	int rc = wait_block(c, *b);
	if (rc) return rc;
	b->used = 0;
	p->cur = b;
	p->cur_dev = di;
	return VAFGPU_OK;
}

ScanArgs scan_args(const vafgpu_ctx *c, const Device &d, const uint8_t *bytes, size_t n, uint32_t *counts)
{
	ScanArgs a{};
	a.bytes = bytes;
	a.n_bytes = n;
	a.counts = counts ? counts : d.counts_to;
	a.stats = d.d_stats;
	a.k = c->k;
	a.stride = c->plan.stride;
	a.len = c->plan.len;
	a.canon = c->canon ? 1 : 0;
	a.filter = d.d_filter;
	a.filter_words = c->filter_words;
	a.defer = c->defer ? 1 : 0;
	a.filter2 = d.d_filter2;
	a.filter2_words = c->filter2_words;
	a.buckets = d.d_buckets;
	a.n_buckets = c->n_buckets;
	a.slots = d.d_slots;
	a.keep_policy = d.keep_policy;
	a.rkeys = d.d_rkeys;
	a.rvals = d.d_rvals;
	a.rbits = c->rbits;
	return a;
}

cudaError_t launch(const vafgpu_ctx *c, const Device &d, const ScanArgs &a, cudaStream_t s)
{
	return (c->flags & VAFGPU_F_REFERENCE_RECIPE) ? launch_recipe_scan(a, d.n_sm, s)
	                                              : launch_anchor_scan(a, d.n_sm, s);
}

/* copy the current block to its device and scan it there, all on the block's stream */
int submit_current(vafgpu_producer *p)
{
	vafgpu_ctx *c = p->c;
	Block *b = p->cur;
	if (!b) return VAFGPU_OK;
	p->cur = nullptr;
	size_t n = (b->used + 15) & ~(size_t)15;
	if (b->used) {
		Device &d = c->devs[p->cur_dev];
		memset(b->h + b->used, '\n', n - b->used);
		CU(c, cudaSetDevice(d.ordinal));
		CU(c, cudaEventRecord(b->e0, b->stream));
		CU(c, cudaMemcpyAsync(b->d, b->h, n, cudaMemcpyHostToDevice, b->stream));
		CU(c, cudaEventRecord(b->e1, b->stream));
		CU(c, launch(c, d, scan_args(c, d, b->d, n, nullptr), b->stream));
		CU(c, cudaEventRecord(b->e2, b->stream));
	}
	{
		std::lock_guard<std::mutex> lk(c->mu);
		if (b->used) {
			b->in_flight = true;
			c->st.n_blocks++;
			c->st.n_bytes += n;
		}
		b->owned = false;
	}
	c->block_free.notify_one();
	return VAFGPU_OK;
}

/* room for `need` more bytes in the current block, submitting / acquiring as required */
int ensure_room(vafgpu_producer *p, size_t need)
{
	if (p->cur && p->cur->used + need > p->c->block_bytes) {
		int rc = submit_current(p);
		if (rc) return rc;
	}
	if (!p->cur) return acquire(p);
	return VAFGPU_OK;
}

void destroy_device(Device &d)
{
	cudaSetDevice(d.ordinal);
	for (Block &b : d.blocks) {
		if (b.stream) cudaStreamSynchronize(b.stream);
		if (b.e0) cudaEventDestroy(b.e0);
		if (b.e1) cudaEventDestroy(b.e1);
		if (b.e2) cudaEventDestroy(b.e2);
		if (b.stream) cudaStreamDestroy(b.stream);
	}
	if (d.h_slab) cudaFreeHost(d.h_slab);
	if (d.d_slab) cudaFree(d.d_slab);
	if (d.main_stream) cudaStreamDestroy(d.main_stream);
	if (d.attached) cudaIpcCloseMemHandle(d.attached);
	cudaFree(d.d_filter);
	cudaFree(d.d_buckets);
	cudaFree(d.d_filter2);
	cudaFree(d.d_slots);
	cudaFree(d.d_rkeys);
	cudaFree(d.d_rvals);
	cudaFree(d.d_counts);
	cudaFree(d.d_stats);
}

} // namespace

extern "C" {

const char *vafgpu_version(void) { return "vafgpu 0.2 (sm_100a; anchor kernel v17)"; }

int vafgpu_plan(int k, int *stride, int *len)
{
	if (k < 1 || k > 31) return VAFGPU_EINVAL;
	Plan p = make_plan(k);
	if (stride) *stride = p.stride;
	if (len) *len = p.len;
	return VAFGPU_OK;
}

const char *vafgpu_strerror(const vafgpu_ctx *ctx)
{
	return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int vafgpu_create(vafgpu_ctx **out, int k, const uint64_t *keys, const uint32_t *vals, uint32_t n_entries,
                  uint32_t n_patterns, size_t block_bytes, int n_buffers, int n_devices, unsigned flags)
{
	if (!out) return fail(nullptr, VAFGPU_EINVAL, "ctx is NULL");
	*out = nullptr;
	if (k < 1 || k > 31) return fail(nullptr, VAFGPU_EINVAL, "k = %d is outside 1..31", k);
	if (n_entries && (!keys || !vals)) return fail(nullptr, VAFGPU_EINVAL, "keys/vals are NULL");
	if (n_patterns > 0x3FFFFFFFu) return fail(nullptr, VAFGPU_EINVAL, "too many patterns (%u)", n_patterns);
	if ((uint64_t)n_entries * 2 * (uint64_t)make_plan(k).stride >= 0x7FFFFFFFull) /* payload indices are 31 bits */
		return fail(nullptr, VAFGPU_EINVAL, "too many k-mers (%u) for the exact table", n_entries);
	const uint64_t kmask = (1ULL << 2 * k) - 1;
	for (uint32_t i = 0; i < n_entries; ++i) {
		if (keys[i] > kmask) return fail(nullptr, VAFGPU_EINVAL, "key %u does not fit 2k bits", i);
		if (vals[i] >= 2ull * n_patterns) return fail(nullptr, VAFGPU_EINVAL, "value %u names pattern %u of %u", i, vals[i] >> 1, n_patterns);
	}
	const bool timing = getenv("VAFGPU_TIMING") != nullptr; /* start-up breakdown on stderr */
	auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double tt = tnow();
	auto lap = [&](const char *what) {
		if (timing) {
			double t = tnow();
			fprintf(stderr, "[vafgpu] %-28s %8.1f ms\n", what, (t - tt) * 1e3);
			tt = t;
		}
	};
	int visible = 0;
	cudaError_t ce = cudaGetDeviceCount(&visible); /* the first runtime call: cuInit */
	lap("CUDA init (device count)");
	if (ce != cudaSuccess || visible < 1)
		return fail(nullptr, VAFGPU_ENOGPU, "no CUDA device: %s", ce == cudaSuccess ? "count is 0" : cudaGetErrorString(ce));
	if (n_devices <= 0 || n_devices > visible) n_devices = visible;
	if (block_bytes == 0) block_bytes = (size_t)16 << 20;
	if (block_bytes < 4096) block_bytes = 4096;
	if (block_bytes > ((size_t)1 << 34)) return fail(nullptr, VAFGPU_EINVAL, "block_bytes too large");
	if (n_buffers <= 0) n_buffers = 3;

	vafgpu_ctx *c = new (std::nothrow) vafgpu_ctx;
	if (!c) return fail(nullptr, VAFGPU_ENOMEM, "out of memory");
	c->k = k;
	c->flags = flags;
	c->n_patterns = n_patterns;
	c->n_counts = (size_t)2 * (n_patterns ? n_patterns : 1);
	c->block_bytes = block_bytes;

	RecipeTable rt;
	AnchorTables at;
	build_recipe_table(k, keys, vals, n_entries, n_patterns, rt);
	build_anchor_tables(k, keys, vals, n_entries, at);
	lap("host tables");
	c->plan = at.plan;
	c->filter_words = at.filter_words;
	c->filter2_words = (uint32_t)at.filter2.size();
	c->n_buckets = at.n_buckets;
	c->canon = at.canon;
	c->defer = at.defer;
	c->rbits = rt.bits;

	int rc = VAFGPU_OK;
	c->devs.resize(n_devices);
	/* every device is set up by a thread of its own: creating a context takes 0.2-2 s, and eight of them one
	 * after the other were most of the start-up of a run on eight GPUs (the breakdown is printed for device 0) */
	auto setup = [&](int i) -> int {
		Device &d = c->devs[i];
		d.ordinal = i;
		double tt = tnow(); /* shadows the caller's: laps of this device */
		auto lap = [&](const char *what) {
			if (timing && i == 0) {
				double t = tnow();
				fprintf(stderr, "[vafgpu] %-28s %8.1f ms\n", what, (t - tt) * 1e3);
				tt = t;
			}
		};
		return [&]() -> int {
			cudaDeviceProp prop;
			CU(c, cudaSetDevice(i));
			CU(c, cudaGetDeviceProperties(&prop, i));
			if (prop.major != 10)
				return fail(c, VAFGPU_ENOGPU, "device %d (%s) is sm_%d%d; this library carries sm_100a code only", i, prop.name, prop.major, prop.minor);
			d.n_sm = prop.multiProcessorCount;
			lap("context");
			CU(c, kernels_make_policy(&d.keep_policy));
			lap("module load + policy kernel");
			if (!(flags & VAFGPU_F_REFERENCE_RECIPE)) {
				ScanArgs form{};
				form.stride = c->plan.stride, form.len = c->plan.len, form.canon = c->canon, form.defer = c->defer;
				CU(c, kernels_prepare(form)); /* dynamic shared-memory opt-in of the instantiation this panel selects */
			}
			CU(c, cudaStreamCreateWithFlags(&d.main_stream, cudaStreamNonBlocking));
			CU(c, cudaMalloc(&d.d_filter, at.filter.size() * 4));
			CU(c, cudaMalloc(&d.d_filter2, at.filter2.size() * 4));
			CU(c, cudaMalloc(&d.d_buckets, at.buckets.size() * 4));
			CU(c, cudaMalloc(&d.d_slots, at.slots.size() * sizeof(vg_slot_t)));
			CU(c, cudaMalloc(&d.d_rkeys, rt.keys.size() * 8));
			CU(c, cudaMalloc(&d.d_rvals, rt.vals.size() * 4));
			CU(c, cudaMalloc(&d.d_counts, c->n_counts * 4));
			CU(c, cudaMalloc(&d.d_stats, ST_N * sizeof(unsigned long long)));
			CU(c, cudaMemcpy(d.d_filter, at.filter.data(), at.filter.size() * 4, cudaMemcpyHostToDevice));
			CU(c, cudaMemcpy(d.d_filter2, at.filter2.data(), at.filter2.size() * 4, cudaMemcpyHostToDevice));
			CU(c, cudaMemcpy(d.d_buckets, at.buckets.data(), at.buckets.size() * 4, cudaMemcpyHostToDevice));
			CU(c, cudaMemcpy(d.d_slots, at.slots.data(), at.slots.size() * sizeof(vg_slot_t), cudaMemcpyHostToDevice));
			CU(c, cudaMemcpy(d.d_rkeys, rt.keys.data(), rt.keys.size() * 8, cudaMemcpyHostToDevice));
			CU(c, cudaMemcpy(d.d_rvals, rt.vals.data(), rt.vals.size() * 4, cudaMemcpyHostToDevice));
			CU(c, cudaMemset(d.d_counts, 0, c->n_counts * 4));
			CU(c, cudaMemset(d.d_stats, 0, ST_N * sizeof(unsigned long long)));
			d.counts_to = d.d_counts;
			lap("tables to device");
			d.blocks.resize(n_buffers);
			const size_t pitch = (block_bytes + 64 + 255) & ~(size_t)255;
			CU(c, cudaHostAlloc(&d.h_slab, pitch * (size_t)n_buffers, cudaHostAllocPortable));
			CU(c, cudaMalloc(&d.d_slab, pitch * (size_t)n_buffers));
			size_t at = 0;
			for (Block &b : d.blocks) {
				b.h = d.h_slab + at;
				b.d = d.d_slab + at;
				at += pitch;
				CU(c, cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
				CU(c, cudaEventCreate(&b.e0));
				CU(c, cudaEventCreate(&b.e1));
				CU(c, cudaEventCreate(&b.e2));
			}
			lap("staging blocks");
			return VAFGPU_OK;
		}();
	};
	if (n_devices == 1) {
		rc = setup(0);
	} else {
		std::vector<int> rcs((size_t)n_devices, VAFGPU_OK);
		std::vector<std::thread> th;
		for (int i = 0; i < n_devices; ++i) th.emplace_back([&, i] { rcs[(size_t)i] = setup(i); });
		for (std::thread &t : th) t.join();
		for (int i = 0; i < n_devices && rc == VAFGPU_OK; ++i) rc = rcs[(size_t)i];
	}
	tt = tnow();
	/* Several devices: every kernel adds into device 0's counter vector through NVLink peer
	 * memory (hits are rare -- tens of thousands per 10^10 bases -- and RED.ADD needs no answer),
	 * so the result is final when the streams have drained.  Without peer access between all
	 * devices and device 0 the counters stay per device and are summed through host memory. */
	c->host_merge = (flags & VAFGPU_F_HOST_MERGE) != 0;
	if (rc == VAFGPU_OK && n_devices > 1 && !c->host_merge) {
		for (int i = 1; i < n_devices && !c->host_merge; ++i) {
			int can = 0, atomics = 0;
			if (cudaDeviceCanAccessPeer(&can, i, 0) != cudaSuccess || !can) c->host_merge = true;
			else if (cudaDeviceGetP2PAttribute(&atomics, cudaDevP2PAttrNativeAtomicSupported, i, 0) != cudaSuccess || !atomics)
				c->host_merge = true;
		}
		for (int i = 1; i < n_devices && !c->host_merge; ++i) {
			cudaSetDevice(i);
			cudaError_t pe = cudaDeviceEnablePeerAccess(0, 0);
			if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
			else if (pe != cudaSuccess) c->host_merge = true;
		}
		cudaGetLastError();
		if (!c->host_merge)
			for (int i = 1; i < n_devices; ++i) c->devs[i].counts_to = c->devs[0].d_counts;
		lap("peer access");
	}
	if (rc != VAFGPU_OK) {
		g_create_error = c->err;
		vafgpu_destroy(c);
		return rc;
	}
	c->st.n_devices = n_devices;
	c->st.anchor_stride = c->plan.stride;
	c->st.anchor_len = c->plan.len;
	c->st.filter_bytes = c->filter_words * 4;
	c->st.table_slots = 3u * c->n_buckets;
	if (!(flags & VAFGPU_F_REFERENCE_RECIPE)) {
		ScanArgs form{};
		form.stride = c->plan.stride, form.len = c->plan.len, form.canon = c->canon, form.defer = c->defer;
		c->st.filter_canon = c->canon, c->st.lookup_deferred = c->defer;
		c->st.kernel_threads = kernels_threads(form);
		c->st.filter2_bytes = c->defer ? c->filter2_words * 4 : 0;
	}
	*out = c;
	return VAFGPU_OK;
}

/* the producer behind the single-producer entry points, made on first use */
static vafgpu_producer *default_producer(vafgpu_ctx *c)
{
	if (!c->def) {
		c->def = new (std::nothrow) vafgpu_producer;
		if (c->def) c->def->c = c;
	}
	return c->def;
}

int vafgpu_producer_create(vafgpu_ctx *c, vafgpu_producer **out)
{
	if (!c || !out) return VAFGPU_EINVAL;
	vafgpu_producer *p = new (std::nothrow) vafgpu_producer;
	if (!p) return fail(c, VAFGPU_ENOMEM, "out of memory");
	p->c = c;
	*out = p;
	return VAFGPU_OK;
}

int vafgpu_producer_add_read(vafgpu_producer *p, const char *seq, size_t len)
{
	if (!p || (!seq && len)) return VAFGPU_EINVAL;
	vafgpu_ctx *c = p->c;
	if (len < (size_t)c->k) return VAFGPU_OK; /* vaf-counter.c:494 */
	p->n_reads++;
	p->n_bases += len;
	const bool simd = !(c->flags & VAFGPU_F_STRICT_BYTES); /* the Makefile builds the reference with -mssse3 (Makefile:44) */
	if (len + 1 <= c->block_bytes) {
		int rc = ensure_room(p, len + 1);
		if (rc) return rc;
		Block *b = p->cur;
		vafgpu_canonicalise_read(seq, len, b->h + b->used, simd);
		b->h[b->used + len] = '\n';
		b->used += len + 1;
		return VAFGPU_OK;
	}
	/* a read longer than a block (a chromosome): canonicalise it whole, since the byte rule
	 * depends on the offset within the read, then cut it into pieces that overlap by k-1
	 * bases so that every k-mer lies in exactly one piece */
	if (p->scratch.size() < len) p->scratch.resize(len);
	vafgpu_canonicalise_read(seq, len, p->scratch.data(), simd);
	const size_t piece = c->block_bytes - 1, step = piece - (size_t)(c->k - 1);
	for (size_t at = 0;; at += step) {
		size_t n = len - at < piece ? len - at : piece;
		int rc = ensure_room(p, n + 1);
		if (rc) return rc;
		Block *b = p->cur;
		memcpy(b->h + b->used, p->scratch.data() + at, n);
		b->h[b->used + n] = '\n';
		b->used += n + 1;
		if (at + n >= len) break;
	}
	return VAFGPU_OK;
}

int vafgpu_producer_flush(vafgpu_producer *p)
{
	if (!p) return VAFGPU_EINVAL;
	int rc = submit_current(p);
	std::lock_guard<std::mutex> lk(p->c->mu);
	p->c->st.n_reads += p->n_reads;
	p->c->st.n_bases += p->n_bases;
	p->n_reads = p->n_bases = 0;
	return rc;
}

int vafgpu_producer_destroy(vafgpu_producer *p)
{
	if (!p) return VAFGPU_OK;
	int rc = vafgpu_producer_flush(p);
	if (p->c->def == p) p->c->def = nullptr;
	delete p;
	return rc;
}

int vafgpu_add_read(vafgpu_ctx *c, const char *seq, size_t len)
{
	if (!c) return VAFGPU_EINVAL;
	vafgpu_producer *p = default_producer(c);
	if (!p) return fail(c, VAFGPU_ENOMEM, "out of memory");
	return vafgpu_producer_add_read(p, seq, len);
}

int vafgpu_submit_stream(vafgpu_ctx *c, const char *bytes, size_t n_bytes, uint64_t n_reads, uint64_t n_bases)
{
	if (!c || (!bytes && n_bytes)) return VAFGPU_EINVAL;
	vafgpu_producer *p = default_producer(c);
	if (!p) return fail(c, VAFGPU_ENOMEM, "out of memory");
	p->n_reads += n_reads;
	p->n_bases += n_bases;
	/* page-locked caller memory is copied to the device as it is; pageable memory goes through
	 * the pinned staging blocks */
	cudaPointerAttributes attr;
	bool pinned = cudaPointerGetAttributes(&attr, bytes) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError(); /* a plain malloc pointer is reported as an error by older drivers */
	size_t at = 0;
	while (at < n_bytes) {
		/* close whatever add_read left open, then fill whole blocks straight from the caller */
		int rc = submit_current(p);
		if (rc) return rc;
		rc = acquire(p);
		if (rc) return rc;
		Block *b = p->cur;
		size_t n = n_bytes - at, advance;
		bool add_nl = false;
		if (n > c->block_bytes) {
			/* cut after the last separator that fits; a single read longer than a block is cut
			 * with a k-1 overlap like in vafgpu_add_read */
			n = c->block_bytes;
			const char *nl = (const char *)memrchr(bytes + at, '\n', n);
			if (nl) {
				n = (size_t)(nl - (bytes + at)) + 1;
				advance = n;
			} else {
				n -= 1;
				add_nl = true;
				advance = n - (size_t)(c->k - 1);
			}
		} else {
			advance = n;
			add_nl = bytes[at + n - 1] != '\n';
		}
		if (!pinned) {
			memcpy(b->h, bytes + at, n);
			if (add_nl) b->h[n++] = '\n';
			b->used = n;
			at += advance;
			continue; /* submitted at the top of the loop or after it */
		}
		/* zero-copy path: H2D from the caller's buffer, separator and padding written on the device */
		Device &d = c->devs[p->cur_dev];
		p->cur = nullptr;
		const size_t n16 = (n + (add_nl ? 1 : 0) + 15) & ~(size_t)15;
		CU(c, cudaSetDevice(d.ordinal));
		CU(c, cudaEventRecord(b->e0, b->stream));
		CU(c, cudaMemcpyAsync(b->d, bytes + at, n, cudaMemcpyHostToDevice, b->stream));
		if (n16 > n) CU(c, cudaMemsetAsync(b->d + n, '\n', n16 - n, b->stream));
		CU(c, cudaEventRecord(b->e1, b->stream));
		CU(c, launch(c, d, scan_args(c, d, b->d, n16, nullptr), b->stream));
		CU(c, cudaEventRecord(b->e2, b->stream));
		{
			std::lock_guard<std::mutex> lk(c->mu);
			b->in_flight = true;
			b->owned = false;
			c->st.n_blocks++;
			c->st.n_bytes += n16;
		}
		c->block_free.notify_one();
		at += advance;
	}
	return vafgpu_producer_flush(p);
}

int vafgpu_count_device(vafgpu_ctx *c, int device, const void *d_bytes, size_t n_bytes, uint32_t *d_counts, void *stream)
{
	if (!c || device < 0 || device >= (int)c->devs.size()) return VAFGPU_EINVAL;
	if (((uintptr_t)d_bytes & 15) || (n_bytes & 15)) return fail(c, VAFGPU_EINVAL, "device stream must be 16-byte aligned and a multiple of 16 bytes");
	Device &d = c->devs[device];
	CU(c, cudaSetDevice(d.ordinal));
	cudaStream_t s = stream ? (cudaStream_t)stream : d.main_stream;
	CU(c, launch(c, d, scan_args(c, d, (const uint8_t *)d_bytes, n_bytes, d_counts), s));
	std::lock_guard<std::mutex> lk(c->mu);
	c->st.n_blocks++;
	c->st.n_bytes += n_bytes;
	return VAFGPU_OK;
}

int vafgpu_finish(vafgpu_ctx *c, uint32_t *counts, vafgpu_stats *stats)
{
	if (!c) return VAFGPU_EINVAL;
	int rc = c->def ? vafgpu_producer_flush(c->def) : VAFGPU_OK;
	if (rc) return rc;
	for (Device &d : c->devs) {
		CU(c, cudaSetDevice(d.ordinal));
		for (Block &b : d.blocks) {
			rc = wait_block(c, b);
			if (rc) return rc;
		}
		CU(c, cudaStreamSynchronize(d.main_stream));
	}
	std::vector<uint32_t> total(c->n_counts, 0);
	if (c->attached) {
		/* the counters live in the context this one is attached to; that one reports them */
	} else if (!c->host_merge) {
		/* one vector: every device's kernels have added into it (peer memory); nothing to merge */
		CU(c, cudaSetDevice(c->devs[0].ordinal));
		CU(c, cudaMemcpy(total.data(), c->devs[0].d_counts, c->n_counts * 4, cudaMemcpyDeviceToHost));
	} else {
		std::vector<uint32_t> part(c->n_counts);
		for (Device &d : c->devs) {
			CU(c, cudaSetDevice(d.ordinal));
			CU(c, cudaMemcpy(part.data(), d.d_counts, c->n_counts * 4, cudaMemcpyDeviceToHost));
			for (size_t j = 0; j < c->n_counts; ++j) total[j] += part[j];
		}
	}
	if (counts) memcpy(counts, total.data(), (size_t)2 * c->n_patterns * 4);
	if (stats) {
		c->st.n_candidates = c->st.n_hits = c->st.n_kmers = 0;
		for (Device &d : c->devs) {
			unsigned long long s[ST_N];
			CU(c, cudaSetDevice(d.ordinal));
			CU(c, cudaMemcpy(s, d.d_stats, sizeof s, cudaMemcpyDeviceToHost));
			c->st.n_candidates += s[ST_CANDIDATES];
			c->st.n_hits += s[ST_HITS];
			c->st.n_kmers += s[ST_KMERS];
		}
		*stats = c->st;
	}
	return VAFGPU_OK;
}

int vafgpu_export_counters(vafgpu_ctx *c, void *handle, size_t handle_bytes)
{
	if (!c || !handle || handle_bytes < sizeof(cudaIpcMemHandle_t)) return VAFGPU_EINVAL;
	if (c->attached || c->host_merge) return fail(c, VAFGPU_ESTATE, "this context does not own a single counter vector");
	cudaIpcMemHandle_t h;
	CU(c, cudaSetDevice(c->devs[0].ordinal));
	CU(c, cudaIpcGetMemHandle(&h, c->devs[0].d_counts));
	memset(handle, 0, handle_bytes);
	memcpy(handle, &h, sizeof h);
	return VAFGPU_OK;
}

int vafgpu_attach_counters(vafgpu_ctx *c, const void *handle, size_t handle_bytes)
{
	if (!c || !handle || handle_bytes < sizeof(cudaIpcMemHandle_t)) return VAFGPU_EINVAL;
	if (c->attached) return fail(c, VAFGPU_ESTATE, "counters are already attached");
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof h);
	for (Device &d : c->devs) {
		CU(c, cudaSetDevice(d.ordinal));
		CU(c, cudaIpcOpenMemHandle(&d.attached, h, cudaIpcMemLazyEnablePeerAccess));
		d.counts_to = static_cast<uint32_t *>(d.attached);
	}
	c->attached = true;
	return VAFGPU_OK;
}

int vafgpu_reset(vafgpu_ctx *c)
{
	if (!c) return VAFGPU_EINVAL;
	int rc = vafgpu_finish(c, nullptr, nullptr);
	if (rc) return rc;
	for (Device &d : c->devs) {
		CU(c, cudaSetDevice(d.ordinal));
		CU(c, cudaMemset(d.d_counts, 0, c->n_counts * 4));
		CU(c, cudaMemset(d.d_stats, 0, ST_N * sizeof(unsigned long long)));
	}
	vafgpu_stats keep = c->st;
	c->st = vafgpu_stats{};
	c->st.n_devices = keep.n_devices;
	c->st.anchor_stride = keep.anchor_stride;
	c->st.anchor_len = keep.anchor_len;
	c->st.filter_bytes = keep.filter_bytes;
	c->st.table_slots = keep.table_slots;
	c->st.filter_canon = keep.filter_canon;
	c->st.lookup_deferred = keep.lookup_deferred;
	c->st.kernel_threads = keep.kernel_threads;
	c->st.filter2_bytes = keep.filter2_bytes;
	return VAFGPU_OK;
}

void vafgpu_destroy(vafgpu_ctx *c)
{
	if (!c) return;
	for (Device &d : c->devs) destroy_device(d);
	delete c->def;
	delete c;
}

} // extern "C"
