/*
 * kcgpu_api.cu -- the C ABI of include/kcgpu.h: one table with its region lists per context,
 * pinned staging blocks with a stream each (the copy of block i+1 overlaps the kernel of block
 * i, which replaces kt_pipeline's three steps, kc-c4.c:130-183), the flush of the lists into
 * the table when they are due, peer allocations for the fused several-GPU form.
 * Host logic only; the kernels are in kcgpu_kernels.cu.
 */
#include "../../include/kcgpu.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "kcgpu_kernels.cuh"

using namespace kcgpu;

namespace {

thread_local std::string g_kc_create_error;

/* One lock for every launch, flush and statistic of every context of the process: reader
 * threads (producers) fill their staging blocks outside it and take it only to fetch a block
 * and to submit one.  Contexts linked into a group flush together, hence not one lock each. */
std::mutex g_kc_mu;
std::condition_variable g_kc_block_free; /* a producer handed a staging block back */

struct KcBlock {
	char *h = nullptr;
	uint8_t *d = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr; /* copy start, copy end, kernel end */
	bool in_flight = false;
	bool owned = false; /* a producer is filling it */
	size_t used = 0;
};

} // namespace

struct kcgpu_ctx {
	int k = 0, device = 0, n_sm = 0;
	uint64_t n_slots = 0, list_cap = 0; /* list_cap: entries per region; 0 = no lists */
	uint32_t region_bits = 0, rslot_bits = 0;
	uint64_t *d_table = nullptr;        /* the whole allocation: table, lists, cursors */
	uint64_t pending_bytes = 0;         /* stream bytes filed since the last flush */
	uint64_t flush_bytes = 0;           /* flush before pending_bytes exceeds this */
	bool external_owners = false;       /* owners set by kcgpu_set_owners: the caller flushes */
	std::vector<std::pair<cudaStream_t, cudaEvent_t>> user_streams; /* last launch on each caller stream */
	cudaEvent_t f0 = nullptr, f1 = nullptr;
	unsigned long long *d_stats = nullptr, *d_hist = nullptr;
	InsertCtl ctl{KC_INS_COUNT, 0, 0}; /* what the insert step does, and the Bloom geometry */
	cudaStream_t main_stream = nullptr;
	size_t block_bytes = 0;
	std::vector<KcBlock *> blocks;
	size_t next_block = 0, n_producers = 0; /* n_producers: those made by kcgpu_producer_create and still alive */
	kcgpu_producer *def = nullptr; /* the producer behind kcgpu_add_read / kcgpu_submit_stream */
	uint32_t n_parts = 1, my_part = 0;
	uint64_t *tables[KC_MAX_PARTS] = {};
	std::vector<void *> ipc_mapped;
	std::vector<kcgpu_ctx *> group; /* contexts linked in this process, this one included */
	kcgpu_stats st{};
	std::string err;
};

/* One stream of reads being packed into staging blocks; one per reader thread. */
struct kcgpu_producer {
	kcgpu_ctx *c = nullptr;
	KcBlock *cur = nullptr;
	uint64_t n_reads = 0, n_bases = 0; /* added to the context's statistics when a block is submitted */
	std::vector<char> scratch;
};

namespace {

int kfail(kcgpu_ctx *c, int code, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (c) c->err = buf;
	else g_kc_create_error = buf;
	return code;
}

#define KCU(c, call)                                                                              \
	do {                                                                                          \
		cudaError_t e_ = (call);                                                                  \
		if (e_ != cudaSuccess)                                                                    \
			return kfail(c, VAFGPU_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

int kc_wait_block(kcgpu_ctx *c, KcBlock &b)
{
	if (!b.in_flight) return VAFGPU_OK;
	KCU(c, cudaEventSynchronize(b.e2));
	float h2d = 0, ker = 0;
	cudaEventElapsedTime(&h2d, b.e0, b.e1);
	cudaEventElapsedTime(&ker, b.e1, b.e2);
	c->st.h2d_ms += h2d;
	c->st.kernel_ms += ker;
	b.in_flight = false;
	return VAFGPU_OK;
}

CountArgs count_args(const kcgpu_ctx *c, const void *bytes, size_t n)
{
	CountArgs a{};
	a.bytes = static_cast<const uint8_t *>(bytes);
	a.n_bytes = n;
	a.first_chunk = 0;
	a.end_chunk = n >> 4;
	a.n_slots = c->n_slots;
	a.list_cap = c->list_cap;
	a.k = c->k;
	a.n_parts = c->n_parts;
	a.region_bits = c->region_bits;
	a.rslot_bits = c->rslot_bits;
	for (uint32_t i = 0; i < c->n_parts; ++i) a.tables[i] = c->tables[i];
	a.stats = c->d_stats;
	a.ctl = c->ctl;
	return a;
}

int kc_flush_group(kcgpu_ctx *c, bool submit_staged = true);

cudaError_t kc_launch_scan(const kcgpu_ctx *c, const CountArgs &a, cudaStream_t s)
{
	if (!c->list_cap) return launch_count(a, s);
	return c->n_parts > 1 ? launch_push(a, s) : launch_partition(a, s);
}

/* what the lists (one owner) or the inbox (several owners) of m hold, into its table */
cudaError_t kc_launch_flush(const kcgpu_ctx *m, cudaStream_t s)
{
	uint64_t *lists = kc_lists_of(m->d_table, m->n_slots);
	unsigned long long *cursors = kc_cursors_of(m->d_table, m->n_slots, m->list_cap, m->region_bits);
	uint32_t *bloom = m->ctl.bloom_bits ? kc_bloom_of(m->d_table, m->n_slots, m->list_cap, m->region_bits) : nullptr;
	if (m->n_parts == 1)
		return launch_flush(m->d_table, lists, cursors, m->list_cap, m->region_bits, m->rslot_bits, bloom, m->ctl, m->d_stats, s);
	/* the inbox (first half of the list area) into the region lists (second half), then those */
	RouteArgs a{};
	a.inbox = lists;
	a.n_ptr = kc_inbox_cursor(m->d_table, m->n_slots, m->list_cap, m->region_bits);
	a.inbox_cap = kc_inbox_cap(m->list_cap, m->region_bits);
	a.lists = lists + a.inbox_cap;
	a.cursors = cursors;
	a.cap = m->list_cap / 2;
	a.table = m->d_table;
	a.region_bits = m->region_bits;
	a.rslot_bits = m->rslot_bits;
	a.bloom = bloom;
	a.ctl = m->ctl;
	a.stats = m->d_stats;
	cudaError_t e = launch_route(a, m->n_sm, s);
	if (e != cudaSuccess) return e;
	return launch_flush(m->d_table, a.lists, cursors, a.cap, m->region_bits, m->rslot_bits, bloom, m->ctl, m->d_stats, s);
}

/* the lists must be able to take n more k-mers: flush first if they might not (a context whose
 * owners the caller named flushes only when told to) */
int kc_make_room(kcgpu_ctx *c, uint64_t n_bytes)
{
	if (!c->list_cap || c->external_owners) return VAFGPU_OK;
	uint64_t pending = 0;
	for (const kcgpu_ctx *m : c->group) pending += m->pending_bytes;
	/* a flush that is merely due does not touch what the built-in producers have staged: their
	 * threads may be writing there */
	if (pending && pending + n_bytes > c->flush_bytes * c->group.size()) return kc_flush_group(c, false);
	return VAFGPU_OK;
}

/* lock held.  One more staging block for the context. */
int kc_add_block(kcgpu_ctx *c)
{
	KcBlock *b = new (std::nothrow) KcBlock;
	if (!b) return kfail(c, VAFGPU_ENOMEM, "out of memory");
	c->blocks.push_back(b);
	KCU(c, cudaSetDevice(c->device));
	KCU(c, cudaHostAlloc(&b->h, c->block_bytes + 64, cudaHostAllocPortable));
	KCU(c, cudaMalloc(&b->d, c->block_bytes + 64));
	KCU(c, cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	KCU(c, cudaEventCreate(&b->e0));
	KCU(c, cudaEventCreate(&b->e1));
	KCU(c, cudaEventCreate(&b->e2));
	return VAFGPU_OK;
}

/* lock held.  Copy the producer's block to the device and scan it there, on the block's stream. */
int kc_submit_current(kcgpu_producer *p)
{
	kcgpu_ctx *c = p->c;
	KcBlock *b = p->cur;
	c->st.n_reads += p->n_reads;
	c->st.n_bases += p->n_bases;
	p->n_reads = p->n_bases = 0;
	if (!b) return VAFGPU_OK;
	p->cur = nullptr;
	b->owned = false;
	g_kc_block_free.notify_one();
	if (!b->used) return VAFGPU_OK;
	const size_t n = (b->used + 15) & ~(size_t)15;
	memset(b->h + b->used, '\n', n - b->used);
	int rc = kc_make_room(c, n);
	if (rc) return rc;
	KCU(c, cudaSetDevice(c->device));
	KCU(c, cudaEventRecord(b->e0, b->stream));
	KCU(c, cudaMemcpyAsync(b->d, b->h, n, cudaMemcpyHostToDevice, b->stream));
	KCU(c, cudaEventRecord(b->e1, b->stream));
	KCU(c, kc_launch_scan(c, count_args(c, b->d, n), b->stream));
	KCU(c, cudaEventRecord(b->e2, b->stream));
	b->in_flight = true;
	c->pending_bytes += n;
	c->st.n_blocks++;
	return VAFGPU_OK;
}

/* Takes the lock.  Room for `need` more bytes in the producer's block: submit it if it is
 * full, fetch a free one if there is none (waiting for a block in flight is the back-pressure
 * that replaces kt_pipeline's "at most three blocks"). */
int kc_ensure_room(kcgpu_producer *p, size_t need)
{
	kcgpu_ctx *c = p->c;
	if (p->cur && p->cur->used + need <= c->block_bytes) return VAFGPU_OK;
	KcBlock *pick = nullptr;
	{
		std::unique_lock<std::mutex> lk(g_kc_mu);
		if (p->cur) {
			int rc = kc_submit_current(p);
			if (rc) return rc;
		}
		for (;;) {
			/* an idle block if there is one, else the first one in flight in ring order */
			KCU(c, cudaSetDevice(c->device));
			for (size_t tries = 0; tries < c->blocks.size(); ++tries) {
				KcBlock *b = c->blocks[(c->next_block + tries) % c->blocks.size()];
				if (b->owned) continue;
				if (!b->in_flight || cudaEventQuery(b->e2) == cudaSuccess) {
					pick = b;
					break;
				}
				if (!pick) pick = b;
			}
			cudaGetLastError(); /* cudaErrorNotReady of the query is not an error */
			if (pick) break;
			g_kc_block_free.wait(lk); /* more producers than blocks: wait for one to be submitted */
		}
		c->next_block = (c->next_block + 1) % c->blocks.size();
		pick->owned = true; /* ours from here on: nobody else looks at it */
	}
	/* wait for the block's last use OUTSIDE the lock: other producers, of this and of every other
	 * context, go on fetching and submitting meanwhile */
	float h2d = 0, ker = 0;
	const bool was_in_flight = pick->in_flight;
	if (was_in_flight) {
		KCU(c, cudaSetDevice(c->device));
		KCU(c, cudaEventSynchronize(pick->e2));
		cudaEventElapsedTime(&h2d, pick->e0, pick->e1);
		cudaEventElapsedTime(&ker, pick->e1, pick->e2);
	}
	{
		std::lock_guard<std::mutex> lk(g_kc_mu);
		if (was_in_flight && pick->in_flight) { /* a flush in between may have accounted for it already */
			c->st.h2d_ms += h2d;
			c->st.kernel_ms += ker;
			pick->in_flight = false;
		}
	}
	pick->used = 0;
	p->cur = pick;
	return VAFGPU_OK;
}

/* lock held.  Submit what the context's own producer has staged (other producers submit
 * their own when they are flushed or destroyed) and wait for everything in flight. */
int kc_sync_one(kcgpu_ctx *c, bool submit_staged = true)
{
	int rc = c->def && submit_staged ? kc_submit_current(c->def) : VAFGPU_OK;
	if (rc) return rc;
	KCU(c, cudaSetDevice(c->device));
	for (KcBlock *b : c->blocks) {
		rc = kc_wait_block(c, *b);
		if (rc) return rc;
	}
	KCU(c, cudaStreamSynchronize(c->main_stream));
	for (auto &us : c->user_streams) KCU(c, cudaEventSynchronize(us.second));
	return VAFGPU_OK;
}

/* every member: wait for what was filed, empty the lists into the table, wait for that */
int kc_flush_group(kcgpu_ctx *c, bool submit_staged)
{
	for (kcgpu_ctx *m : c->group) {
		int rc = kc_sync_one(m, submit_staged);
		if (rc) {
			if (m != c) c->err = m->err;
			return rc;
		}
	}
	for (kcgpu_ctx *m : c->group) {
		if (!m->list_cap) continue;
		KCU(m, cudaSetDevice(m->device));
		KCU(m, cudaEventRecord(m->f0, m->main_stream));
		KCU(m, kc_launch_flush(m, m->main_stream));
		KCU(m, cudaMemsetAsync(kc_cursors_of(m->d_table, m->n_slots, m->list_cap, m->region_bits), 0,
		                       (size_t)kc_cursor_bytes(m->region_bits), m->main_stream));
		KCU(m, cudaEventRecord(m->f1, m->main_stream));
	}
	for (kcgpu_ctx *m : c->group) {
		if (!m->list_cap) continue;
		KCU(m, cudaSetDevice(m->device));
		KCU(m, cudaStreamSynchronize(m->main_stream));
		float ms = 0;
		cudaEventElapsedTime(&ms, m->f0, m->f1);
		m->st.kernel_ms += ms;
		m->st.n_flushes++;
		m->pending_bytes = 0;
	}
	return VAFGPU_OK;
}

/* a launch went to a stream of the caller's: remember where it ends */
int kc_note_user_stream(kcgpu_ctx *c, cudaStream_t s)
{
	if (s == c->main_stream) return VAFGPU_OK;
	for (auto &us : c->user_streams)
		if (us.first == s) {
			KCU(c, cudaEventRecord(us.second, s));
			return VAFGPU_OK;
		}
	cudaEvent_t e;
	KCU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	c->user_streams.emplace_back(s, e);
	KCU(c, cudaEventRecord(e, s));
	return VAFGPU_OK;
}

int kc_read_stats(kcgpu_ctx *c)
{
	unsigned long long s[KC_ST_N];
	KCU(c, cudaSetDevice(c->device));
	KCU(c, cudaMemcpy(s, c->d_stats, sizeof s, cudaMemcpyDeviceToHost));
	c->st.n_kmers = s[KC_ST_KMERS];
	c->st.n_distinct = s[KC_ST_NEW];
	c->st.n_overflow = s[KC_ST_OVERFLOW];
	c->st.n_dropped = s[KC_ST_DROPPED];
	c->st.n_direct = s[KC_ST_DIRECT];
	return VAFGPU_OK;
}

} // namespace

extern "C" {

uint64_t kcgpu_hash64(uint64_t key, int k)
{
	if (k < 1 || k > 31) return 0;
	return kc_hash64(key, (1ULL << 2 * k) - 1);
}

int kcgpu_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

const char *kcgpu_strerror(const kcgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_kc_create_error.c_str(); }

int kcgpu_create(kcgpu_ctx **out, int k, uint64_t table_slots, uint64_t list_slots, size_t block_bytes, int device)
{
	return kcgpu_create_filtered(out, k, table_slots, list_slots, block_bytes, device, 0, 0);
}

int kcgpu_create_filtered(kcgpu_ctx **out, int k, uint64_t table_slots, uint64_t list_slots, size_t block_bytes, int device,
                          int bloom_bits, int bloom_hashes)
{
	if (!out) return kfail(nullptr, VAFGPU_EINVAL, "ctx is NULL");
	if (bloom_bits < 0 || bloom_bits > 40 || bloom_hashes < 0 || bloom_hashes > 64)
		return kfail(nullptr, VAFGPU_EINVAL, "Bloom filter of 2^%d bits with %d hash functions", bloom_bits, bloom_hashes);
	if (bloom_bits && bloom_bits < 13) bloom_bits = 13; /* at least a kilobyte */
	if (!bloom_hashes) bloom_bits = 0;
	*out = nullptr;
	if (k < 1 || k > 31) return kfail(nullptr, VAFGPU_EINVAL, "k = %d is outside 1..31", k);
	int visible = 0;
	cudaError_t ce = cudaGetDeviceCount(&visible);
	if (ce != cudaSuccess || visible < 1)
		return kfail(nullptr, VAFGPU_ENOGPU, "no CUDA device: %s", ce == cudaSuccess ? "count is 0" : cudaGetErrorString(ce));
	if (device < 0 || device >= visible) return kfail(nullptr, VAFGPU_EINVAL, "device %d of %d", device, visible);
	if (block_bytes == 0) block_bytes = (size_t)16 << 20;
	if (block_bytes < 4096) block_bytes = 4096;
	if (block_bytes > ((size_t)1 << 34)) return kfail(nullptr, VAFGPU_EINVAL, "block_bytes too large");

	kcgpu_ctx *c = new (std::nothrow) kcgpu_ctx;
	if (!c) return kfail(nullptr, VAFGPU_ENOMEM, "out of memory");
	c->k = k;
	c->device = device;
	c->block_bytes = block_bytes;
	int rc = [&]() -> int {
		cudaDeviceProp prop;
		KCU(c, cudaSetDevice(device));
		KCU(c, cudaGetDeviceProperties(&prop, device));
		if (prop.major != 10)
			return kfail(c, VAFGPU_ENOGPU, "device %d (%s) is sm_%d%d; this library carries sm_100a code only", device, prop.name,
			             prop.major, prop.minor);
		c->n_sm = prop.multiProcessorCount;
		const uint32_t need_bits = kc_region_bits(k); /* regions the slot word needs (tag beside the count) */
		const uint64_t min_slots = (uint64_t)4096 > ((uint64_t)16 << need_bits) ? (uint64_t)4096 : ((uint64_t)16 << need_bits);
		const bool lists = list_slots != KCGPU_NO_LISTS;
		size_t free_b = 0, total_b = 0;
		KCU(c, cudaMemGetInfo(&free_b, &total_b));
		/* a filter is worth at most a quarter of the device */
		while (bloom_bits > 13 && kc_bloom_bytes((uint32_t)bloom_bits) > (uint64_t)free_b / 4) --bloom_bits;
		free_b -= (size_t)kc_bloom_bytes((uint32_t)bloom_bits);
		c->ctl.bloom_bits = (uint32_t)bloom_bits;
		c->ctl.bloom_hashes = (uint32_t)bloom_hashes;
		if (table_slots == 0) {
			uint64_t budget = (uint64_t)free_b / 4 * 3 / 8; /* 8-byte words */
			if (lists) budget = list_slots ? (budget > list_slots ? budget - list_slots : 0) : budget / 3 * 2;
			table_slots = min_slots;
			while (table_slots * 2 <= budget) table_slots *= 2;
		}
		uint64_t n = min_slots;
		while (n < table_slots) {
			if (n >> 40) return kfail(c, VAFGPU_EINVAL, "table_slots too large");
			n *= 2;
		}
		/* the geometry of a table of n slots; a request that does not fit the device is halved
		 * until it does (the caller reads the size it got from kcgpu_stats.table_slots) */
		const uint64_t want_lists = list_slots;
		uint64_t alloc = 0;
		for (;; n /= 2) {
			c->n_slots = n;
			uint32_t bits = 0;
			while ((1ull << bits) < n) ++bits;
			/* regions: what the slot word needs, and small enough (64 MiB) to stay in L2 while the
			 * lists of one region are emptied into it */
			c->region_bits = need_bits;
			if (lists && bits > KC_REGION_SLOT_BITS && bits - KC_REGION_SLOT_BITS > c->region_bits) c->region_bits = bits - KC_REGION_SLOT_BITS;
			if (c->region_bits > 20) c->region_bits = 20;
			c->rslot_bits = bits - c->region_bits;
			c->list_cap = 0;
			if (lists) {
				list_slots = want_lists ? want_lists : n / 2;
				uint64_t cap = (list_slots >> c->region_bits) + 31 & ~(uint64_t)31;
				if (cap < 64) cap = 64;
				if (cap >> 40) return kfail(c, VAFGPU_EINVAL, "list_slots too large");
				c->list_cap = cap;
				c->flush_bytes = (cap << c->region_bits) / 20 * 19; /* a byte is at most one k-mer; what a list cannot take goes to the table */
				if (c->flush_bytes > ((uint64_t)1 << 40)) c->flush_bytes = (uint64_t)1 << 40;
			}
			alloc = kc_alloc_bytes(n, c->list_cap, c->region_bits, c->ctl.bloom_bits);
			/* (a table with more regions than a tile is sorted by would not fit any device: 2^35 slots) */
			if (lists && c->region_bits > KC_TILE_REGION_BITS && n > min_slots) continue;
			if (alloc - kc_bloom_bytes(c->ctl.bloom_bits) + ((uint64_t)256 << 20) <= (uint64_t)free_b || n <= min_slots) break;
		}
		cudaError_t me = cudaMalloc(&c->d_table, alloc);
		if (me != cudaSuccess) {
			cudaGetLastError();
			return kfail(c, VAFGPU_ENOMEM, "cannot allocate %llu MiB on device %d for %llu slots and lists of %llu: %s",
			             (unsigned long long)(alloc >> 20), device, (unsigned long long)n,
			             (unsigned long long)(c->list_cap << c->region_bits), cudaGetErrorString(me));
		}
		KCU(c, cudaMalloc(&c->d_stats, KC_ST_N * sizeof(unsigned long long)));
		KCU(c, cudaMalloc(&c->d_hist, 1024 * sizeof(unsigned long long)));
		KCU(c, cudaMemset(c->d_table, 0, n * 8));
		if (c->list_cap)
			KCU(c, cudaMemset(kc_cursors_of(c->d_table, n, c->list_cap, c->region_bits), 0, (size_t)kc_cursor_bytes(c->region_bits)));
		if (c->ctl.bloom_bits)
			KCU(c, cudaMemset(kc_bloom_of(c->d_table, n, c->list_cap, c->region_bits), 0, (size_t)kc_bloom_bytes(c->ctl.bloom_bits)));
		KCU(c, cudaMemset(c->d_stats, 0, KC_ST_N * sizeof(unsigned long long)));
		KCU(c, cudaStreamCreateWithFlags(&c->main_stream, cudaStreamNonBlocking));
		KCU(c, cudaEventCreate(&c->f0));
		KCU(c, cudaEventCreate(&c->f1));
		for (int i = 0; i < 3; ++i) {
			int brc = kc_add_block(c);
			if (brc) return brc;
		}
		c->def = new (std::nothrow) kcgpu_producer;
		if (!c->def) return kfail(c, VAFGPU_ENOMEM, "out of memory");
		c->def->c = c;
		KCU(c, cudaDeviceSynchronize());
		return VAFGPU_OK;
	}();
	if (rc != VAFGPU_OK) {
		g_kc_create_error = c->err;
		kcgpu_destroy(c);
		return rc;
	}
	c->tables[0] = c->d_table;
	c->group.push_back(c);
	c->st.table_slots = c->n_slots;
	c->st.list_slots = c->list_cap << c->region_bits;
	c->st.flush_bytes = c->flush_bytes;
	*out = c;
	return VAFGPU_OK;
}

int kcgpu_producer_create(kcgpu_ctx *c, kcgpu_producer **out)
{
	if (!c || !out) return VAFGPU_EINVAL;
	kcgpu_producer *p = new (std::nothrow) kcgpu_producer;
	if (!p) return kfail(c, VAFGPU_ENOMEM, "out of memory");
	p->c = c;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	/* one block per live reader to fill, on top of the three in flight; the blocks of readers that
	 * are gone (the first pass of a two-pass count) are taken over */
	if (c->blocks.size() < 3 + c->n_producers + 1) {
		int rc = kc_add_block(c);
		if (rc) {
			delete p;
			return rc;
		}
	}
	c->n_producers++;
	*out = p;
	return VAFGPU_OK;
}

int kcgpu_producer_add_read(kcgpu_producer *p, const char *seq, size_t len)
{
	if (!p || (!seq && len)) return VAFGPU_EINVAL;
	kcgpu_ctx *c = p->c;
	if (len < (size_t)c->k) return VAFGPU_OK; /* kc-c4.c:141 */
	p->n_reads++;
	p->n_bases += len;
	if (len + 1 <= c->block_bytes) {
		int rc = kc_ensure_room(p, len + 1);
		if (rc) return rc;
		KcBlock *b = p->cur;
		vafgpu_canonicalise_read(seq, len, b->h + b->used, 0); /* strict table: kc-c4.c:21-38 */
		b->h[b->used + len] = '\n';
		b->used += len + 1;
		return VAFGPU_OK;
	}
	/* a read longer than a block (a chromosome): pieces that overlap by k-1 bases, so that
	 * every k-mer lies in exactly one piece */
	if (p->scratch.size() < len) p->scratch.resize(len);
	vafgpu_canonicalise_read(seq, len, p->scratch.data(), 0);
	const size_t piece = c->block_bytes - 1, step = piece - (size_t)(c->k - 1);
	for (size_t at = 0;; at += step) {
		const size_t n = len - at < piece ? len - at : piece;
		int rc = kc_ensure_room(p, n + 1);
		if (rc) return rc;
		KcBlock *b = p->cur;
		memcpy(b->h + b->used, p->scratch.data() + at, n);
		b->h[b->used + n] = '\n';
		b->used += n + 1;
		if (at + n >= len) break;
	}
	return VAFGPU_OK;
}

int kcgpu_producer_flush(kcgpu_producer *p)
{
	if (!p) return VAFGPU_EINVAL;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	return kc_submit_current(p);
}

int kcgpu_producer_destroy(kcgpu_producer *p)
{
	if (!p) return VAFGPU_OK;
	int rc = kcgpu_producer_flush(p);
	{
		std::lock_guard<std::mutex> lk(g_kc_mu);
		if (p->c->n_producers) p->c->n_producers--;
	}
	delete p;
	return rc;
}

int kcgpu_add_read(kcgpu_ctx *c, const char *seq, size_t len)
{
	if (!c) return VAFGPU_EINVAL;
	return kcgpu_producer_add_read(c->def, seq, len);
}

int kcgpu_submit_stream(kcgpu_ctx *c, const char *bytes, size_t n_bytes)
{
	if (!c || (!bytes && n_bytes)) return VAFGPU_EINVAL;
	kcgpu_producer *p = c->def;
	cudaPointerAttributes attr;
	const bool pinned = cudaPointerGetAttributes(&attr, bytes) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError(); /* a plain malloc pointer is reported as an error by older drivers */
	size_t at = 0;
	while (at < n_bytes) {
		/* a fresh block: what kcgpu_add_read or the last round left open is submitted first */
		int rc = kc_ensure_room(p, c->block_bytes + 1);
		if (rc) return rc;
		KcBlock *b = p->cur;
		size_t n = n_bytes - at, advance;
		bool add_nl = false;
		if (n > c->block_bytes) {
			/* cut after the last separator that fits; a read longer than a block is cut with a
			 * k-1 overlap, as in kcgpu_add_read */
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
		at += advance;
		if (!pinned) {
			memcpy(b->h, bytes + at - advance, n);
			if (add_nl) b->h[n++] = '\n';
			b->used = n;
			continue; /* submitted by the next kc_ensure_room or after the loop */
		}
		/* zero-copy: H2D straight from the caller's buffer, separator and padding written on the device */
		std::lock_guard<std::mutex> lk(g_kc_mu);
		p->cur = nullptr;
		b->owned = false;
		g_kc_block_free.notify_one();
		const size_t n16 = (n + (add_nl ? 1 : 0) + 15) & ~(size_t)15;
		rc = kc_make_room(c, n16);
		if (rc) return rc;
		KCU(c, cudaSetDevice(c->device));
		KCU(c, cudaEventRecord(b->e0, b->stream));
		KCU(c, cudaMemcpyAsync(b->d, bytes + at - advance, n, cudaMemcpyHostToDevice, b->stream));
		if (n16 > n) KCU(c, cudaMemsetAsync(b->d + n, '\n', n16 - n, b->stream));
		KCU(c, cudaEventRecord(b->e1, b->stream));
		KCU(c, kc_launch_scan(c, count_args(c, b->d, n16), b->stream));
		KCU(c, cudaEventRecord(b->e2, b->stream));
		b->in_flight = true;
		c->pending_bytes += n16;
		c->st.n_blocks++;
	}
	std::lock_guard<std::mutex> lk(g_kc_mu);
	return kc_submit_current(p);
}

int kcgpu_count_device(kcgpu_ctx *c, const void *d_bytes, size_t n_bytes, void *stream)
{
	if (!c) return VAFGPU_EINVAL;
	if (((uintptr_t)d_bytes & 15) || (n_bytes & 15))
		return kfail(c, VAFGPU_EINVAL, "device stream must be 16-byte aligned and a multiple of 16 bytes");
	std::lock_guard<std::mutex> lk(g_kc_mu);
	cudaStream_t s = stream ? (cudaStream_t)stream : c->main_stream;
	/* as much at a time as the lists are sure to take, a flush in between */
	const uint64_t n_chunks = n_bytes >> 4;
	uint64_t step = n_chunks;
	if (c->list_cap && !c->external_owners) {
		step = c->flush_bytes * c->group.size() >> 4;
		if (step < 4096) step = 4096;
	}
	for (uint64_t lo = 0; lo < n_chunks;) {
		uint64_t pending = 0;
		for (const kcgpu_ctx *m : c->group) pending += m->pending_bytes;
		uint64_t room = step > (pending >> 4) ? step - (pending >> 4) : 0;
		if (c->list_cap && !c->external_owners && room < 4096 && pending) {
			int rc = kc_flush_group(c);
			if (rc) return rc;
			continue;
		}
		const uint64_t hi = lo + room < n_chunks && room ? lo + room : n_chunks;
		CountArgs a = count_args(c, d_bytes, n_bytes);
		a.first_chunk = lo;
		a.end_chunk = hi;
		KCU(c, cudaSetDevice(c->device));
		KCU(c, kc_launch_scan(c, a, s));
		int rc = kc_note_user_stream(c, s);
		if (rc) return rc;
		c->pending_bytes += (hi - lo) << 4;
		c->st.n_blocks++;
		lo = hi;
	}
	return VAFGPU_OK;
}

int kcgpu_extract_device(kcgpu_ctx *c, const void *d_bytes, size_t n_bytes, int n_parts, uint64_t *d_keys,
                         size_t cap_per_part, uint32_t *d_part_counts, void *stream)
{
	if (!c) return VAFGPU_EINVAL;
	if (n_parts < 1 || n_parts > KC_MAX_PARTS) return kfail(c, VAFGPU_EINVAL, "n_parts = %d is outside 1..%d", n_parts, KC_MAX_PARTS);
	if (((uintptr_t)d_bytes & 15) || (n_bytes & 15))
		return kfail(c, VAFGPU_EINVAL, "device stream must be 16-byte aligned and a multiple of 16 bytes");
	if (!d_keys || !d_part_counts) return kfail(c, VAFGPU_EINVAL, "output lists are NULL");
	if (cap_per_part >= ((size_t)1 << 32)) return kfail(c, VAFGPU_EINVAL, "cap_per_part must be below 2^32");
	CountArgs a = count_args(c, d_bytes, n_bytes);
	a.n_parts = (uint32_t)n_parts;
	a.out_keys = d_keys;
	a.cap_per_part = cap_per_part;
	a.part_counts = d_part_counts;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	KCU(c, cudaSetDevice(c->device));
	cudaStream_t s = stream ? (cudaStream_t)stream : c->main_stream;
	KCU(c, launch_extract(a, s));
	c->st.n_blocks++;
	return kc_note_user_stream(c, s);
}

int kcgpu_insert_device(kcgpu_ctx *c, const uint64_t *d_hashed_keys, size_t n, int n_parts, void *stream)
{
	if (!c) return VAFGPU_EINVAL;
	if (n_parts < 1 || n_parts > KC_MAX_PARTS) return kfail(c, VAFGPU_EINVAL, "n_parts = %d is outside 1..%d", n_parts, KC_MAX_PARTS);
	if (n && !d_hashed_keys) return kfail(c, VAFGPU_EINVAL, "keys are NULL");
	InsertArgs a{};
	a.hashed = d_hashed_keys;
	a.n = n;
	a.n_parts = (uint32_t)n_parts;
	a.region_bits = c->region_bits;
	a.rslot_bits = c->rslot_bits;
	a.table = c->d_table;
	a.n_slots = c->n_slots;
	a.list_cap = c->list_cap;
	a.ctl = c->ctl;
	a.stats = c->d_stats;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	KCU(c, cudaSetDevice(c->device));
	cudaStream_t s = stream ? (cudaStream_t)stream : c->main_stream;
	KCU(c, launch_insert(a, c->n_sm, s));
	return kc_note_user_stream(c, s);
}

int kcgpu_table(kcgpu_ctx *c, void **d_table, uint64_t *table_slots)
{
	if (!c) return VAFGPU_EINVAL;
	if (d_table) *d_table = c->d_table;
	if (table_slots) *table_slots = c->n_slots;
	return VAFGPU_OK;
}

int kcgpu_ipc_export(kcgpu_ctx *c, void *handle)
{
	if (!c || !handle) return VAFGPU_EINVAL;
	static_assert(sizeof(cudaIpcMemHandle_t) == KCGPU_IPC_HANDLE_BYTES, "IPC handle size");
	cudaIpcMemHandle_t h;
	KCU(c, cudaSetDevice(c->device));
	KCU(c, cudaIpcGetMemHandle(&h, c->d_table));
	memcpy(handle, &h, sizeof h);
	return VAFGPU_OK;
}

int kcgpu_ipc_open(kcgpu_ctx *c, const void *handle, void **d_peer_table)
{
	if (!c || !handle || !d_peer_table) return VAFGPU_EINVAL;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof h);
	void *p = nullptr;
	KCU(c, cudaSetDevice(c->device));
	KCU(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	c->ipc_mapped.push_back(p);
	*d_peer_table = p;
	return VAFGPU_OK;
}

static int kc_set_owners(kcgpu_ctx *c, int n_parts, int my_part, void *const *tables);

int kcgpu_set_owners(kcgpu_ctx *c, int n_parts, int my_part, void *const *tables)
{
	if (!c) return VAFGPU_EINVAL;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	return kc_set_owners(c, n_parts, my_part, tables);
}

/* lock held */
static int kc_set_owners(kcgpu_ctx *c, int n_parts, int my_part, void *const *tables)
{
	if (n_parts < 1 || n_parts > KC_MAX_PARTS || my_part < 0 || my_part >= n_parts || !tables)
		return kfail(c, VAFGPU_EINVAL, "owner %d of %d", my_part, n_parts);
	/* nothing this context filed may still be running, or waiting in a list, under the old
	 * owners.  A context that has filed nothing leaves its lists alone: a peer that already
	 * knows this allocation may be filing into them right now */
	int rc = c->pending_bytes ? kc_flush_group(c) : kc_sync_one(c);
	if (rc) return rc;
	c->external_owners = true;
	for (int i = 0; i < n_parts; ++i) {
		if (!tables[i] && i != my_part) return kfail(c, VAFGPU_EINVAL, "table of owner %d is NULL", i);
		c->tables[i] = tables[i] ? static_cast<uint64_t *>(tables[i]) : c->d_table;
	}
	c->n_parts = (uint32_t)n_parts;
	c->my_part = (uint32_t)my_part;
	/* several owners: half of the list area is the inbox, the other half the region lists */
	c->flush_bytes = (c->list_cap << c->region_bits) / 20 * 19 / (n_parts > 1 ? 2 : 1);
	c->st.flush_bytes = c->flush_bytes;
	return VAFGPU_OK;
}

int kcgpu_link(kcgpu_ctx *const *ctxs, int n)
{
	if (!ctxs || n < 1 || n > KC_MAX_PARTS) return VAFGPU_EINVAL;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	for (int i = 0; i < n; ++i) {
		if (!ctxs[i]) return VAFGPU_EINVAL;
		if (ctxs[i]->n_slots != ctxs[0]->n_slots || ctxs[i]->k != ctxs[0]->k || ctxs[i]->list_cap != ctxs[0]->list_cap ||
		    ctxs[i]->ctl.bloom_bits != ctxs[0]->ctl.bloom_bits || ctxs[i]->ctl.bloom_hashes != ctxs[0]->ctl.bloom_hashes)
			return kfail(ctxs[i], VAFGPU_EINVAL, "linked contexts must share k, the table size, the list size and the Bloom filter size");
	}
	for (int i = 0; i < n; ++i) {
		kcgpu_ctx *c = ctxs[i];
		KCU(c, cudaSetDevice(c->device));
		for (int j = 0; j < n; ++j) {
			if (ctxs[j]->device == c->device) continue;
			int ok = 0;
			KCU(c, cudaDeviceCanAccessPeer(&ok, c->device, ctxs[j]->device));
			if (!ok) return kfail(c, VAFGPU_ECUDA, "device %d cannot reach device %d's memory", c->device, ctxs[j]->device);
			cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
			if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
			else if (e != cudaSuccess) return kfail(c, VAFGPU_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
		}
	}
	std::vector<void *> tables(n);
	for (int i = 0; i < n; ++i) tables[i] = ctxs[i]->d_table;
	for (int i = 0; i < n; ++i) {
		int rc = kc_set_owners(ctxs[i], n, i, tables.data());
		if (rc) return rc;
	}
	for (int i = 0; i < n; ++i) {
		ctxs[i]->group.assign(ctxs, ctxs + n);
		ctxs[i]->external_owners = false; /* the group is in this process: it flushes itself */
	}
	return VAFGPU_OK;
}

static int kc_sync_group(kcgpu_ctx *c);

int kcgpu_sync(kcgpu_ctx *c)
{
	if (!c) return VAFGPU_EINVAL;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	return kc_sync_group(c);
}

/* lock held */
static int kc_sync_group(kcgpu_ctx *c)
{
	for (kcgpu_ctx *m : c->group) {
		int rc = kc_sync_one(m);
		if (rc) {
			if (m != c) c->err = m->err;
			return rc;
		}
	}
	return VAFGPU_OK;
}

int kcgpu_flush(kcgpu_ctx *c)
{
	if (!c) return VAFGPU_EINVAL;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	return kc_flush_group(c);
}

int kcgpu_histogram(kcgpu_ctx *c, uint64_t hist[256], kcgpu_stats *stats)
{
	if (!c) return VAFGPU_EINVAL;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	int rc = kc_flush_group(c);
	if (rc) return rc;
	KCU(c, cudaSetDevice(c->device));
	if (hist) {
		KCU(c, cudaMemsetAsync(c->d_hist, 0, 256 * sizeof(unsigned long long), c->main_stream));
		KCU(c, launch_histogram(c->d_table, c->n_slots, c->d_hist, c->n_sm, c->main_stream));
		unsigned long long h[256];
		KCU(c, cudaMemcpyAsync(h, c->d_hist, sizeof h, cudaMemcpyDeviceToHost, c->main_stream));
		KCU(c, cudaStreamSynchronize(c->main_stream));
		for (int i = 0; i < 256; ++i) hist[i] = h[i];
		hist[0] = 0;
	}
	if (stats) {
		rc = kc_read_stats(c);
		if (rc) return rc;
		*stats = c->st;
	}
	return VAFGPU_OK;
}

int kcgpu_set_pass(kcgpu_ctx *c, int pass)
{
	if (!c) return VAFGPU_EINVAL;
	if (pass != KCGPU_PASS_COUNT && pass != KCGPU_PASS_CLAIM && pass != KCGPU_PASS_LOOKUP)
		return kfail(c, VAFGPU_EINVAL, "pass %d", pass);
	std::lock_guard<std::mutex> lk(g_kc_mu);
	/* what was filed under the old rule is inserted under the old rule */
	int rc = c->external_owners ? kc_sync_group(c) : kc_flush_group(c);
	if (rc) return rc;
	for (kcgpu_ctx *m : c->group) m->ctl.mode = pass;
	return VAFGPU_OK;
}

int kcgpu_histogram1024(kcgpu_ctx *c, uint64_t hist[1024], int min_count, int max_count, kcgpu_stats *stats)
{
	if (!c) return VAFGPU_EINVAL;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	int rc = kc_flush_group(c);
	if (rc) return rc;
	KCU(c, cudaSetDevice(c->device));
	if (hist) {
		KCU(c, cudaMemsetAsync(c->d_hist, 0, 1024 * sizeof(unsigned long long), c->main_stream));
		KCU(c, launch_histogram1024(c->d_table, c->n_slots, c->d_hist, c->n_sm, c->main_stream));
		unsigned long long h[1024];
		KCU(c, cudaMemcpyAsync(h, c->d_hist, sizeof h, cudaMemcpyDeviceToHost, c->main_stream));
		KCU(c, cudaStreamSynchronize(c->main_stream));
		for (int i = 0; i < 1024; ++i) hist[i] = i >= min_count && i <= max_count ? h[i] : 0; /* yak_ch_shrink, yak-count.c:247-282 */
	}
	if (stats) {
		rc = kc_read_stats(c);
		if (rc) return rc;
		*stats = c->st;
	}
	return VAFGPU_OK;
}

int kcgpu_reset(kcgpu_ctx *c)
{
	if (!c) return VAFGPU_EINVAL;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	int rc = kc_sync_group(c);
	if (rc) return rc;
	KCU(c, cudaSetDevice(c->device));
	KCU(c, cudaMemsetAsync(c->d_table, 0, c->n_slots * 8, c->main_stream));
	if (c->list_cap)
		KCU(c, cudaMemsetAsync(kc_cursors_of(c->d_table, c->n_slots, c->list_cap, c->region_bits), 0,
		                       (size_t)kc_cursor_bytes(c->region_bits), c->main_stream));
	if (c->ctl.bloom_bits)
		KCU(c, cudaMemsetAsync(kc_bloom_of(c->d_table, c->n_slots, c->list_cap, c->region_bits), 0, (size_t)kc_bloom_bytes(c->ctl.bloom_bits),
		                       c->main_stream));
	KCU(c, cudaMemsetAsync(c->d_stats, 0, KC_ST_N * sizeof(unsigned long long), c->main_stream));
	KCU(c, cudaStreamSynchronize(c->main_stream));
	const kcgpu_stats keep = c->st;
	c->st = kcgpu_stats{};
	c->st.table_slots = keep.table_slots;
	c->st.list_slots = keep.list_slots;
	c->st.flush_bytes = keep.flush_bytes;
	c->pending_bytes = 0;
	return VAFGPU_OK;
}

void kcgpu_destroy(kcgpu_ctx *c)
{
	if (!c) return;
	std::lock_guard<std::mutex> lk(g_kc_mu);
	cudaSetDevice(c->device);
	for (kcgpu_ctx *m : c->group) /* the others must not wait for a context that is gone */
		if (m != c) {
			for (size_t i = 0; i < m->group.size(); ++i)
				if (m->group[i] == c) {
					m->group.erase(m->group.begin() + i);
					break;
				}
		}
	for (KcBlock *b : c->blocks) {
		if (b->stream) cudaStreamSynchronize(b->stream);
		if (b->h) cudaFreeHost(b->h);
		if (b->d) cudaFree(b->d);
		if (b->e0) cudaEventDestroy(b->e0);
		if (b->e1) cudaEventDestroy(b->e1);
		if (b->e2) cudaEventDestroy(b->e2);
		if (b->stream) cudaStreamDestroy(b->stream);
		delete b;
	}
	delete c->def;
	if (c->main_stream) {
		cudaStreamSynchronize(c->main_stream);
		cudaStreamDestroy(c->main_stream);
	}
	for (auto &us : c->user_streams) cudaEventDestroy(us.second);
	if (c->f0) cudaEventDestroy(c->f0);
	if (c->f1) cudaEventDestroy(c->f1);
	for (void *p : c->ipc_mapped) cudaIpcCloseMemHandle(p);
	cudaFree(c->d_table);
	cudaFree(c->d_stats);
	cudaFree(c->d_hist);
	delete c;
}

} // extern "C"
