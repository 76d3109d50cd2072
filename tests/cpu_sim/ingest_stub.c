/*
 * ingest_stub.c -- TEST-ONLY stand-in for the engine behind host/ingest.c: the producer entry
 * points of include/vafgpu.h collect an order-independent digest of the reads they are handed,
 * so that the CPU test-suite can check that sliced, multi-threaded ingest hands over exactly
 * the reads the sequential reader does.  Built by tests/util.py into tests/_build/; never part
 * of the product.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../kmer-cnt_b200/host/ingest.h"

struct vafgpu_ctx { int unused; };
struct vafgpu_producer { uint64_t n, bases, sum, xr; };

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static uint64_t g_n, g_bases, g_sum, g_xor;
static int g_k;

static uint64_t fnv(const char *s, size_t n)
{
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)s[i]) * 1099511628211ull;
	return h;
}

int vafgpu_producer_create(vafgpu_ctx *ctx, vafgpu_producer **p)
{
	(void)ctx;
	*p = (vafgpu_producer *)calloc(1, sizeof **p);
	return *p ? VAFGPU_OK : VAFGPU_ENOMEM;
}
int vafgpu_producer_add_read(vafgpu_producer *p, const char *seq, size_t len)
{
	if (len < (size_t)g_k) return VAFGPU_OK;
	uint64_t h = fnv(seq, len);
	p->n++, p->bases += len, p->sum += h, p->xr ^= h * 0x9E3779B97F4A7C15ull;
	return VAFGPU_OK;
}
int vafgpu_producer_flush(vafgpu_producer *p) { (void)p; return VAFGPU_OK; }
int vafgpu_producer_destroy(vafgpu_producer *p)
{
	pthread_mutex_lock(&g_mu);
	g_n += p->n, g_bases += p->bases, g_sum += p->sum, g_xor ^= p->xr;
	pthread_mutex_unlock(&g_mu);
	free(p);
	return VAFGPU_OK;
}
const char *vafgpu_strerror(const vafgpu_ctx *ctx) { (void)ctx; return "stub"; }

/* out: [reads, bases, sum of hashes, xor of hashes, per-file seqs..., per-file bases..., per-file slices...] */
int stub_ingest(int n_files, char **files, int k, int block_len, int n_threads, uint64_t *out)
{
	vafgpu_ctx ctx;
	ingest_file_t *pf = (ingest_file_t *)calloc((size_t)n_files + 1, sizeof *pf);
	g_n = g_bases = g_sum = g_xor = 0;
	g_k = k;
	int rc = ingest_files(&ctx, n_files, files, k, block_len, n_threads, pf);
	out[0] = g_n, out[1] = g_bases, out[2] = g_sum, out[3] = g_xor;
	for (int i = 0; i < n_files; ++i) {
		out[4 + i] = pf[i].seqs;
		out[4 + n_files + i] = pf[i].bases;
		out[4 + 2 * n_files + i] = (uint64_t)pf[i].sliced | (uint64_t)pf[i].opened << 32;
	}
	free(pf);
	return rc;
}
