/*
 * ingest.c -- see ingest.h.
 */
#define _GNU_SOURCE
#include "ingest.h"

#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include "fastx.h"

#define DEFAULT_SLICE ((size_t)32 << 20)

static double now(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + tv.tv_usec * 1e-6;
}

/* ---- strictly formed four-line FASTQ records in a memory image ---- */

/* The record starting at o.  Returns 1 and sets the sequence and the offset of the next record;
 * 0 at end of input (only line feeds may follow); -1 if what is there is not a strictly formed
 * record (the caller then gives the whole file to the sequential reader). */
static int strict_record(const unsigned char *m, size_t n, size_t o, size_t *seq, size_t *len, size_t *next)
{
	const unsigned char *p;
	size_t s, l, plus, q;
	if (o >= n) return 0;
	if (m[o] != '@') {
		for (; o < n; ++o)
			if (m[o] != '\n') return -1;
		return 0;
	}
	p = (const unsigned char *)memchr(m + o, '\n', n - o); /* header line */
	if (!p) return -1;
	s = (size_t)(p - m) + 1;
	if (s >= n || m[s] == '@' || m[s] == '+' || m[s] == '>' || m[s] == '\n') return -1;
	p = (const unsigned char *)memchr(m + s, '\n', n - s); /* the one sequence line */
	if (!p) return -1;
	l = (size_t)(p - m) - s;
	if (m[s + l - 1] == '\r') return -1;
	plus = s + l + 1;
	if (plus >= n || m[plus] != '+') return -1;
	p = (const unsigned char *)memchr(m + plus, '\n', n - plus);
	if (!p) return -1;
	q = (size_t)(p - m) + 1; /* the one quality line: exactly l bytes */
	if (q + l > n || memchr(m + q, '\n', l)) return -1;
	if (q + l < n && m[q + l] != '\n') return -1;
	*seq = s;
	*len = l;
	*next = q + l < n ? q + l + 1 : n;
	return 1;
}

/* first offset >= lo that looks like the start of a record: a line that starts with '@' whose
 * second successor starts with '+' (a quality line may start with '@', but then the line two
 * further on is a sequence line).  n if there is none nearby. */
static size_t guess_record(const unsigned char *m, size_t n, size_t lo)
{
	size_t o = lo;
	int tries;
	if (lo > 0) {
		const unsigned char *p = (const unsigned char *)memchr(m + lo - 1, '\n', n - (lo - 1));
		if (!p) return n;
		o = (size_t)(p - m) + 1;
	}
	for (tries = 0; tries < 8 && o < n; ++tries) {
		const unsigned char *p1 = (const unsigned char *)memchr(m + o, '\n', n - o), *p2;
		if (!p1) return n;
		if (m[o] == '@') {
			p2 = (const unsigned char *)memchr(p1 + 1, '\n', n - (size_t)(p1 + 1 - m));
			if (p2 && (size_t)(p2 - m) + 1 < n && p2[1] == '+') return o;
		}
		o = (size_t)(p1 - m) + 1;
	}
	return n;
}

/* ---- units and the thread pool ---- */

typedef struct {
	const char *fn;
	ingest_file_t *out;
	const unsigned char *map; /* NULL: sequential */
	size_t size, slice;
	int n_slices;
	size_t *guess, *end; /* per slice */
	int *bad;
	double t0, t1;
} file_t;

typedef struct {
	int file, slice; /* slice < 0: the whole file through the sequential reader */
} unit_t;

typedef struct {
	const ingest_sink_t *sink;
	int k, block_len, pass;
	file_t *files;
	unit_t *units;
	int n_units;
	int next; /* work counter */
	int failed;
	pthread_mutex_t mu;
} pool_t;

static void add_totals(pool_t *pl, file_t *f, uint64_t seqs, uint64_t bases, double t0, double t1)
{
	pthread_mutex_lock(&pl->mu);
	f->out->seqs += seqs;
	f->out->bases += bases;
	if (f->t0 == 0 || t0 < f->t0) f->t0 = t0;
	if (t1 > f->t1) f->t1 = t1;
	pthread_mutex_unlock(&pl->mu);
}

static int engine_failed(pool_t *pl)
{
	pthread_mutex_lock(&pl->mu);
	if (!pl->failed) fprintf(stderr, "Error: %s\n", pl->sink->error(pl->sink->engine));
	pl->failed = 1;
	pthread_mutex_unlock(&pl->mu);
	return -1;
}

/* one whole file: the step-0 loop of vaf-counter.c:486-517 feeding the engine.  The reference
 * closes a block when it holds >= block_len bases or the reader returns < 0.  An empty block
 * retires the pipeline worker that read it (kthread.c:97-125: a worker leaves when ITS step 0
 * returns NULL), the other two of kt_pipeline(3, ...) go on calling step 0 in order, so the file
 * ends with the third empty block; reproduced so that a malformed FASTQ record ends (or does
 * not end) the file at the same place.  kc-c4.c:133-183 is the same loop. */
static int run_sequential(pool_t *pl, file_t *f, void *prod)
{
	fastx_t *fx = fastx_open(f->fn);
	uint64_t seqs = 0, bases = 0;
	double t0 = now();
	if (!fx) return 0; /* vaf-counter.c:557: silently skipped */
	f->out->opened = 1;
	int lives = 3; /* the three workers of kt_pipeline(3, ...), see below */
	for (;;) {
		long l, sum_len = 0;
		const char *s;
		while ((l = fastx_next(fx, &s)) >= 0) {
			if (l < pl->k) continue;
			if (pl->sink->producer_add_read(prod, s, (size_t)l) != 0) {
				fastx_close(fx);
				return engine_failed(pl);
			}
			sum_len += l;
			++seqs;
			bases += (uint64_t)l;
			if (sum_len >= pl->block_len) break;
		}
		if (sum_len == 0 && --lives == 0) break;
	}
	fastx_close(fx);
	add_totals(pl, f, seqs, bases, t0, now());
	return 0;
}

/* one slice: pass 1 walks and validates, pass 2 walks again and feeds the engine */
static int run_slice(pool_t *pl, file_t *f, int i, void *prod)
{
	const size_t hi = (size_t)(i + 1) * f->slice < f->size ? (size_t)(i + 1) * f->slice : f->size;
	size_t o, seq, len, next;
	uint64_t seqs = 0, bases = 0;
	double t0 = now();
	int rc = 1;
	if (pl->pass == 1) f->guess[i] = guess_record(f->map, f->size, (size_t)i * f->slice);
	o = f->guess[i];
	while (o < hi && (rc = strict_record(f->map, f->size, o, &seq, &len, &next)) == 1) {
		if (pl->pass == 2 && len >= (size_t)pl->k) {
			if (pl->sink->producer_add_read(prod, (const char *)f->map + seq, len) != 0) return engine_failed(pl);
			++seqs;
			bases += len;
		}
		o = next;
	}
	if (pl->pass == 1) {
		f->end[i] = rc == 0 ? f->size : o; /* rc == 0: nothing but line feeds up to the end of the file */
		f->bad[i] = rc < 0;
	} else add_totals(pl, f, seqs, bases, t0, now());
	return 0;
}

static void *worker(void *arg)
{
	pool_t *pl = (pool_t *)arg;
	void *prod = NULL;
	if (pl->pass == 2 && pl->sink->producer_create(pl->sink->engine, &prod) != 0) {
		engine_failed(pl);
		return NULL;
	}
	for (;;) {
		int u;
		pthread_mutex_lock(&pl->mu);
		u = pl->failed ? pl->n_units : pl->next++;
		pthread_mutex_unlock(&pl->mu);
		if (u >= pl->n_units) break;
		file_t *f = &pl->files[pl->units[u].file];
		if (pl->units[u].slice < 0) run_sequential(pl, f, prod);
		else run_slice(pl, f, pl->units[u].slice, prod);
	}
	if (prod && pl->sink->producer_destroy(prod) != 0) engine_failed(pl);
	return NULL;
}

static void run_pool(pool_t *pl, int n_threads)
{
	pthread_t *t;
	int i, n = n_threads < pl->n_units ? n_threads : pl->n_units;
	if (n < 1) return;
	pl->next = 0;
	t = (pthread_t *)calloc((size_t)n, sizeof *t);
	for (i = 1; i < n; ++i) pthread_create(&t[i], NULL, worker, pl);
	worker(pl); /* the calling thread is reader 0 */
	for (i = 1; i < n; ++i) pthread_join(t[i], NULL);
	free(t);
}

/* map a file if it is a plain four-line FASTQ worth cutting up */
static void try_map(file_t *f, size_t slice)
{
	struct stat st;
	int fd = open(f->fn, O_RDONLY);
	unsigned char magic[2];
	void *m;
	if (fd < 0) return;
	if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || (size_t)st.st_size < 2 * slice ||
	    pread(fd, magic, 2, 0) != 2 || magic[0] != '@') { /* gzip starts 1f 8b, FASTA '>' */
		close(fd);
		return;
	}
	m = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (m == MAP_FAILED) return;
	madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
	f->map = (const unsigned char *)m;
	f->size = (size_t)st.st_size;
	f->slice = slice;
	f->n_slices = (int)((f->size + slice - 1) / slice);
	f->guess = (size_t *)calloc((size_t)f->n_slices, sizeof(size_t));
	f->end = (size_t *)calloc((size_t)f->n_slices, sizeof(size_t));
	f->bad = (int *)calloc((size_t)f->n_slices, sizeof(int));
}

static void unmap(file_t *f)
{
	if (f->map) munmap((void *)f->map, f->size);
	free(f->guess);
	free(f->end);
	free(f->bad);
	f->map = NULL;
	f->guess = f->end = NULL;
	f->bad = NULL;
	f->n_slices = 0;
}

int ingest_files_to(const ingest_sink_t *sink, int n_files, char **files, int k, int block_len, int n_threads,
                    ingest_file_t *out)
{
	pool_t pl;
	file_t *fs = (file_t *)calloc((size_t)(n_files > 0 ? n_files : 1), sizeof *fs);
	const char *env = getenv("VAFGPU_SLICE_BYTES"); /* testing knob */
	size_t slice = env && atoll(env) > 0 ? (size_t)atoll(env) : DEFAULT_SLICE;
	int i, j, n_units = 0, total_slices = 0;
	if (n_threads < 1) n_threads = 1;
	memset(&pl, 0, sizeof pl);
	pthread_mutex_init(&pl.mu, NULL);
	pl.sink = sink, pl.k = k, pl.block_len = block_len, pl.files = fs;
	for (i = 0; i < n_files; ++i) {
		memset(&out[i], 0, sizeof out[i]);
		fs[i].fn = files[i];
		fs[i].out = &out[i];
		if (n_threads > 1) try_map(&fs[i], slice);
		total_slices += fs[i].n_slices;
	}
	pl.units = (unit_t *)calloc((size_t)(total_slices + n_files + 1), sizeof(unit_t));

	/* pass 1: validate the slices */
	for (i = 0; i < n_files; ++i)
		for (j = 0; j < fs[i].n_slices; ++j) pl.units[n_units].file = i, pl.units[n_units++].slice = j;
	pl.n_units = n_units;
	pl.pass = 1;
	run_pool(&pl, n_threads);
	for (i = 0; i < n_files; ++i) {
		file_t *f = &fs[i];
		int ok = f->n_slices > 0 && f->guess[0] == 0 && f->end[f->n_slices - 1] == f->size;
		for (j = 0; ok && j < f->n_slices; ++j) ok = !f->bad[j] && (j + 1 == f->n_slices || f->end[j] == f->guess[j + 1]);
		if (!ok) unmap(f); /* not one chain of strictly formed records: read it sequentially */
	}

	/* pass 2: whole files first (they are the long poles), then the slices */
	n_units = 0;
	for (i = 0; i < n_files; ++i)
		if (!fs[i].map) pl.units[n_units].file = i, pl.units[n_units++].slice = -1;
	for (i = 0; i < n_files; ++i)
		for (j = 0; j < fs[i].n_slices; ++j) pl.units[n_units].file = i, pl.units[n_units++].slice = j;
	pl.n_units = n_units;
	pl.pass = 2;
	run_pool(&pl, n_threads);

	for (i = 0; i < n_files; ++i) {
		if (fs[i].map) out[i].opened = 1;
		out[i].sliced = fs[i].n_slices;
		out[i].seconds = fs[i].t1 > fs[i].t0 ? fs[i].t1 - fs[i].t0 : 0;
		unmap(&fs[i]);
	}
	free(pl.units);
	free(fs);
	pthread_mutex_destroy(&pl.mu);
	return pl.failed ? -1 : 0;
}

/* the vaf-counter engine as a sink */
static int vaf_producer_create(void *engine, void **producer) { return vafgpu_producer_create((vafgpu_ctx *)engine, (vafgpu_producer **)producer); }
static int vaf_producer_add_read(void *producer, const char *seq, size_t len) { return vafgpu_producer_add_read((vafgpu_producer *)producer, seq, len); }
static int vaf_producer_destroy(void *producer) { return vafgpu_producer_destroy((vafgpu_producer *)producer); }
static const char *vaf_error(void *engine) { return vafgpu_strerror((const vafgpu_ctx *)engine); }

int ingest_files(vafgpu_ctx *ctx, int n_files, char **files, int k, int block_len, int n_threads, ingest_file_t *out)
{
	ingest_sink_t sink = {ctx, vaf_producer_create, vaf_producer_add_read, vaf_producer_destroy, vaf_error};
	return ingest_files_to(&sink, n_files, files, k, block_len, n_threads, out);
}
