/*
 * kc_count.c -- the kc-c4 command line on top of libvafgpu's counting mode (include/kcgpu.h).
 *
 * Same options, input and output as the reference tool (kc-c4.c:217-252):
 *   kc-c4 [-k INT] [-p INT] [-b INT] [-t INT] <in.fa>
 * prints the 255 lines "count<TAB>number of distinct canonical k-mers seen that often" (the
 * last line collects 255 and more) on stdout.  The host keeps option parsing, FASTA/FASTQ
 * parsing and the printing; count_seq_buf, the partitioned insert and the histogram scan
 * (kc-c4.c:74-128,186-215) run on the GPUs.
 *   -p  only checked (>= 10, kc-c4.c:243-246): the partition into 2^p tables is an internal of
 *       the reference; here the hash space is split over the visible GPUs instead
 *   -b  bases per turn when reads are dealt to several GPUs (the reference's block size)
 *   -t  number of host reader threads (the insert it used to parallelise runs on the GPU): a
 *       plain four-line FASTQ is cut into slices read in parallel (ingest.h), anything else
 *       goes through the one sequential reader
 * Environment: CUDA_VISIBLE_DEVICES / KCGPU_DEVICES=n select the GPUs (default: as many of the visible ones as
 *              the table of a file this size needs -- one up to ~50 GB of input -- more if it fills up);
 *              KCGPU_TABLE_SLOTS=n slots per GPU to start with (default: from the file size);
 *              either way a table that fills up is doubled and the file counted again.
 * Deviations: k outside 1..31 is rejected (the reference shifts by >= 64 bits there), and a
 * file that cannot be opened is an error (the reference dereferences NULL, kc-c4.c:166,247-248).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include "../../include/kcgpu.h"
#include "fastx.h"
#include "ingest.h"

/* staging blocks of 4 MiB: the readers set the pace, and page-locking one block per reader is start-up time */
#define STAGING_BYTES ((size_t)4 << 20)

static double now(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + tv.tv_usec * 1e-6;
}

/* one context per GPU, each made by a thread of its own: creating a CUDA context takes 0.2-2 s */
typedef struct {
	kcgpu_ctx **out;
	int k, device, rc;
	uint64_t slots, list_slots;
} create_job_t;

static void *create_one(void *arg)
{
	create_job_t *j = (create_job_t *)arg;
	j->rc = kcgpu_create(j->out, j->k, j->slots, j->list_slots, STAGING_BYTES, j->device);
	return NULL;
}

static int create_all(kcgpu_ctx **ctx, int n_dev, int k, uint64_t slots, uint64_t list_slots)
{
	create_job_t job[KCGPU_MAX_OWNERS];
	pthread_t th[KCGPU_MAX_OWNERS];
	int i, rc = VAFGPU_OK;
	for (i = 0; i < n_dev; ++i) {
		job[i].out = &ctx[i], job[i].k = k, job[i].device = i, job[i].rc = VAFGPU_OK, job[i].slots = slots, job[i].list_slots = list_slots;
		if (n_dev == 1 || pthread_create(&th[i], NULL, create_one, &job[i]) != 0) {
			create_one(&job[i]);
			th[i] = 0;
		}
	}
	for (i = 0; i < n_dev; ++i) {
		if (n_dev > 1 && th[i]) pthread_join(th[i], NULL);
		if (job[i].rc != VAFGPU_OK && rc == VAFGPU_OK) rc = job[i].rc;
	}
	return rc;
}

static uint64_t guess_slots(const char *fn, int n_dev)
{
	struct stat sb;
	uint64_t est;
	FILE *fp;
	unsigned char magic[2] = {0, 0};
	if (stat(fn, &sb) != 0 || !S_ISREG(sb.st_mode)) return 0; /* a pipe: as large as fits */
	est = (uint64_t)sb.st_size;
	if ((fp = fopen(fn, "rb")) != NULL) {
		if (fread(magic, 1, 2, fp) == 2 && magic[0] == 0x1f && magic[1] == 0x8b) est *= 4; /* gzip */
		fclose(fp);
	}
	est = est / (uint64_t)n_dev; /* at most one k-mer per byte, and in a FASTQ half the bytes are qualities:
	                              * the table is at most half full; a FASTA of all-distinct k-mers fills
	                              * it, and the file is then counted again with twice the slots */
	if (est < (1u << 20)) est = 1u << 20;
	return est;
}

/* the counting engine as a sink of the readers: every reader thread deals its reads to the
 * GPUs in turns of -b bases (kc-c4.c:150: a block is closed when it holds that many) */
typedef struct {
	kcgpu_ctx **ctx;
	int n_dev, block_size;
} engine_t;

typedef struct {
	engine_t *e;
	kcgpu_producer *prod[KCGPU_MAX_OWNERS];
	int turn;
	uint64_t in_turn;
} reader_t;

static int reader_destroy(void *producer)
{
	reader_t *r = (reader_t *)producer;
	int rc = 0;
	for (int i = 0; i < r->e->n_dev; ++i)
		if (r->prod[i] && kcgpu_producer_destroy(r->prod[i]) != VAFGPU_OK) rc = -1;
	free(r);
	return rc;
}

static int reader_create(void *engine, void **producer)
{
	static int next_turn = 0;
	engine_t *e = (engine_t *)engine;
	reader_t *r = (reader_t *)calloc(1, sizeof *r);
	if (!r) return -1;
	r->e = e;
	r->turn = __atomic_fetch_add(&next_turn, 1, __ATOMIC_RELAXED) % e->n_dev; /* readers start on different GPUs */
	for (int i = 0; i < e->n_dev; ++i)
		if (kcgpu_producer_create(e->ctx[i], &r->prod[i]) != VAFGPU_OK) {
			reader_destroy(r);
			return -1;
		}
	*producer = r;
	return 0;
}

static int reader_add_read(void *producer, const char *seq, size_t len)
{
	reader_t *r = (reader_t *)producer;
	if (kcgpu_producer_add_read(r->prod[r->turn], seq, len) != VAFGPU_OK) return -1;
	r->in_turn += len;
	if (r->in_turn >= (uint64_t)r->e->block_size) {
		r->in_turn = 0;
		r->turn = (r->turn + 1) % r->e->n_dev;
	}
	return 0;
}

static const char *engine_error(void *engine)
{
	engine_t *e = (engine_t *)engine;
	for (int i = 0; i < e->n_dev; ++i)
		if (kcgpu_strerror(e->ctx[i])[0]) return kcgpu_strerror(e->ctx[i]);
	return "unknown error";
}

int main(int argc, char *argv[])
{
	int c, k = 31, p = 10, block_size = 10000000, n_thread = 4, n_dev = 0, i;
	while ((c = getopt(argc, argv, "k:p:b:t:")) >= 0) {
		if (c == 'k') k = atoi(optarg);
		else if (c == 'p') p = atoi(optarg);
		else if (c == 'b') block_size = atoi(optarg);
		else if (c == 't') n_thread = atoi(optarg);
	}
	if (argc - optind < 1) { /* kc-c4.c:235-242 */
		fprintf(stderr, "Usage: kc-c4 [options] <in.fa>\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -k INT     k-mer size [%d]\n", k);
		fprintf(stderr, "  -p INT     prefix length [%d]\n", p);
		fprintf(stderr, "  -b INT     block size [%d]\n", block_size);
		fprintf(stderr, "  -t INT     number of worker threads [%d]\n", n_thread);
		return 1;
	}
	if (p < 10) { /* kc-c4.c:243-246 */
		fprintf(stderr, "ERROR: -p should be at least %d\n", 10);
		return 1;
	}
	if (k < 1 || k > 31) {
		fprintf(stderr, "ERROR: -k should be between 1 and 31\n");
		return 1;
	}
	if (block_size < 1) block_size = 1;
	const char *fn = argv[optind];

	n_dev = kcgpu_device_count();
	if (n_dev < 1) {
		fprintf(stderr, "ERROR: no CUDA device (this build has no CPU path)\n");
		return 1;
	}
	if (n_dev > KCGPU_MAX_OWNERS) n_dev = KCGPU_MAX_OWNERS;
	const int n_visible = n_dev;
	if (getenv("KCGPU_DEVICES") && atoi(getenv("KCGPU_DEVICES")) > 0) {
		if (atoi(getenv("KCGPU_DEVICES")) < n_dev) n_dev = atoi(getenv("KCGPU_DEVICES"));
	} else {
		/* as many GPUs as the table needs, not as many as there are: one GPU counts faster than the readers
		 * parse, and every further one costs a context and the peer mappings (seconds) before the first read */
		const uint64_t est = guess_slots(fn, 1); /* 0: a pipe, size unknown */
		const uint64_t per_gpu = (uint64_t)6 << 30; /* slots: 48 GB of table beside its lists and filter */
		if (est) {
			uint64_t need = (est + per_gpu - 1) / per_gpu;
			if (need < (uint64_t)n_dev) n_dev = (int)(need > 0 ? need : 1);
		}
	}

	const int direct = getenv("KCGPU_DIRECT") && atoi(getenv("KCGPU_DIRECT")); /* no region lists (development) */
	uint64_t slots = getenv("KCGPU_TABLE_SLOTS") ? strtoull(getenv("KCGPU_TABLE_SLOTS"), NULL, 10) : guess_slots(fn, n_dev);
	const int timing = getenv("KCGPU_TIMING") != NULL; /* phase times on stderr */
	uint64_t got_before = 0; /* slots per GPU the previous attempt really had */
	for (int attempt = 0;; ++attempt) {
		double t0 = now(), t1;
		kcgpu_ctx *ctx[KCGPU_MAX_OWNERS] = {0};
		uint64_t hist[256], part[256], overflow = 0;
		kcgpu_stats st;
		if (create_all(ctx, n_dev, k, slots, direct ? KCGPU_NO_LISTS : 0) != VAFGPU_OK) {
			fprintf(stderr, "ERROR: %s\n", kcgpu_strerror(NULL));
			return 1;
		}
		if (n_dev > 1 && kcgpu_link(ctx, n_dev) != VAFGPU_OK) {
			fprintf(stderr, "ERROR: %s\n", kcgpu_strerror(ctx[0]));
			return 1;
		}
		t1 = now();
		if (timing) fprintf(stderr, "[kc-c4] contexts (%llu slots per GPU)  %8.1f ms\n", (unsigned long long)slots, (t1 - t0) * 1e3);
		t0 = t1;
		engine_t eng = {ctx, n_dev, block_size};
		ingest_sink_t sink = {&eng, reader_create, reader_add_read, reader_destroy, engine_error};
		ingest_file_t info;
		char *files[1] = {(char *)fn};
		if (ingest_files_to(&sink, 1, files, k, block_size, n_thread, &info) != 0) return 1;
		if (!info.opened) {
			fprintf(stderr, "ERROR: cannot open %s\n", fn);
			return 1;
		}
		t1 = now();
		if (timing)
			fprintf(stderr, "[kc-c4] reading + counting (%d slices) %8.1f ms, %.1f Mbases/s\n", info.sliced, (t1 - t0) * 1e3,
			        info.bases / (t1 - t0) / 1e6);
		t0 = t1;
		memset(hist, 0, sizeof hist);
		for (i = 0; i < n_dev; ++i) {
			if (kcgpu_histogram(ctx[i], part, &st) != VAFGPU_OK) {
				fprintf(stderr, "ERROR: %s\n", kcgpu_strerror(ctx[i]));
				return 1;
			}
			for (int j = 0; j < 256; ++j) hist[j] += part[j];
			overflow += st.n_overflow;
			slots = st.table_slots;
		}
		t1 = now();
		if (timing) fprintf(stderr, "[kc-c4] flush + histogram               %8.1f ms (kernels %.1f ms, copies %.1f ms)\n", (t1 - t0) * 1e3, st.kernel_ms, st.h2d_ms);
		t0 = t1;
		for (i = 0; i < n_dev; ++i) kcgpu_destroy(ctx[i]);
		if (timing) fprintf(stderr, "[kc-c4] destroy                         %8.1f ms\n", (now() - t0) * 1e3);
		if (overflow) { /* never print a histogram with k-mers missing */
			struct stat sb;
			/* a pipe cannot be read twice; a request the device cannot hold is cut down by
			 * kcgpu_create, so a table that did not grow is the largest one that fits */
			if (attempt >= 12 || stat(fn, &sb) != 0 || !S_ISREG(sb.st_mode)) {
				fprintf(stderr, "ERROR: the k-mer table (%llu slots per GPU) is full\n", (unsigned long long)slots);
				return 1;
			}
			if (attempt && slots <= got_before) { /* the table cannot grow on this GPU: take more GPUs if there are any */
				if (n_dev >= n_visible) {
					fprintf(stderr, "ERROR: the k-mer table (%llu slots per GPU) is full\n", (unsigned long long)slots);
					return 1;
				}
				n_dev = n_dev * 2 < n_visible ? n_dev * 2 : n_visible;
				fprintf(stderr, "[kc-c4] table full, counting again on %d GPUs\n", n_dev);
				got_before = 0;
				continue;
			}
			got_before = slots;
			slots *= 2;
			fprintf(stderr, "[kc-c4] table full, counting again with %llu slots per GPU\n", (unsigned long long)slots);
			continue;
		}
		for (i = 1; i < 256; ++i) printf("%d\t%ld\n", i, (long)hist[i]); /* kc-c4.c:232-233 */
		return 0;
	}
}
