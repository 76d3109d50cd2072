/*
 * yak_count.c -- the yak-count command line on top of libvafgpu's counting mode (include/kcgpu.h).
 *
 * Same options, input and output as the reference tool (yak-count.c:460-507):
 *   yak-count [-k INT] [-p INT] [-b INT] [-H INT] [-t INT] [-K INT] <in.fa> [in.fa]
 * prints the 1023 lines "count<TAB>number of distinct canonical k-mers seen that often" on
 * stdout.  Without -b every k-mer is counted in one pass.  With -b the first file is read once
 * to find the k-mers worth an entry -- those a Bloom filter of 2^b bits has seen before -- the
 * second file (or the first one again) is read to count them, and entries seen fewer than twice
 * are dropped (yak-count.c:445-456).  The host keeps option parsing, FASTA/FASTQ parsing and the
 * printing; extraction, filter, tables and the histogram scan run on the GPUs.
 *   -p  only checked (>= 10, yak-count.c:492-495): the partition into 2^p tables is an internal of
 *       the reference; here the hash space is split over the visible GPUs instead
 *   -K  bases per turn when reads are dealt to several GPUs (the reference's chunk size)
 *   -t  number of host reader threads
 *   -b  the filter gets 2^b bits on every GPU (cut down to a quarter of its memory)
 * Environment: CUDA_VISIBLE_DEVICES / KCGPU_DEVICES=n select the GPUs (default: as many of the visible ones as
 *              the table of a file this size needs -- one up to ~50 GB of input -- more if it fills up);
 *              KCGPU_TABLE_SLOTS=n slots per GPU to start with (default: from the file size);
 *              either way a table that fills up is doubled and the files counted again;
 *              KCGPU_TIMING: phase times on stderr.
 * Deviations: k outside 1..31 is rejected; a file that cannot be opened is an error (the
 * reference dereferences NULL); with -b AND a second file the entries kept can differ from the
 * reference's by false positives of the filter (include/kcgpu.h), as they do between two values
 * of -b in the reference itself.
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

typedef struct {
	kcgpu_ctx **ctx;
	int n_dev, chunk;
} engine_t;

/* staging blocks of 4 MiB: the readers set the pace, and page-locking one block per reader is start-up time */
#define STAGING_BYTES ((size_t)4 << 20)

static double now(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + tv.tv_usec * 1e-6;
}

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
	r->turn = __atomic_fetch_add(&next_turn, 1, __ATOMIC_RELAXED) % e->n_dev;
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
	if (r->in_turn >= (uint64_t)r->e->chunk) {
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

/* one context per GPU, each made by a thread of its own: creating a CUDA context takes 0.2-2 s */
typedef struct {
	kcgpu_ctx **out;
	int k, device, rc;
	uint64_t slots;
	int bloom_bits, bloom_hashes;
} create_job_t;

static void *create_one(void *arg)
{
	create_job_t *j = (create_job_t *)arg;
	j->rc = kcgpu_create_filtered(j->out, j->k, j->slots, 0, STAGING_BYTES, j->device, j->bloom_bits, j->bloom_hashes);
	return NULL;
}

static int create_all(kcgpu_ctx **ctx, int n_dev, int k, uint64_t slots, int bloom_bits, int bloom_hashes)
{
	create_job_t job[KCGPU_MAX_OWNERS];
	pthread_t th[KCGPU_MAX_OWNERS];
	int i, rc = VAFGPU_OK;
	for (i = 0; i < n_dev; ++i) {
		job[i].out = &ctx[i], job[i].k = k, job[i].device = i, job[i].rc = VAFGPU_OK, job[i].slots = slots, job[i].bloom_bits = bloom_bits, job[i].bloom_hashes = bloom_hashes;
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
	est /= (uint64_t)n_dev;
	return est < (1u << 20) ? 1u << 20 : est;
}

/* one pass over one file; 0, or 1 after printing what went wrong */
static int count_file(engine_t *eng, const char *fn, int k, int n_thread)
{
	ingest_sink_t sink = {eng, reader_create, reader_add_read, reader_destroy, engine_error};
	ingest_file_t info;
	char *files[1] = {(char *)fn};
	if (ingest_files_to(&sink, 1, files, k, eng->chunk, n_thread, &info) != 0) return 1;
	if (!info.opened) {
		fprintf(stderr, "ERROR: cannot open %s\n", fn);
		return 1;
	}
	fprintf(stderr, "[M] processed %llu sequences of %s\n", (unsigned long long)info.seqs, fn);
	return 0;
}

int main(int argc, char *argv[])
{
	int c, k = 31, pre = 10, bf_shift = 0, bf_n_hash = 4, n_thread = 4, chunk = 10000000, n_dev, i; /* yak-count.c:312-321 */
	while ((c = getopt(argc, argv, "k:p:K:t:b:H:")) >= 0) {
		if (c == 'k') k = atoi(optarg);
		else if (c == 'p') pre = atoi(optarg);
		else if (c == 'K') chunk = atoi(optarg);
		else if (c == 't') n_thread = atoi(optarg);
		else if (c == 'b') bf_shift = atoi(optarg);
		else if (c == 'H') bf_n_hash = atoi(optarg);
	}
	if (argc - optind < 1) { /* yak-count.c:481-491 */
		fprintf(stderr, "Usage: yak-count [options] <in.fa> [in.fa]\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -k INT     k-mer size [%d]\n", k);
		fprintf(stderr, "  -p INT     prefix length [%d]\n", pre);
		fprintf(stderr, "  -b INT     set Bloom filter size to 2**INT bits; 0 to disable [%d]\n", bf_shift);
		fprintf(stderr, "  -H INT     use INT hash functions for Bloom filter [%d]\n", bf_n_hash);
		fprintf(stderr, "  -t INT     number of worker threads [%d]\n", n_thread);
		fprintf(stderr, "  -K INT     chunk size [100m]\n");
		fprintf(stderr, "Note: -b37 is recommended for human reads\n");
		return 1;
	}
	if (pre < 10) { /* yak-count.c:492-495 */
		fprintf(stderr, "ERROR: -p should be at least %d\n", 10);
		return 1;
	}
	if (k < 1 || k > 31) {
		fprintf(stderr, "ERROR: -k should be between 1 and 31\n");
		return 1;
	}
	if (chunk < 1) chunk = 1;
	const char *fn1 = argv[optind], *fn2 = argc - optind >= 2 ? argv[optind + 1] : fn1;
	const int two_pass = bf_shift > 0; /* yak-count.c:449: decided by -b alone, whether or not a filter comes of it */
	/* the reference builds a filter when -H > 0 and -b > -p, one per partition of at least one block (yak-count.c:75,117) */
	const int filter = bf_n_hash > 0 && bf_shift > pre && bf_shift - pre >= 9;

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
		const uint64_t est = guess_slots(fn1, 1); /* 0: a pipe, size unknown */
		const uint64_t per_gpu = (uint64_t)6 << 30; /* slots: 48 GB of table beside its lists and filter */
		if (est) {
			uint64_t need = (est + per_gpu - 1) / per_gpu;
			if (need < (uint64_t)n_dev) n_dev = (int)(need > 0 ? need : 1);
		}
	}

	uint64_t slots = getenv("KCGPU_TABLE_SLOTS") ? strtoull(getenv("KCGPU_TABLE_SLOTS"), NULL, 10) : guess_slots(fn1, n_dev);
	uint64_t got_before = 0;
	const int timing = getenv("KCGPU_TIMING") != NULL;
	for (int attempt = 0;; ++attempt) {
		double t0 = now(), t1;
		kcgpu_ctx *ctx[KCGPU_MAX_OWNERS] = {0};
		uint64_t hist[1024], part[1024], overflow = 0, tot = 0;
		kcgpu_stats st;
		if (create_all(ctx, n_dev, k, slots, filter ? (bf_shift > 40 ? 40 : bf_shift) : 0, bf_n_hash > 64 ? 64 : bf_n_hash) != VAFGPU_OK) {
			fprintf(stderr, "ERROR: %s\n", kcgpu_strerror(NULL));
			return 1;
		}
		if (n_dev > 1 && kcgpu_link(ctx, n_dev) != VAFGPU_OK) {
			fprintf(stderr, "ERROR: %s\n", kcgpu_strerror(ctx[0]));
			return 1;
		}
		engine_t eng = {ctx, n_dev, chunk};
		if (timing) fprintf(stderr, "[yak-count] contexts (%llu slots per GPU)  %8.1f ms\n", (unsigned long long)slots, ((t1 = now()) - t0) * 1e3), t0 = t1;
		/* yak_count_file, yak-count.c:445-456 */
		/* with -b the counts of the first pass are thrown away (yak-count.c:452): it only makes the
		 * entries -- for the k-mers the filter has seen before or, where no filter comes of -b
		 * (yak-count.c:75,117), for all of them */
		if (kcgpu_set_pass(ctx[0], two_pass ? KCGPU_PASS_CLAIM : KCGPU_PASS_COUNT) != VAFGPU_OK || count_file(&eng, fn1, k, n_thread)) {
			if (kcgpu_strerror(ctx[0])[0]) fprintf(stderr, "ERROR: %s\n", kcgpu_strerror(ctx[0]));
			return 1;
		}
		if (timing) fprintf(stderr, "[yak-count] first pass                      %8.1f ms\n", ((t1 = now()) - t0) * 1e3), t0 = t1;
		if (two_pass && (kcgpu_set_pass(ctx[0], KCGPU_PASS_LOOKUP) != VAFGPU_OK || count_file(&eng, fn2, k, n_thread))) {
			if (kcgpu_strerror(ctx[0])[0]) fprintf(stderr, "ERROR: %s\n", kcgpu_strerror(ctx[0]));
			return 1;
		}
		if (timing && two_pass) fprintf(stderr, "[yak-count] second pass                     %8.1f ms\n", ((t1 = now()) - t0) * 1e3), t0 = t1;
		memset(hist, 0, sizeof hist);
		for (i = 0; i < n_dev; ++i) {
			if (kcgpu_histogram1024(ctx[i], part, two_pass ? 2 : 0, 1023, &st) != VAFGPU_OK) { /* the shrink, yak-count.c:453 */
				fprintf(stderr, "ERROR: %s\n", kcgpu_strerror(ctx[i]));
				return 1;
			}
			for (int j = 0; j < 1024; ++j) hist[j] += part[j], tot += part[j];
			overflow += st.n_overflow;
			slots = st.table_slots;
		}
		if (timing) fprintf(stderr, "[yak-count] flush + histogram               %8.1f ms (kernels %.1f ms, copies %.1f ms)\n", ((t1 = now()) - t0) * 1e3, st.kernel_ms, st.h2d_ms), t0 = t1;
		for (i = 0; i < n_dev; ++i) kcgpu_destroy(ctx[i]);
		if (timing) fprintf(stderr, "[yak-count] destroy                         %8.1f ms\n", (now() - t0) * 1e3);
		if (overflow) { /* never print a histogram with k-mers missing */
			struct stat sb;
			if (attempt >= 12 || stat(fn1, &sb) != 0 || !S_ISREG(sb.st_mode) || stat(fn2, &sb) != 0 || !S_ISREG(sb.st_mode)) {
				fprintf(stderr, "ERROR: the k-mer table (%llu slots per GPU) is full\n", (unsigned long long)slots);
				return 1;
			}
			if (attempt && slots <= got_before) { /* the table cannot grow on this GPU: take more GPUs if there are any */
				if (n_dev >= n_visible) {
					fprintf(stderr, "ERROR: the k-mer table (%llu slots per GPU) is full\n", (unsigned long long)slots);
					return 1;
				}
				n_dev = n_dev * 2 < n_visible ? n_dev * 2 : n_visible;
				fprintf(stderr, "[yak-count] table full, counting again on %d GPUs\n", n_dev);
				got_before = 0;
				continue;
			}
			got_before = slots;
			slots *= 2;
			fprintf(stderr, "[yak-count] table full, counting again with %llu slots per GPU\n", (unsigned long long)slots);
			continue;
		}
		fprintf(stderr, "[M::%s] %ld distinct k-mers after shrinking\n", __func__, (long)tot); /* yak-count.c:499 */
		for (i = 1; i < 1024; ++i) printf("%d\t%lld\n", i, (long long)hist[i]);                    /* yak-count.c:503 */
		return 0;
	}
}
