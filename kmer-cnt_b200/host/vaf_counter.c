/*
 * vaf_counter.c -- the vaf-counter command line on top of libvafgpu.
 *
 * Same options, inputs, messages and output as the reference tool (vaf-counter.c:584-738):
 *   vaf-counter [-k INT] [-t INT] [-b INT] [-v] -p patterns.txt -o out.vaf reads.fq [...]
 * The host keeps what the reference keeps on the host -- option parsing, the pattern file,
 * the first-insert-wins k-mer map, FASTA/FASTQ parsing, the VAF writer -- and hands every
 * parsed read to the GPU engine instead of the kt_pipeline extract/lookup steps.
 *   -t  is the number of host reader threads (the lookup it used to parallelise runs on the
 *       GPU); -t 1 reads every file with the one sequential reader
 *   -b  is the staging block size in bases, as in the reference
 * Environment: CUDA_VISIBLE_DEVICES selects the GPUs, VAFGPU_DEVICES=n how many of them are used (default 1: one GPU
 *              scans hundreds of times faster than the readers parse; 0 = all visible);
 *              VAFGPU_RECIPE=1 runs the literal on-device recipe (verification mode).
 */
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#include "../../include/vafgpu.h"
#include "ingest.h"

typedef struct { /* one line of patterns.txt, vaf-counter.c:92-103 */
	char chr[256];
	int start, end;
	char rsid[256];
	char ref, alt;
	char ref_kmer[128], alt_kmer[128];
} pattern_t;

typedef struct {
	int n, m;
	pattern_t *a;
} pattern_db_t;

static double now(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + tv.tv_usec * 1e-6;
}

/* vaf-counter.c:149-184: eight white-space separated fields, stop at the first bad record */
static pattern_db_t *load_patterns(const char *fn)
{
	FILE *fp = fopen(fn, "r");
	pattern_db_t *db;
	pattern_t p;
	if (!fp) return NULL;
	db = (pattern_db_t *)calloc(1, sizeof(*db));
	while (fscanf(fp, "%255s%d%d%255s %c %c%127s%127s", p.chr, &p.start, &p.end, p.rsid, &p.ref,
	              &p.alt, p.ref_kmer, p.alt_kmer) == 8) {
		if (db->n == db->m) {
			pattern_t *t;
			db->m = db->m ? db->m * 2 : 16;
			t = (pattern_t *)realloc(db->a, (size_t)db->m * sizeof(pattern_t));
			if (!t) break;
			db->a = t;
		}
		db->a[db->n++] = p;
	}
	fclose(fp);
	return db;
}

/* strict base table of vaf-counter.c:73-90, used for pattern k-mers (vaf-counter.c:117-127) */
static int base_code(unsigned char b)
{
	if (b < 4) return b;
	switch (b | 0x20) {
	case 'a': return 0;
	case 'c': return 1;
	case 'g': return 2;
	case 't': case 'u': return 3;
	}
	return -1;
}

/* canonical k-mer of the first k characters, or UINT64_MAX if one is not a base */
static uint64_t canonical_of(const char *s, int k)
{
	uint64_t f = 0, r = 0;
	for (int i = 0; i < k; ++i) {
		int c = base_code((unsigned char)s[i]);
		if (c < 0) return UINT64_MAX;
		f = f << 2 | (uint64_t)c;
		r = r >> 2 | (uint64_t)(3 - c) << 2 * (k - 1);
	}
	return f < r ? f : r;
}

/* The combined map of vaf-counter.c:198-252 as a flat list: canonical(ref) -> i<<1,
 * canonical(alt) -> i<<1|1, in file order, a key that is already present keeps its first
 * value and is counted as a collision.  Returns the number of distinct keys. */
static uint32_t build_key_list(const pattern_db_t *db, int k, uint64_t **keys_out, uint32_t **vals_out,
                               int *n_collisions)
{
	size_t cap = 16;
	uint32_t n = 0;
	while (cap < (size_t)db->n * 4 + 16) cap <<= 1;
	uint64_t *slot = (uint64_t *)malloc(cap * 8);
	uint64_t *keys = (uint64_t *)malloc(((size_t)db->n * 2 + 1) * 8);
	uint32_t *vals = (uint32_t *)malloc(((size_t)db->n * 2 + 1) * 4);
	memset(slot, 0xFF, cap * 8);
	*n_collisions = 0;
	for (int i = 0; i < db->n; ++i)
		for (int alt = 0; alt < 2; ++alt) {
			uint64_t c = canonical_of(alt ? db->a[i].alt_kmer : db->a[i].ref_kmer, k);
			size_t h;
			if (c == UINT64_MAX) continue; /* vaf-counter.c:223,235 */
			h = (size_t)((c * 0x9E3779B97F4A7C15ULL) >> 20) & (cap - 1);
			while (slot[h] != UINT64_MAX && slot[h] != c) h = (h + 1) & (cap - 1);
			if (slot[h] == c) {
				++*n_collisions;
				continue;
			}
			slot[h] = c;
			keys[n] = c;
			vals[n++] = (uint32_t)i << 1 | (uint32_t)alt;
		}
	free(slot);
	*keys_out = keys;
	*vals_out = vals;
	return n;
}

int main(int argc, char *argv[])
{
	int c, k = 21, n_thread = 4, block_size = 10000000, verbose = 0, n_collisions = 0;
	char *pattern_fn = 0, *out_fn = 0;
	double t_start, t0, t_load, t_map, t_count, t_write;
	vafgpu_ctx *ctx = NULL;
	vafgpu_stats st;

	while ((c = getopt(argc, argv, "k:p:o:t:b:v")) >= 0) {
		if (c == 'k') k = atoi(optarg);
		else if (c == 'p') pattern_fn = optarg;
		else if (c == 'o') out_fn = optarg;
		else if (c == 't') n_thread = atoi(optarg);
		else if (c == 'b') block_size = atoi(optarg);
		else if (c == 'v') verbose = 1;
	}
	if (!pattern_fn || !out_fn || argc - optind < 1) {
		fprintf(stderr, "Usage: vaf-counter [options] -p <patterns.txt> -o <output.vaf> <reads.fq> [reads2.fq ...]\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -k INT    k-mer length [%d]\n", k);
		fprintf(stderr, "  -p FILE   input pattern file\n");
		fprintf(stderr, "  -o FILE   output VAF file\n");
		fprintf(stderr, "  -t INT    number of threads [%d]\n", n_thread);
		fprintf(stderr, "  -b INT    block size [%d]\n", block_size);
		fprintf(stderr, "  -v        verbose mode (report performance statistics)\n");
		return 1;
	}
	if (k < 1 || k > 31) { /* the reference's 64-bit words hold at most 31 bases (vaf-counter.c:352) */
		fprintf(stderr, "Error: k must be between 1 and 31\n");
		return 1;
	}
	t_start = now();

	fprintf(stderr, "[M::%s] Loading patterns...\n", __func__);
	t0 = now();
	pattern_db_t *db = load_patterns(pattern_fn);
	if (!db) {
		fprintf(stderr, "Error: failed to load pattern file\n");
		return 1;
	}
	t_load = now() - t0;
	fprintf(stderr, "[M::%s] Loaded %d patterns in %.3f sec\n", __func__, db->n, t_load);

	fprintf(stderr, "[M::%s] Creating k-mer map...\n", __func__);
	t0 = now();
	if (db->n > (INT32_MAX >> 1)) {
		fprintf(stderr, "Error: too many patterns (%d), maximum is %d\n", db->n, INT32_MAX >> 1);
		fprintf(stderr, "Error: failed to create k-mer map\n");
		return 1;
	}
	uint64_t *keys;
	uint32_t *vals;
	uint32_t n_keys = build_key_list(db, k, &keys, &vals, &n_collisions);
	if (n_collisions > 0)
		fprintf(stderr, "[W::%s] Warning: %d k-mer collisions detected. Some patterns may have overlapping k-mers.\n",
		        "create_combined_kmer_map", n_collisions);
	unsigned flags = 0;
	const char *env = getenv("VAFGPU_RECIPE");
	if (env && atoi(env)) flags |= VAFGPU_F_REFERENCE_RECIPE;
	/* staging blocks per device: one for every reader to fill plus two in flight */
	int n_buffers = n_thread > 1 ? n_thread + 2 : 3;
	if (n_buffers > 66) n_buffers = 66;
	/* -b bounds a staging block from above; more than 2 MiB buys nothing here (the readers, not the copies, set the
	 * pace) and page-locking the blocks is start-up time: 18 blocks of 16 / 4 / 2 MiB take 280 / 110 / 30 ms */
	size_t staging = block_size > 0 ? (size_t)block_size : 0;
	if (staging == 0 || staging > ((size_t)2 << 20)) staging = (size_t)2 << 20;
	/* one GPU unless VAFGPU_DEVICES asks for more (0 = all): it scans 300 times faster than -t readers parse, and
	 * every further GPU is a context (and its share of seconds) before the first read */
	int n_gpus = getenv("VAFGPU_DEVICES") ? atoi(getenv("VAFGPU_DEVICES")) : 1;
	if (vafgpu_create(&ctx, k, keys, vals, n_keys, (uint32_t)db->n, staging, n_buffers, n_gpus < 0 ? 1 : n_gpus, flags) != VAFGPU_OK) {
		fprintf(stderr, "Error: failed to create k-mer map: %s\n", vafgpu_strerror(NULL));
		return 1;
	}
	t_map = now() - t0;
	if (verbose)
		fprintf(stderr, "[V::%s] Created k-mer map with %u entries in %.3f sec\n", __func__, n_keys, t_map);

	fprintf(stderr, "[M::%s] Counting k-mers in FASTQ files with %d threads...\n", __func__, n_thread);
	t0 = now();
	/* the reference walks the files one after the other with one reader (vaf-counter.c:647-650);
	 * here -t readers share them, large plain FASTQ files cut into slices (ingest.h) */
	int n_files = argc - optind;
	ingest_file_t *per_file = (ingest_file_t *)calloc((size_t)n_files, sizeof *per_file);
	for (int i = optind; i < argc; ++i) fprintf(stderr, "[M::%s] Processing %s...\n", __func__, argv[i]);
	if (ingest_files(ctx, n_files, argv + optind, k, block_size, n_thread, per_file) != 0) return 1;
	if (verbose)
		for (int i = 0; i < n_files; ++i) {
			const ingest_file_t *f = &per_file[i];
			if (!f->opened) continue;
			fprintf(stderr, "[V::%s] Processed %s: %llu sequences, %llu bases in %.2f sec (%.2f Mbases/sec)\n",
			        "count_fastq_kmers", argv[optind + i], (unsigned long long)f->seqs, (unsigned long long)f->bases, f->seconds,
			        f->seconds > 0 ? f->bases / f->seconds / 1e6 : 0.0);
		}
	uint32_t *counts = (uint32_t *)calloc((size_t)2 * (db->n > 0 ? db->n : 1), 4);
	if (vafgpu_finish(ctx, counts, &st) != VAFGPU_OK) {
		fprintf(stderr, "Error: %s\n", vafgpu_strerror(ctx));
		return 1;
	}
	t_count = now() - t0;

	uint64_t total_ref = 0, total_alt = 0; /* vaf-counter.c:654-658 */
	for (int i = 0; i < db->n; ++i) {
		total_ref += counts[2 * i];
		total_alt += counts[2 * i + 1];
	}
	double avg_depth = (double)(total_ref + total_alt) / (db->n > 0 ? db->n : 1);

	fprintf(stderr, "[M::%s] Writing VAF file...\n", __func__);
	t0 = now();
	FILE *out = fopen(out_fn, "w");
	if (!out) {
		fprintf(stderr, "Error: failed to open output file\n");
		return 1;
	}
	fprintf(out, "# Average depth: %.2f\n", avg_depth); /* vaf-counter.c:668-678 */
	fprintf(out, "CHR\tPOS\tRSID\tREF\tALT\tREF_COUNT\tALT_COUNT\tTOTAL_COUNT\tVAF\n");
	for (int i = 0; i < db->n; ++i) {
		const pattern_t *p = &db->a[i];
		uint32_t r = counts[2 * i], a = counts[2 * i + 1], total = r + a;
		double vaf = total > 0 ? (double)a / total : 0.0;
		fprintf(out, "%s\t%d\t%s\t%c\t%c\t%u\t%u\t%u\t%.4f\n", p->chr, p->start, p->rsid, p->ref, p->alt, r, a, total, vaf);
	}
	fclose(out);
	t_write = now() - t0;
	fprintf(stderr, "[M::%s] Done. Average depth: %.2f\n", __func__, avg_depth);

	if (verbose) { /* vaf-counter.c:686-732, with the GPU's figures where the CPU's were */
		double total_time = now() - t_start;
		fprintf(stderr, "\n=== Performance Statistics ===\n");
		fprintf(stderr, "Total runtime:           %.3f sec\n", total_time);
		fprintf(stderr, "  Pattern loading:       %.3f sec (%.1f%%)\n", t_load, 100.0 * t_load / total_time);
		fprintf(stderr, "  K-mer map creation:    %.3f sec (%.1f%%)\n", t_map, 100.0 * t_map / total_time);
		fprintf(stderr, "  K-mer counting:        %.3f sec (%.1f%%)\n", t_count, 100.0 * t_count / total_time);
		fprintf(stderr, "  Output writing:        %.3f sec (%.1f%%)\n", t_write, 100.0 * t_write / total_time);
		fprintf(stderr, "\nThroughput:\n");
		fprintf(stderr, "  Sequences processed:   %llu\n", (unsigned long long)st.n_reads);
		fprintf(stderr, "  Bases processed:       %llu (%.2f Mbases)\n", (unsigned long long)st.n_bases, st.n_bases / 1e6);
		if (st.n_kmers)
			fprintf(stderr, "  K-mers extracted:      %llu (%.2f million)\n", (unsigned long long)st.n_kmers, st.n_kmers / 1e6);
		if (t_count > 0) fprintf(stderr, "  Speed:                 %.2f Mbases/sec\n", st.n_bases / t_count / 1e6);
		fprintf(stderr, "\nGPU:\n");
		fprintf(stderr, "  Devices:               %d\n", st.n_devices);
		fprintf(stderr, "  Blocks / bytes:        %llu / %llu\n", (unsigned long long)st.n_blocks, (unsigned long long)st.n_bytes);
		fprintf(stderr, "  Copy time (sum):       %.3f ms\n", st.h2d_ms);
		fprintf(stderr, "  Kernel time (sum):     %.3f ms", st.kernel_ms);
		if (st.kernel_ms > 0) fprintf(stderr, " (%.2f Gbases/sec)", st.n_bases / st.kernel_ms / 1e6);
		fprintf(stderr, "\n  Anchor plan:           one %d-mer every %d bases\n", st.anchor_len, st.anchor_stride);
		fprintf(stderr, "  Filter / table:        %u bytes shared / %u slots in L2\n", st.filter_bytes, st.table_slots);
		fprintf(stderr, "  Filter survivors:      %llu\n", (unsigned long long)st.n_candidates);
		fprintf(stderr, "  Pattern k-mer hits:    %llu\n", (unsigned long long)st.n_hits);
		fprintf(stderr, "\nMemory:\n");
		fprintf(stderr, "  Patterns:              %d\n", db->n);
		fprintf(stderr, "  Hash table entries:    %u\n", n_keys);
		fprintf(stderr, "  Threads:               %d readers\n", n_thread);
		fprintf(stderr, "==============================\n");
	}
	t0 = now();
	vafgpu_destroy(ctx);
	if (getenv("VAFGPU_TIMING")) fprintf(stderr, "[vafgpu] %-28s %8.1f ms\n", "destroy", (now() - t0) * 1e3);
	free(counts);
	free(per_file);
	free(keys);
	free(vals);
	free(db->a);
	free(db);
	return 0;
}
