/*
 * snp_pattern_gen.c -- the snp-pattern-gen command line with its genome scan on the GPU.
 *
 * Same options, inputs, messages and output as the reference tool (snp-pattern-gen.c:219-366):
 *   snp-pattern-gen -k INT -b snps.bed -f ref.fa -o patterns.txt
 * A SNP is kept when its reference k-mer occurs exactly once in the genome and its alternative
 * k-mer never (both strands).  The host keeps what is cheap and sequential: the FASTA in memory
 * (snp-pattern-gen.c:70-104), the candidate k-mers of the BED rows (pass 1, :262-301) and the
 * selection (pass 3, :318-356).  Pass 2 (count_candidate_kmers, :162-190: every canonical
 * k-mer of the genome looked up in the candidate set, minutes for a human genome) is the
 * vaf-counter hot path with the candidates as the panel, so it runs on the same engine:
 * include/vafgpu.h with VAFGPU_F_STRICT_BYTES (the scan uses the strict base table everywhere).
 * Chromosomes are handed over by a few reader threads, one producer each.
 *
 * Deviation: k outside 1..31 is rejected (the reference shifts by >= 64 bits from k = 32 on).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/vafgpu.h"
#include "fastx.h"

typedef struct {
	char *name, *seq;
	long len;
} fasta_seq_t;

typedef struct {
	int n, m;
	fasta_seq_t *a;
} fasta_db_t;

typedef struct { /* one BED row, snp-pattern-gen.c:60-67 */
	char chr[256];
	int start, end;
	char rsid[256];
	char ref, alt;
} snp_t;

/* snp-pattern-gen.c:70-104: every record of the file, name up to the first white space */
static fasta_db_t *load_fasta(const char *fn)
{
	fastx_t *fx = fastx_open(fn);
	fasta_db_t *db;
	const char *s;
	long l;
	if (!fx) return NULL;
	db = (fasta_db_t *)calloc(1, sizeof *db);
	while ((l = fastx_next(fx, &s)) >= 0) {
		if (db->n == db->m) {
			db->m = db->m ? db->m << 1 : 16;
			db->a = (fasta_seq_t *)realloc(db->a, (size_t)db->m * sizeof *db->a);
		}
		fasta_seq_t *q = &db->a[db->n++];
		q->name = strdup(fastx_name(fx));
		q->seq = (char *)malloc((size_t)l + 1);
		memcpy(q->seq, s, (size_t)l);
		q->seq[l] = 0;
		q->len = l;
	}
	fastx_close(fx);
	return db;
}

static fasta_seq_t *find_seq(fasta_db_t *db, const char *chr) /* snp-pattern-gen.c:118-126: the first of that name */
{
	for (int i = 0; i < db->n; ++i)
		if (strcmp(db->a[i].name, chr) == 0) return &db->a[i];
	return NULL;
}

static int base_code(unsigned char b) /* snp-pattern-gen.c:30-47 */
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

/* canonical k-mer of k characters, UINT64_MAX if one is not a base (snp-pattern-gen.c:129-159) */
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

/* snp-pattern-gen.c:193-216: the k characters around the SNP, as they stand in the genome */
static int extract_snp_kmer(const fasta_seq_t *seq, int pos, char alt, int k, char *ref_kmer, char *alt_kmer)
{
	const int flank = k / 2;
	const long start = (long)pos - flank;
	if (start < 0 || start + k > seq->len) return 0;
	for (int i = 0; i < k; ++i)
		if (base_code((unsigned char)seq->seq[start + i]) < 0) return 0;
	memcpy(ref_kmer, seq->seq + start, (size_t)k);
	ref_kmer[k] = 0;
	memcpy(alt_kmer, seq->seq + start, (size_t)k);
	alt_kmer[flank] = alt;
	alt_kmer[k] = 0;
	return 1;
}

/* the candidate set: canonical k-mer -> index, in order of first appearance */
typedef struct {
	uint64_t *slot_key;
	uint32_t *slot_idx;
	size_t cap;
	uint64_t *keys;
	uint32_t n, m;
} cand_t;

static size_t cand_slot(const cand_t *c, uint64_t key)
{
	size_t h = (size_t)((key * 0x9E3779B97F4A7C15ULL) >> 20) & (c->cap - 1);
	while (c->slot_key[h] != UINT64_MAX && c->slot_key[h] != key) h = (h + 1) & (c->cap - 1);
	return h;
}

static void cand_grow(cand_t *c)
{
	size_t cap = c->cap ? c->cap * 2 : 1024;
	free(c->slot_key);
	free(c->slot_idx);
	c->slot_key = (uint64_t *)malloc(cap * 8);
	c->slot_idx = (uint32_t *)malloc(cap * 4);
	memset(c->slot_key, 0xFF, cap * 8);
	c->cap = cap;
	for (uint32_t i = 0; i < c->n; ++i) {
		size_t h = cand_slot(c, c->keys[i]);
		c->slot_key[h] = c->keys[i];
		c->slot_idx[h] = i;
	}
}

static int cand_put(cand_t *c, uint64_t key) /* 1 if it was absent */
{
	if ((size_t)(c->n + 1) * 2 > c->cap) cand_grow(c);
	size_t h = cand_slot(c, key);
	if (c->slot_key[h] == key) return 0;
	if (c->n == c->m) {
		c->m = c->m ? c->m * 2 : 1024;
		c->keys = (uint64_t *)realloc(c->keys, (size_t)c->m * 8);
	}
	c->slot_key[h] = key;
	c->slot_idx[h] = c->n;
	c->keys[c->n++] = key;
	return 1;
}

static long cand_get(const cand_t *c, uint64_t key) /* index or -1 */
{
	if (!c->cap) return -1;
	size_t h = cand_slot(c, key);
	return c->slot_key[h] == key ? (long)c->slot_idx[h] : -1;
}

/* pass 2: the chromosomes to the engine, longest first, a few reader threads */
typedef struct {
	vafgpu_ctx *ctx;
	fasta_db_t *db;
	int *order, next, failed;
	pthread_mutex_t mu;
} feed_t;

static void *feeder(void *arg)
{
	feed_t *f = (feed_t *)arg;
	vafgpu_producer *p = NULL;
	if (vafgpu_producer_create(f->ctx, &p) != VAFGPU_OK) {
		f->failed = 1;
		return NULL;
	}
	for (;;) {
		pthread_mutex_lock(&f->mu);
		int i = f->failed ? f->db->n : f->next++;
		pthread_mutex_unlock(&f->mu);
		if (i >= f->db->n) break;
		const fasta_seq_t *q = &f->db->a[f->order[i]];
		if (vafgpu_producer_add_read(p, q->seq, (size_t)q->len) != VAFGPU_OK) f->failed = 1;
	}
	if (vafgpu_producer_destroy(p) != VAFGPU_OK) f->failed = 1;
	return NULL;
}

static fasta_db_t *g_sort_db;
static int by_len_desc(const void *a, const void *b)
{
	long la = g_sort_db->a[*(const int *)a].len, lb = g_sort_db->a[*(const int *)b].len;
	return la < lb ? 1 : la > lb ? -1 : *(const int *)a - *(const int *)b;
}

int main(int argc, char *argv[])
{
	int c, k = 21;
	char *bed_fn = 0, *fasta_fn = 0, *out_fn = 0;
	FILE *bed_fp, *out_fp;
	snp_t snp;
	char ref_kmer[128], alt_kmer[128];
	int n_total = 0, n_unique = 0, n_candidate = 0;
	cand_t cand;
	memset(&cand, 0, sizeof cand);

	while ((c = getopt(argc, argv, "k:b:f:o:")) >= 0) {
		if (c == 'k') k = atoi(optarg);
		else if (c == 'b') bed_fn = optarg;
		else if (c == 'f') fasta_fn = optarg;
		else if (c == 'o') out_fn = optarg;
	}
	if (k % 2 == 0) { /* snp-pattern-gen.c:238-241 */
		fprintf(stderr, "Error: k must be odd\n");
		return 1;
	}
	if (!bed_fn || !fasta_fn || !out_fn) {
		fprintf(stderr, "Usage: snp-pattern-gen -k %d -b <snps.bed> -f <ref.fa> -o <patterns.txt>\n", k);
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -k INT    k-mer length (must be odd) [%d]\n", k);
		fprintf(stderr, "  -b FILE   input BED file with SNPs\n");
		fprintf(stderr, "  -f FILE   input reference genome FASTA file\n");
		fprintf(stderr, "  -o FILE   output pattern file\n");
		return 1;
	}
	if (k < 1 || k > 31) {
		fprintf(stderr, "Error: k must be between 1 and 31\n");
		return 1;
	}

	fprintf(stderr, "[M::%s] Loading reference genome...\n", __func__);
	fasta_db_t *db = load_fasta(fasta_fn);
	if (!db) {
		fprintf(stderr, "Error: failed to load FASTA file\n");
		return 1;
	}
	fprintf(stderr, "[M::%s] Loaded %d sequences\n", __func__, db->n);

	/* pass 1: candidate k-mers of the BED rows (snp-pattern-gen.c:262-301) */
	fprintf(stderr, "[M::%s] Generating candidate k-mers from BED file...\n", __func__);
	bed_fp = fopen(bed_fn, "r");
	if (!bed_fp) {
		fprintf(stderr, "Error: failed to open BED file\n");
		return 1;
	}
	while (fscanf(bed_fp, "%254s%d%d%254s %c %c", snp.chr, &snp.start, &snp.end, snp.rsid, &snp.ref, &snp.alt) == 6) {
		fasta_seq_t *seq = find_seq(db, snp.chr);
		if (!seq) continue;
		if (extract_snp_kmer(seq, snp.start, snp.alt, k, ref_kmer, alt_kmer)) {
			uint64_t ref_can = canonical_of(ref_kmer, k), alt_can = canonical_of(alt_kmer, k);
			if (ref_can == UINT64_MAX || alt_can == UINT64_MAX) continue;
			n_candidate += cand_put(&cand, ref_can);
			n_candidate += cand_put(&cand, alt_can);
		}
	}
	fclose(bed_fp);
	fprintf(stderr, "[M::%s] Generated %d candidate k-mers\n", __func__, n_candidate);

	/* pass 2 on the GPU: candidate i is "allele i & 1 of pattern i >> 1" of a vaf-counter panel */
	fprintf(stderr, "[M::%s] Counting candidate k-mers in genome...\n", __func__);
	uint32_t *vals = (uint32_t *)malloc(((size_t)cand.n + 1) * 4);
	uint32_t n_pairs = (cand.n + 1) / 2 ? (cand.n + 1) / 2 : 1;
	uint32_t *counts = (uint32_t *)calloc((size_t)2 * n_pairs, 4);
	for (uint32_t i = 0; i < cand.n; ++i) vals[i] = i;
	vafgpu_ctx *ctx = NULL;
	int n_feed = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (n_feed > 8) n_feed = 8;
	if (n_feed > db->n) n_feed = db->n;
	if (n_feed < 1) n_feed = 1;
	/* one GPU: a human genome is a second of kernel time, more devices only add start-up */
	if (vafgpu_create(&ctx, k, cand.keys, vals, cand.n, n_pairs, 0, n_feed + 2, 1, VAFGPU_F_STRICT_BYTES) != VAFGPU_OK) {
		fprintf(stderr, "Error: %s\n", vafgpu_strerror(NULL));
		return 1;
	}
	feed_t feed;
	memset(&feed, 0, sizeof feed);
	feed.ctx = ctx, feed.db = db;
	feed.order = (int *)malloc(((size_t)db->n + 1) * sizeof(int));
	for (int i = 0; i < db->n; ++i) feed.order[i] = i;
	g_sort_db = db;
	qsort(feed.order, (size_t)db->n, sizeof(int), by_len_desc);
	pthread_mutex_init(&feed.mu, NULL);
	pthread_t th[8];
	for (int i = 1; i < n_feed; ++i) pthread_create(&th[i], NULL, feeder, &feed);
	feeder(&feed);
	for (int i = 1; i < n_feed; ++i) pthread_join(th[i], NULL);
	if (feed.failed || vafgpu_finish(ctx, counts, NULL) != VAFGPU_OK) {
		fprintf(stderr, "Error: %s\n", vafgpu_strerror(ctx));
		return 1;
	}
	vafgpu_destroy(ctx);
	fprintf(stderr, "[M::%s] Finished counting k-mers\n", __func__);

	/* pass 3: the rows whose pair is unique (snp-pattern-gen.c:303-356) */
	bed_fp = fopen(bed_fn, "r");
	if (!bed_fp) {
		fprintf(stderr, "Error: failed to open BED file\n");
		return 1;
	}
	out_fp = fopen(out_fn, "w");
	if (!out_fp) {
		fprintf(stderr, "Error: failed to open output file\n");
		return 1;
	}
	fprintf(stderr, "[M::%s] Processing SNPs...\n", __func__);
	while (fscanf(bed_fp, "%254s%d%d%254s %c %c", snp.chr, &snp.start, &snp.end, snp.rsid, &snp.ref, &snp.alt) == 6) {
		fasta_seq_t *seq = find_seq(db, snp.chr);
		++n_total;
		if (!seq) {
			fprintf(stderr, "Warning: chromosome %s not found\n", snp.chr);
			continue;
		}
		if (extract_snp_kmer(seq, snp.start, snp.alt, k, ref_kmer, alt_kmer)) {
			uint64_t ref_can = canonical_of(ref_kmer, k), alt_can = canonical_of(alt_kmer, k);
			if (ref_can == UINT64_MAX || alt_can == UINT64_MAX) continue;
			long ri = cand_get(&cand, ref_can), ai = cand_get(&cand, alt_can);
			if (ri >= 0 && counts[ri] == 1 && ai >= 0 && counts[ai] == 0) {
				fprintf(out_fp, "%s\t%d\t%d\t%s\t%c\t%c\t%s\t%s\n", snp.chr, snp.start, snp.end, snp.rsid, snp.ref, snp.alt,
				        ref_kmer, alt_kmer);
				++n_unique;
			}
		}
	}
	fprintf(stderr, "[M::%s] Total SNPs: %d, Unique k-mer pairs: %d\n", __func__, n_total, n_unique);
	fclose(bed_fp);
	fclose(out_fp);
	return 0;
}
