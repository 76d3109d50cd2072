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

/* the genome in memory: every record of the FASTA file, named by its header up to the first
 * white space (what snp-pattern-gen.c:70-104 keeps) */
typedef struct {
	char *name, *bases;
	long n_bases;
} contig_t;

typedef struct {
	contig_t *contig;
	int n_contigs, room;
} genome_t;

static genome_t *genome_load(const char *fn)
{
	fastx_t *fx = fastx_open(fn);
	if (!fx) return NULL;
	genome_t *g = (genome_t *)calloc(1, sizeof *g);
	const char *s;
	for (long l; (l = fastx_next(fx, &s)) >= 0;) {
		if (g->n_contigs == g->room) {
			g->room = g->room ? g->room * 2 : 16;
			g->contig = (contig_t *)realloc(g->contig, (size_t)g->room * sizeof *g->contig);
		}
		contig_t *c = &g->contig[g->n_contigs++];
		c->name = strdup(fastx_name(fx));
		c->bases = (char *)malloc((size_t)l + 1);
		memcpy(c->bases, s, (size_t)l);
		c->bases[l] = 0;
		c->n_bases = l;
	}
	fastx_close(fx);
	return g;
}

/* the first contig of that name (snp-pattern-gen.c:118-126), NULL if there is none */
static const contig_t *genome_find(const genome_t *g, const char *name)
{
	for (int i = 0; i < g->n_contigs; ++i)
		if (strcmp(g->contig[i].name, name) == 0) return &g->contig[i];
	return NULL;
}

static int base_code(unsigned char b) /* the strict table, snp-pattern-gen.c:30-47 */
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

/* canonical k-mer of k characters that are all bases (snp-pattern-gen.c:129-159) */
static uint64_t canonical_of(const char *s, int k)
{
	uint64_t f = 0, r = 0;
	for (int i = 0; i < k; ++i) {
		const uint64_t c = (uint64_t)base_code((unsigned char)s[i]);
		f = f << 2 | c;
		r = r >> 2 | (3 - c) << 2 * (k - 1);
	}
	return f < r ? f : r;
}

/* One row of the BED file (snp-pattern-gen.c:60-67) with everything both host passes need of it,
 * worked out once: where it lies, the k characters around it as they stand in the genome
 * (snp-pattern-gen.c:193-216: the window must lie inside the contig and hold bases only) with
 * the reference and with the alternative allele in the middle, and their canonical words. */
typedef struct {
	char chr[256], rsid[256];
	int start, end;
	char ref, alt;
	int on_genome;   /* its contig exists */
	int usable;      /* ... and the window is whole and free of non-bases, the alternative allele a base */
	char ref_kmer[32], alt_kmer[32];
	uint64_t ref_key, alt_key;
} site_t;

static void site_resolve(site_t *s, const genome_t *g, int k)
{
	const contig_t *c = genome_find(g, s->chr);
	const long first = (long)s->start - k / 2;
	s->on_genome = c != NULL;
	s->usable = 0;
	if (!c || first < 0 || first + k > c->n_bases) return;
	for (int i = 0; i < k; ++i)
		if (base_code((unsigned char)c->bases[first + i]) < 0) return;
	if (base_code((unsigned char)s->alt) < 0) return; /* the reference finds this out when it encodes the k-mer (:283,335) */
	memcpy(s->ref_kmer, c->bases + first, (size_t)k);
	memcpy(s->alt_kmer, c->bases + first, (size_t)k);
	s->alt_kmer[k / 2] = s->alt;
	s->ref_kmer[k] = s->alt_kmer[k] = 0;
	s->ref_key = canonical_of(s->ref_kmer, k);
	s->alt_key = canonical_of(s->alt_kmer, k);
	s->usable = 1;
}

/* every row of the BED file the reference's fscanf loop would read (snp-pattern-gen.c:270,327) */
static site_t *sites_load(const char *fn, const genome_t *g, int k, int *n_out)
{
	FILE *fp = fopen(fn, "r");
	site_t *sites = NULL, row;
	int n = 0, room = 0;
	if (!fp) return NULL;
	memset(&row, 0, sizeof row);
	while (fscanf(fp, "%254s%d%d%254s %c %c", row.chr, &row.start, &row.end, row.rsid, &row.ref, &row.alt) == 6) {
		if (n == room) {
			room = room ? room * 2 : 1024;
			sites = (site_t *)realloc(sites, (size_t)room * sizeof *sites);
		}
		site_resolve(&row, g, k);
		sites[n++] = row;
	}
	fclose(fp);
	if (!sites) sites = (site_t *)calloc(1, sizeof *sites);
	*n_out = n;
	return sites;
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
	const genome_t *genome;
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
		int i = f->failed ? f->genome->n_contigs : f->next++;
		pthread_mutex_unlock(&f->mu);
		if (i >= f->genome->n_contigs) break;
		const contig_t *q = &f->genome->contig[f->order[i]];
		if (vafgpu_producer_add_read(p, q->bases, (size_t)q->n_bases) != VAFGPU_OK) f->failed = 1;
	}
	if (vafgpu_producer_destroy(p) != VAFGPU_OK) f->failed = 1;
	return NULL;
}

static const genome_t *g_sort_genome;
static int longest_first(const void *a, const void *b)
{
	long la = g_sort_genome->contig[*(const int *)a].n_bases, lb = g_sort_genome->contig[*(const int *)b].n_bases;
	return la < lb ? 1 : la > lb ? -1 : *(const int *)a - *(const int *)b;
}

int main(int argc, char *argv[])
{
	int c, k = 21;
	const char *bed_fn = NULL, *fasta_fn = NULL, *out_fn = NULL;
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
	genome_t *genome = genome_load(fasta_fn);
	if (!genome) {
		fprintf(stderr, "Error: failed to load FASTA file\n");
		return 1;
	}
	fprintf(stderr, "[M::%s] Loaded %d sequences\n", __func__, genome->n_contigs);

	/* the BED rows, read once and resolved against the genome; the reference reads the file twice
	 * (snp-pattern-gen.c:262-301 and :318-356) and resolves every row in both passes */
	fprintf(stderr, "[M::%s] Generating candidate k-mers from BED file...\n", __func__);
	int n_sites = 0;
	site_t *sites = sites_load(bed_fn, genome, k, &n_sites);
	if (!sites) {
		fprintf(stderr, "Error: failed to open BED file\n");
		return 1;
	}
	/* the candidate set: both k-mers of every usable row */
	cand_t cand;
	memset(&cand, 0, sizeof cand);
	int n_candidate = 0;
	for (int i = 0; i < n_sites; ++i)
		if (sites[i].usable) n_candidate += cand_put(&cand, sites[i].ref_key) + cand_put(&cand, sites[i].alt_key);
	fprintf(stderr, "[M::%s] Generated %d candidate k-mers\n", __func__, n_candidate);

	/* the genome scan on the GPU: candidate i is "allele i & 1 of pattern i >> 1" of a vaf-counter panel */
	fprintf(stderr, "[M::%s] Counting candidate k-mers in genome...\n", __func__);
	uint32_t *vals = (uint32_t *)malloc(((size_t)cand.n + 1) * 4);
	uint32_t n_pairs = (cand.n + 1) / 2 ? (cand.n + 1) / 2 : 1;
	uint32_t *counts = (uint32_t *)calloc((size_t)2 * n_pairs, 4);
	for (uint32_t i = 0; i < cand.n; ++i) vals[i] = i;
	vafgpu_ctx *ctx = NULL;
	int n_feed = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if (n_feed > 8) n_feed = 8;
	if (n_feed > genome->n_contigs) n_feed = genome->n_contigs;
	if (n_feed < 1) n_feed = 1;
	/* one GPU: a human genome is a second of kernel time, more devices only add start-up */
	if (vafgpu_create(&ctx, k, cand.keys, vals, cand.n, n_pairs, 0, n_feed + 2, 1, VAFGPU_F_STRICT_BYTES) != VAFGPU_OK) {
		fprintf(stderr, "Error: %s\n", vafgpu_strerror(NULL));
		return 1;
	}
	feed_t feed;
	memset(&feed, 0, sizeof feed);
	feed.ctx = ctx, feed.genome = genome;
	feed.order = (int *)malloc(((size_t)genome->n_contigs + 1) * sizeof(int));
	for (int i = 0; i < genome->n_contigs; ++i) feed.order[i] = i;
	g_sort_genome = genome;
	qsort(feed.order, (size_t)genome->n_contigs, sizeof(int), longest_first);
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

	/* the selection: rows whose reference k-mer occurs exactly once in the genome and whose
	 * alternative k-mer never does (snp-pattern-gen.c:1-16,338-356), in the order of the BED file */
	FILE *out_fp = fopen(out_fn, "w");
	if (!out_fp) {
		fprintf(stderr, "Error: failed to open output file\n");
		return 1;
	}
	fprintf(stderr, "[M::%s] Processing SNPs...\n", __func__);
	int n_unique = 0;
	for (int i = 0; i < n_sites; ++i) {
		const site_t *s = &sites[i];
		if (!s->on_genome) {
			fprintf(stderr, "Warning: chromosome %s not found\n", s->chr);
			continue;
		}
		if (!s->usable) continue;
		const long ri = cand_get(&cand, s->ref_key), ai = cand_get(&cand, s->alt_key);
		if (ri < 0 || ai < 0 || counts[ri] != 1 || counts[ai] != 0) continue;
		fprintf(out_fp, "%s\t%d\t%d\t%s\t%c\t%c\t%s\t%s\n", s->chr, s->start, s->end, s->rsid, s->ref, s->alt, s->ref_kmer, s->alt_kmer);
		++n_unique;
	}
	fprintf(stderr, "[M::%s] Total SNPs: %d, Unique k-mer pairs: %d\n", __func__, n_sites, n_unique);
	fclose(out_fp);
	return 0;
}
