/*
 * spg_oracle.c -- CPU restatement of snp-pattern-gen (snp-pattern-gen.c:70-216,219-366).
 * TEST INFRASTRUCTURE ONLY (see vaf_oracle.h for the rules).  Parity status: PINNED --
 * tests/test_spg.py compares its output byte for byte with tests/golden/spg/ (written by the
 * unmodified reference binary, tests/golden/make_spg_golden.py) and with the live
 * oracle/_ref/snp-pattern-gen where that exists.
 *
 *   spg_oracle -k K -b snps.bed -f ref.fa -o patterns.txt
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "oracle_reader.h"
#include "vaf_oracle.h"

typedef struct {
	char *name, *seq;
	long len;
} rec_t;

static rec_t *g_rec;
static int g_n;

static rec_t *find(const char *chr) /* snp-pattern-gen.c:118-126 */
{
	for (int i = 0; i < g_n; ++i)
		if (!strcmp(g_rec[i].name, chr)) return &g_rec[i];
	return NULL;
}

/* canonical code of the k characters around pos with `alt` in the middle (alt = 0: as it
 * stands); UINT64_MAX when the window leaves the contig or holds a non-base
 * (snp-pattern-gen.c:193-216 + 129-159) */
static uint64_t window(const rec_t *r, int pos, int k, char alt, char *text)
{
	long start = (long)pos - k / 2;
	if (start < 0 || start + k > r->len) return VO_NO_KMER;
	for (int i = 0; i < k; ++i)
		if (vo_nt4_strict((uint8_t)r->seq[start + i]) > 3) return VO_NO_KMER;
	memcpy(text, r->seq + start, (size_t)k);
	text[k] = 0;
	if (alt) text[k / 2] = alt;
	uint64_t x = vo_encode_kmer(text, k);
	return x == VO_NO_KMER ? x : vo_canonical(x, k);
}

/* candidate set: sorted array + counts */
static uint64_t *g_key;
static uint32_t *g_cnt;
static size_t g_nk, g_mk;

static int cmp64(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
	return x < y ? -1 : x > y;
}

static long lookup(uint64_t key)
{
	size_t lo = 0, hi = g_nk;
	while (lo < hi) {
		size_t mid = (lo + hi) / 2;
		if (g_key[mid] < key) lo = mid + 1;
		else hi = mid;
	}
	return lo < g_nk && g_key[lo] == key ? (long)lo : -1;
}

int main(int argc, char **argv)
{
	int c, k = 21;
	const char *bed = NULL, *fa = NULL, *out = NULL;
	while ((c = getopt(argc, argv, "k:b:f:o:")) >= 0) {
		if (c == 'k') k = atoi(optarg);
		else if (c == 'b') bed = optarg;
		else if (c == 'f') fa = optarg;
		else if (c == 'o') out = optarg;
	}
	if (k % 2 == 0 || k < 1 || k > 31 || !bed || !fa || !out) return 1;
	orr_t *fx = orr_open(fa);
	if (!fx) return 1;
	const char *s;
	long l;
	while ((l = orr_next(fx, &s)) >= 0) { /* snp-pattern-gen.c:70-104 */
		g_rec = (rec_t *)realloc(g_rec, (size_t)(g_n + 1) * sizeof *g_rec);
		g_rec[g_n].name = strdup(orr_name(fx));
		g_rec[g_n].seq = (char *)malloc((size_t)l + 1);
		memcpy(g_rec[g_n].seq, s, (size_t)l);
		g_rec[g_n].seq[l] = 0;
		g_rec[g_n++].len = l;
	}
	orr_close(fx);

	char chr[256], rsid[256], ref, alt, rk[64], ak[64];
	int start, end;
	FILE *fp = fopen(bed, "r");
	if (!fp) return 1;
	while (fscanf(fp, "%254s%d%d%254s %c %c", chr, &start, &end, rsid, &ref, &alt) == 6) { /* pass 1, :262-301 */
		rec_t *r = find(chr);
		if (!r) continue;
		uint64_t a = window(r, start, k, 0, rk), b = window(r, start, k, alt, ak);
		if (a == VO_NO_KMER || b == VO_NO_KMER) continue;
		if (g_nk + 2 > g_mk) g_key = (uint64_t *)realloc(g_key, (g_mk = g_mk ? g_mk * 2 : 1024) * 8);
		g_key[g_nk++] = a;
		g_key[g_nk++] = b;
	}
	fclose(fp);
	if (g_nk) {
		qsort(g_key, g_nk, 8, cmp64);
		size_t w = 1;
		for (size_t i = 1; i < g_nk; ++i)
			if (g_key[i] != g_key[w - 1]) g_key[w++] = g_key[i];
		g_nk = w;
	}
	g_cnt = (uint32_t *)calloc(g_nk + 1, 4);

	for (int i = 0; i < g_n; ++i) { /* pass 2, :162-190 */
		const uint64_t mask = (1ULL << 2 * k) - 1;
		uint64_t fw = 0, rv = 0;
		int run = 0;
		for (long j = 0; j < g_rec[i].len; ++j) {
			int code = vo_nt4_strict((uint8_t)g_rec[i].seq[j]);
			if (code > 3) {
				run = 0, fw = rv = 0;
				continue;
			}
			fw = (fw << 2 | (uint64_t)code) & mask;
			rv = rv >> 2 | (uint64_t)(3 - code) << 2 * (k - 1);
			if (++run >= k) {
				long at = lookup(fw < rv ? fw : rv);
				if (at >= 0) g_cnt[at]++;
			}
		}
	}

	fp = fopen(bed, "r");
	FILE *op = fopen(out, "w");
	if (!fp || !op) return 1;
	while (fscanf(fp, "%254s%d%d%254s %c %c", chr, &start, &end, rsid, &ref, &alt) == 6) { /* pass 3, :318-356 */
		rec_t *r = find(chr);
		if (!r) continue;
		uint64_t a = window(r, start, k, 0, rk), b = window(r, start, k, alt, ak);
		if (a == VO_NO_KMER || b == VO_NO_KMER) continue;
		long ia = lookup(a), ib = lookup(b);
		if (ia >= 0 && g_cnt[ia] == 1 && ib >= 0 && g_cnt[ib] == 0)
			fprintf(op, "%s\t%d\t%d\t%s\t%c\t%c\t%s\t%s\n", chr, start, end, rsid, ref, alt, rk, ak);
	}
	fclose(fp);
	fclose(op);
	return 0;
}
