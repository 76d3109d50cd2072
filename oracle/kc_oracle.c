/*
 * kc_oracle.c -- CPU restatement of kc-c4's full k-mer counting path.  TEST INFRASTRUCTURE
 * ONLY (see kc_oracle.h).  Parity pinned against oracle/_ref/kc-c4.
 *
 * The reference keeps 2^p khashl sets keyed by hash64(k-mer) >> p with a 10-bit saturating
 * count in the low bits (kc-c4.c:11-15,116-128).  hash64 is a bijection on 2k-bit words, so
 * what those tables hold is exactly one entry per distinct canonical k-mer; the restatement
 * keeps one growing open-addressing table of (hash64, count) and is independent of p.
 */
#include "kc_oracle.h"

#include <stdlib.h>
#include <string.h>

#include "oracle_reader.h"

#define KCO_MAX 1023u /* kc-c4.c:11-12 */

int kco_nt4(uint8_t b) /* kc-c4.c:21-38 */
{
	switch (b) {
	case 0: case 'A': case 'a': return 0;
	case 1: case 'C': case 'c': return 1;
	case 2: case 'G': case 'g': return 2;
	case 3: case 'T': case 't': case 'U': case 'u': return 3;
	default: return 4;
	}
}

uint64_t kco_hash64(uint64_t key, int k) /* kc-c4.c:40-50 */
{
	const uint64_t m = (1ULL << 2 * k) - 1;
	key = (~key + (key << 21)) & m;
	key ^= key >> 24;
	key = (key * 265) & m;
	key ^= key >> 14;
	key = (key * 21) & m;
	key ^= key >> 28;
	key = (key + (key << 31)) & m;
	return key;
}

long kco_hashed_kmers(const char *seq, long len, int k, uint64_t *out) /* kc-c4.c:74-90 */
{
	const uint64_t m = (1ULL << 2 * k) - 1;
	const int top = 2 * (k - 1);
	uint64_t fw = 0, rv = 0;
	long n = 0;
	int run = 0;
	for (long i = 0; i < len; ++i) {
		int c = kco_nt4((uint8_t)seq[i]);
		if (c > 3) {
			run = 0, fw = rv = 0;
			continue;
		}
		fw = (fw << 2 | (uint64_t)c) & m;
		rv = rv >> 2 | (uint64_t)(3 - c) << top;
		if (++run >= k) out[n++] = kco_hash64(fw < rv ? fw : rv, k);
	}
	return n;
}

struct kco {
	int k;
	uint64_t cap, used, instances; /* cap is a power of two */
	uint64_t *key;                 /* hash64 + 1; 0 = free   */
	uint16_t *cnt;
	uint64_t *scratch;
	long scratch_cap;
};

static uint64_t slot_of(uint64_t h, uint64_t cap) { return (h * 0x9E3779B97F4A7C15ULL) >> 20 & (cap - 1); }

static void grow(kco_t *o)
{
	uint64_t ncap = o->cap * 2;
	uint64_t *nk = (uint64_t *)calloc(ncap, sizeof *nk);
	uint16_t *nc = (uint16_t *)calloc(ncap, sizeof *nc);
	if (!nk || !nc) abort();
	for (uint64_t i = 0; i < o->cap; ++i) {
		if (!o->key[i]) continue;
		uint64_t s = slot_of(o->key[i] - 1, ncap);
		while (nk[s]) s = (s + 1) & (ncap - 1);
		nk[s] = o->key[i];
		nc[s] = o->cnt[i];
	}
	free(o->key);
	free(o->cnt);
	o->key = nk, o->cnt = nc, o->cap = ncap;
}

kco_t *kco_create(int k)
{
	if (k < 1 || k > 31) return NULL;
	kco_t *o = (kco_t *)calloc(1, sizeof *o);
	if (!o) return NULL;
	o->k = k;
	o->cap = 1 << 16;
	o->key = (uint64_t *)calloc(o->cap, sizeof *o->key);
	o->cnt = (uint16_t *)calloc(o->cap, sizeof *o->cnt);
	if (!o->key || !o->cnt) abort();
	return o;
}

void kco_destroy(kco_t *o)
{
	if (!o) return;
	free(o->key);
	free(o->cnt);
	free(o->scratch);
	free(o);
}

void kco_add_hashed(kco_t *o, uint64_t h) /* kc-c4.c:116-128 */
{
	if (o->used * 10 >= o->cap * 6) grow(o);
	uint64_t s = slot_of(h, o->cap);
	while (o->key[s] && o->key[s] != h + 1) s = (s + 1) & (o->cap - 1);
	if (!o->key[s]) {
		o->key[s] = h + 1;
		o->used++;
	}
	if (o->cnt[s] < KCO_MAX) o->cnt[s]++;
	o->instances++;
}

void kco_add_read(kco_t *o, const char *seq, long len)
{
	if (len < o->k) return; /* kc-c4.c:141 */
	if (o->scratch_cap < len) {
		o->scratch_cap = len + 1024;
		o->scratch = (uint64_t *)realloc(o->scratch, (size_t)o->scratch_cap * sizeof *o->scratch);
		if (!o->scratch) abort();
	}
	long n = kco_hashed_kmers(seq, len, o->k, o->scratch);
	for (long i = 0; i < n; ++i) kco_add_hashed(o, o->scratch[i]);
}

/* kc-c4.c:133-183 step 0: a block is read until it holds block_len bases or the reader
 * reports anything below zero -- end of input, or a FASTQ record whose quality does not fit
 * (kseq.h:230).  A bad record in the middle of a block only closes the block: the next one goes
 * on behind it.  An empty block retires the pipeline worker that read it (kthread.c:97-125: a
 * worker leaves when ITS step 0 returns NULL) while the other workers of kt_pipeline(3, ...)
 * (kc-c4.c:173) go on calling step 0, in order: the file ends with the third empty block. */
int kco_add_file(kco_t *o, const char *fn, long block_len)
{
	orr_t *fx = orr_open(fn);
	const char *seq;
	long len;
	if (!fx) return -1;
	int lives = 3;
	for (;;) {
		long sum_len = 0;
		while ((len = orr_next(fx, &seq)) >= 0) {
			if (len < o->k) continue;
			kco_add_read(o, seq, len);
			sum_len += len;
			if (sum_len >= block_len) break;
		}
		if (sum_len == 0 && --lives == 0) break;
	}
	orr_close(fx);
	return 0;
}

void kco_hist(const kco_t *o, uint64_t hist[256]) /* kc-c4.c:186-215 */
{
	memset(hist, 0, 256 * sizeof hist[0]);
	for (uint64_t i = 0; i < o->cap; ++i)
		if (o->key[i]) hist[o->cnt[i] < 255 ? o->cnt[i] : 255]++;
}

uint64_t kco_distinct(const kco_t *o) { return o->used; }
uint64_t kco_instances(const kco_t *o) { return o->instances; }

void kco_print_hist(const uint64_t hist[256], FILE *fp) /* kc-c4.c:232-233 */
{
	for (int i = 1; i < 256; ++i) fprintf(fp, "%d\t%ld\n", i, (long)hist[i]);
}
