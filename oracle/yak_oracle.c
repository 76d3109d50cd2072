/*
 * yak_oracle.c -- CPU restatement of yak-count.  TEST INFRASTRUCTURE ONLY (see yak_oracle.h).
 * Parity pinned against oracle/_ref/yak-count.
 *
 * The reference keeps 2^p khashl sets keyed by hash64(k-mer) >> p with a 10-bit saturating
 * count in the low bits, and, with -b, one blocked Bloom filter of 2^(b-p) bits beside each
 * (yak-count.c:106-124).  hash64 is a bijection, so (partition, key) names a k-mer; the
 * restatement keeps ONE growing table over the full hash and the 2^p Bloom filters as they are.
 * K-mers are handed over in stream order: each partition then sees its own in the order the
 * reference's per-partition buffers hold them (yak-count.c:330-341,363-369), which is what
 * decides the Bloom filter's false positives.
 */
#include "yak_oracle.h"

#include <stdlib.h>
#include <string.h>

#include "kc_oracle.h" /* the byte table and hash64 are the same as kc-c4's (yak-count.c:47-57,291-308) */
#include "oracle_reader.h"

#define YKO_MAX 1023u   /* yak-count.c:9-12 */
#define YKO_BLK_SHIFT 9 /* a Bloom block is 512 bits (yak-count.c:14-15) */

struct yko {
	int k, pre, bf_shift, n_hash;
	int part_shift;     /* log2 of one partition's Bloom filter in bits, 0 = no filter */
	uint8_t **bloom;    /* one per partition */
	uint64_t cap, used; /* cap is a power of two */
	uint64_t *key;      /* hash64 + 1; 0 = free */
	uint16_t *cnt;
	uint64_t *scratch;
	long scratch_cap;
};

int yko_bf_insert(uint8_t *bits, int n_shift, int n_hash, uint64_t hash) /* yak-count.c:86-104 */
{
	const int x = n_shift - YKO_BLK_SHIFT;
	uint8_t *block = bits + ((hash & ((1ULL << x) - 1)) << (YKO_BLK_SHIFT - 3));
	int z = (int)(hash >> x & 511), step = (int)(hash >> n_shift & 511), seen = 0;
	if ((step & 31) == 0) step = (step + 1) & 511;
	for (int i = 0; i < n_hash; ++i, z = (z + step) & 511) {
		const uint8_t bit = (uint8_t)(1u << (z & 7));
		seen += (block[z >> 3] & bit) != 0;
		block[z >> 3] |= bit;
	}
	return seen;
}

yko_t *yko_create(int k, int pre, int bf_shift, int bf_n_hash)
{
	if (k < 1 || k > 31 || pre < 10 || pre > 30) return NULL; /* yak-count.c:110,492-495 */
	yko_t *o = (yko_t *)calloc(1, sizeof *o);
	o->k = k, o->pre = pre, o->bf_shift = bf_shift, o->n_hash = bf_n_hash;
	/* a filter per partition if -H > 0 and -b > -p (yak-count.c:117-121), and if its size is one
	 * the filter accepts: at least one block, at most 2^55 bits (yak-count.c:75) */
	if (bf_n_hash > 0 && bf_shift > pre && bf_shift - pre >= YKO_BLK_SHIFT && bf_shift - pre + YKO_BLK_SHIFT <= 64) {
		o->part_shift = bf_shift - pre;
		o->bloom = (uint8_t **)calloc((size_t)1 << pre, sizeof *o->bloom);
		for (long i = 0; i < 1L << pre; ++i) o->bloom[i] = (uint8_t *)calloc((size_t)1 << (o->part_shift - 3), 1);
	}
	o->cap = 1 << 16;
	o->key = (uint64_t *)calloc(o->cap, sizeof *o->key);
	o->cnt = (uint16_t *)calloc(o->cap, sizeof *o->cnt);
	return o;
}

static void drop_bloom(yko_t *o) /* yak-count.c:126-135 */
{
	if (!o->bloom) return;
	for (long i = 0; i < 1L << o->pre; ++i) free(o->bloom[i]);
	free(o->bloom);
	o->bloom = NULL;
}

void yko_destroy(yko_t *o)
{
	if (!o) return;
	drop_bloom(o);
	free(o->key), free(o->cnt), free(o->scratch), free(o);
}

static uint64_t slot_of(uint64_t h, uint64_t cap) { return (h * 0x9E3779B97F4A7C15ULL) >> 20 & (cap - 1); }

static void grow(yko_t *o)
{
	const uint64_t ncap = o->cap * 2;
	uint64_t *nk = (uint64_t *)calloc(ncap, sizeof *nk);
	uint16_t *nc = (uint16_t *)calloc(ncap, sizeof *nc);
	for (uint64_t i = 0; i < o->cap; ++i) {
		if (!o->key[i]) continue;
		uint64_t s = slot_of(o->key[i] - 1, ncap);
		while (nk[s]) s = (s + 1) & (ncap - 1);
		nk[s] = o->key[i], nc[s] = o->cnt[i];
	}
	free(o->key), free(o->cnt);
	o->key = nk, o->cnt = nc, o->cap = ncap;
}

void yko_add_hashed(yko_t *o, uint64_t h, int create_new) /* yak-count.c:150-177 */
{
	if (create_new) {
		if (o->bloom) { /* only a k-mer whose bits were all set already gets an entry */
			const uint64_t part = h & ((1ULL << o->pre) - 1);
			if (yko_bf_insert(o->bloom[part], o->part_shift, o->n_hash, h >> o->pre) != o->n_hash) return;
		}
		if (o->used * 10 >= o->cap * 6) grow(o);
	}
	uint64_t s = slot_of(h, o->cap);
	while (o->key[s] && o->key[s] != h + 1) s = (s + 1) & (o->cap - 1);
	if (!o->key[s]) {
		if (!create_new) return; /* second pass: only what the first pass let in is counted */
		o->key[s] = h + 1;
		o->used++;
	}
	if (o->cnt[s] < YKO_MAX) o->cnt[s]++;
}

void yko_add_read(yko_t *o, const char *seq, long len, int create_new) /* yak-count.c:345-361,388 */
{
	if (len < o->k) return;
	if (o->scratch_cap < len) {
		o->scratch_cap = len + 1024;
		o->scratch = (uint64_t *)realloc(o->scratch, (size_t)o->scratch_cap * sizeof *o->scratch);
	}
	const long n = kco_hashed_kmers(seq, len, o->k, o->scratch);
	for (long i = 0; i < n; ++i) yko_add_hashed(o, o->scratch[i], create_new);
}

/* yak-count.c:378-400 under kt_pipeline(3, ...) (:440): blocks of chunk_size bases; a record the
 * reader rejects closes the block; the file ends with the third empty block (kthread.c:97-125),
 * as in kc-c4 */
int yko_add_file(yko_t *o, const char *fn, long chunk_size, int create_new)
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
			yko_add_read(o, seq, len, create_new);
			sum_len += len;
			if (sum_len >= chunk_size) break;
		}
		if (sum_len == 0 && --lives == 0) break;
	}
	orr_close(fx);
	return 0;
}

void yko_second_pass(yko_t *o) /* yak-count.c:451-452 */
{
	drop_bloom(o);
	memset(o->cnt, 0, o->cap * sizeof *o->cnt);
}

uint64_t yko_shrink(yko_t *o, int min, int max) /* yak-count.c:247-282 */
{
	uint64_t *nk = (uint64_t *)calloc(o->cap, sizeof *nk);
	uint16_t *nc = (uint16_t *)calloc(o->cap, sizeof *nc);
	uint64_t kept = 0;
	for (uint64_t i = 0; i < o->cap; ++i) {
		if (!o->key[i] || o->cnt[i] < min || o->cnt[i] > max) continue;
		uint64_t s = slot_of(o->key[i] - 1, o->cap);
		while (nk[s]) s = (s + 1) & (o->cap - 1);
		nk[s] = o->key[i], nc[s] = o->cnt[i];
		++kept;
	}
	free(o->key), free(o->cnt);
	o->key = nk, o->cnt = nc, o->used = kept;
	return kept;
}

int yko_count_files(yko_t *o, const char *fn1, const char *fn2, long chunk_size) /* yak-count.c:445-456 */
{
	if (yko_add_file(o, fn1, chunk_size, 1) != 0) return -1;
	if (o->bf_shift > 0) {
		yko_second_pass(o);
		if (yko_add_file(o, fn2 ? fn2 : fn1, chunk_size, 0) != 0) return -1;
		yko_shrink(o, 2, YKO_MAX);
	}
	return 0;
}

void yko_hist(const yko_t *o, uint64_t hist[1024]) /* yak-count.c:205-239 */
{
	memset(hist, 0, 1024 * sizeof hist[0]);
	for (uint64_t i = 0; i < o->cap; ++i)
		if (o->key[i]) hist[o->cnt[i]]++;
}

uint64_t yko_distinct(const yko_t *o) { return o->used; }

void yko_print_hist(const uint64_t hist[1024], FILE *fp) /* yak-count.c:503 */
{
	for (int i = 1; i < 1024; ++i) fprintf(fp, "%d\t%lld\n", i, (long long)hist[i]);
}
