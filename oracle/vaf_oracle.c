/*
 * vaf_oracle.c -- CPU restatement of the reference's extract-and-lookup path.
 * TEST INFRASTRUCTURE ONLY (see vaf_oracle.h).  Parity: pinned against oracle/_ref.
 */
#include "vaf_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_reader.h"

/* ---- byte classification ------------------------------------------------------- */

/* vaf-counter.c:73-90: A/a C/c G/g T/t U/u and the control bytes 0..3 are bases. */
int vo_nt4_strict(uint8_t b)
{
	if (b < 4) return b;
	switch (b | 0x20) {
	case 'a': return 0;
	case 'c': return 1;
	case 'g': return 2;
	case 't':
	case 'u': return 3;
	}
	return 4;
}

/* vaf-counter.c:272-283: PSHUFB over (byte & 15) with the 16-entry table
 * {4,0,4,1,3,3,4,2,4,...,4}.  (A byte >= 0x80 would make PSHUFB return 0, but the
 * index is masked with 0x0F first, so only the low nibble matters.) */
int vo_nt4_nibble(uint8_t b)
{
	static const int8_t lut[16] = {4, 0, 4, 1, 3, 3, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4};
	return lut[b & 15];
}

/* vaf-counter.c:278-290 */
int vo_code_at(const char *seq, int len, int i, int simd)
{
	uint8_t b = (uint8_t)seq[i];
	if (simd && i < (len & ~15)) return vo_nt4_nibble(b);
	return vo_nt4_strict(b);
}

/* ---- k-mer arithmetic ---------------------------------------------------------- */

uint64_t vo_encode_kmer(const char *s, int k)
{
	uint64_t x = 0;
	for (int i = 0; i < k; ++i) {
		int c = vo_nt4_strict((uint8_t)s[i]);
		if (c > 3) return VO_NO_KMER;
		x = x << 2 | (uint64_t)c;
	}
	return x;
}

uint64_t vo_revcomp(uint64_t x, int k)
{
	uint64_t r = 0;
	for (int i = 0; i < k; ++i, x >>= 2) r = r << 2 | (3 - (x & 3));
	return r;
}

uint64_t vo_canonical(uint64_t x, int k)
{
	uint64_t r = vo_revcomp(x, k);
	return x < r ? x : r;
}

uint32_t vo_kmer_hash(uint64_t key)
{
	key ^= key >> 33;
	key *= 0xff51afd7ed558ccdULL;
	key ^= key >> 33;
	return (uint32_t)key;
}

uint32_t vo_h2b(uint32_t hash, uint32_t bits)
{
	return hash * 2654435769u >> (32 - bits);
}

/* ---- patterns -------------------------------------------------------------------- */

vo_patterns_t *vo_load_patterns(const char *fn)
{
	FILE *fp = fopen(fn, "r");
	vo_patterns_t *db;
	vo_pattern_t p;
	if (!fp) return NULL;
	db = (vo_patterns_t *)calloc(1, sizeof(*db));
	/* eight white-space separated fields; the first malformed record ends the load */
	while (fscanf(fp, "%255s%d%d%255s %c %c%127s%127s", p.chr, &p.start, &p.end, p.rsid,
	              &p.ref, &p.alt, p.ref_kmer, p.alt_kmer) == 8) {
		if (db->n == db->m) {
			db->m = db->m ? db->m * 2 : 16;
			db->a = (vo_pattern_t *)realloc(db->a, (size_t)db->m * sizeof(vo_pattern_t));
		}
		p.ref_count = p.alt_count = 0;
		db->a[db->n++] = p;
	}
	fclose(fp);
	return db;
}

void vo_patterns_free(vo_patterns_t *db)
{
	if (!db) return;
	free(db->a);
	free(db);
}

/* ---- the map --------------------------------------------------------------------- */

#define USED(m, i) ((m)->used[(i) >> 5] >> ((i) & 31) & 1u)

/* khashl.h:152-160: buckets = smallest power of two >= the request, at least 4 */
static uint32_t bits_for(uint32_t want)
{
	uint32_t j = 0, x = want;
	while ((x >>= 1) != 0) ++j;
	if (want & (want - 1)) ++j;
	return j > 2 ? j : 2;
}

static void map_alloc(vo_map_t *m, uint32_t bits)
{
	uint32_t nb = 1u << bits;
	m->bits = bits;
	m->used = (uint32_t *)calloc(nb < 32 ? 1 : nb >> 5, 4);
	m->key = (uint64_t *)malloc((size_t)nb * 8);
	m->val = (uint32_t *)malloc((size_t)nb * 4);
}

/* khashl.h:197-221: probe linearly from h2b(hash); stop at an empty bucket or an equal
 * key.  Growth at 75 % load (khashl.h:202) is restated as a rebuild in bucket order,
 * which yields the same membership; with the n*3 pre-sizing of vaf-counter.c:216 it
 * never triggers. */
static uint32_t map_put(vo_map_t *m, uint64_t key, int *absent);

static void map_grow(vo_map_t *m)
{
	vo_map_t old = *m;
	uint32_t nb_old = 1u << old.bits;
	map_alloc(m, bits_for(nb_old + 1));
	m->count = 0;
	for (uint32_t j = 0; j < nb_old; ++j)
		if (USED(&old, j)) {
			int a;
			uint32_t i = map_put(m, old.key[j], &a);
			m->val[i] = old.val[j];
		}
	free(old.used);
	free(old.key);
	free(old.val);
}

static uint32_t map_put(vo_map_t *m, uint64_t key, int *absent)
{
	uint32_t nb = 1u << m->bits, mask, i, first;
	if (m->count >= (nb >> 1) + (nb >> 2)) {
		map_grow(m);
		nb = 1u << m->bits;
	}
	mask = nb - 1;
	i = first = vo_h2b(vo_kmer_hash(key), m->bits);
	while (USED(m, i) && m->key[i] != key) {
		i = (i + 1) & mask;
		if (i == first) break;
	}
	if (!USED(m, i)) {
		m->key[i] = key;
		m->used[i >> 5] |= 1u << (i & 31);
		++m->count;
		*absent = 1;
	} else *absent = 0;
	return i;
}

vo_map_t *vo_map_build(const vo_patterns_t *db, int k)
{
	vo_map_t *m = (vo_map_t *)calloc(1, sizeof(*m));
	map_alloc(m, bits_for((uint32_t)db->n * 3)); /* vaf-counter.c:216 */
	for (int i = 0; i < db->n; ++i) {
		for (int alt = 0; alt < 2; ++alt) { /* ref first, then alt: vaf-counter.c:221-243 */
			uint64_t x = vo_encode_kmer(alt ? db->a[i].alt_kmer : db->a[i].ref_kmer, k);
			int absent;
			uint32_t b;
			if (x == VO_NO_KMER) continue;
			b = map_put(m, vo_canonical(x, k), &absent);
			if (absent) m->val[b] = (uint32_t)i << 1 | (uint32_t)alt;
			else ++m->n_collisions; /* first insert keeps the bucket */
		}
	}
	return m;
}

void vo_map_free(vo_map_t *m)
{
	if (!m) return;
	free(m->used);
	free(m->key);
	free(m->val);
	free(m);
}

uint32_t vo_map_get(const vo_map_t *m, uint64_t key)
{
	uint32_t nb = 1u << m->bits, mask = nb - 1, i, first;
	i = first = vo_h2b(vo_kmer_hash(key), m->bits);
	while (USED(m, i) && m->key[i] != key) {
		i = (i + 1) & mask;
		if (i == first) return nb;
	}
	return USED(m, i) ? i : nb;
}

uint32_t vo_map_export(const vo_map_t *m, uint64_t *keys, uint32_t *vals)
{
	uint32_t nb = 1u << m->bits, n = 0;
	for (uint32_t i = 0; i < nb; ++i)
		if (USED(m, i)) {
			if (keys) keys[n] = m->key[i];
			if (vals) vals[n] = m->val[i];
			++n;
		}
	return n;
}

/* ---- extraction + lookup ------------------------------------------------------------ */

/* vaf-counter.c:349-427: forward word shifts left, reverse-complement word shifts right,
 * the smaller of the two is emitted once k valid bases have been seen in a row; any
 * other byte clears the run. */
static inline uint64_t roll(int k, const char *seq, int len, int simd, uint64_t *out,
                            const vo_map_t *m, uint32_t *counts)
{
	const uint64_t mask = (1ULL << 2 * k) - 1;
	const int shift = 2 * (k - 1);
	const int simd_len = simd ? (len & ~15) : 0;
	uint64_t fw = 0, rc = 0, n = 0;
	int run = 0;
	for (int i = 0; i < len; ++i) {
		uint8_t b = (uint8_t)seq[i];
		int c = i < simd_len ? vo_nt4_nibble(b) : vo_nt4_strict(b);
		if (c > 3) {
			run = 0;
			fw = rc = 0;
			continue;
		}
		fw = (fw << 2 | (uint64_t)c) & mask;
		rc = rc >> 2 | (uint64_t)(3 - c) << shift;
		if (++run < k) continue;
		uint64_t y = fw < rc ? fw : rc;
		if (out) out[n] = y;
		if (m) { /* vaf-counter.c:462-477 */
			uint32_t b2 = vo_map_get(m, y);
			if (b2 != 1u << m->bits) ++counts[m->val[b2]]; /* val = idx<<1 | is_alt */
		}
		++n;
	}
	return n;
}

uint64_t vo_count_read(const vo_map_t *m, int k, const char *seq, int len, int simd,
                       uint32_t *counts)
{
	return roll(k, seq, len, simd, NULL, m, counts);
}

uint64_t vo_extract_read(int k, const char *seq, int len, int simd, uint64_t *out)
{
	return roll(k, seq, len, simd, out, NULL, NULL);
}

/* ---- output -------------------------------------------------------------------------- */

int vo_write_vaf(FILE *fp, const vo_patterns_t *db, const uint32_t *counts)
{
	uint64_t tot_ref = 0, tot_alt = 0;
	for (int i = 0; i < db->n; ++i) {
		tot_ref += counts[2 * i];
		tot_alt += counts[2 * i + 1];
	}
	fprintf(fp, "# Average depth: %.2f\n",
	        (double)(tot_ref + tot_alt) / (db->n > 0 ? db->n : 1));
	fprintf(fp, "CHR\tPOS\tRSID\tREF\tALT\tREF_COUNT\tALT_COUNT\tTOTAL_COUNT\tVAF\n");
	for (int i = 0; i < db->n; ++i) {
		const vo_pattern_t *p = &db->a[i];
		uint32_t r = counts[2 * i], a = counts[2 * i + 1], t = r + a;
		fprintf(fp, "%s\t%d\t%s\t%c\t%c\t%u\t%u\t%u\t%.4f\n", p->chr, p->start, p->rsid,
		        p->ref, p->alt, r, a, t, t > 0 ? (double)a / t : 0.0);
	}
	return ferror(fp) ? -1 : 0;
}

/* ---- whole-file driver ----------------------------------------------------------------- */

typedef struct {
	const vo_map_t *m;
	int k, simd, tid, nt;
	int n;
	char **seq;
	int *len;
	uint32_t *counts;
	uint64_t n_kmers;
} job_t;

static void *job_run(void *arg)
{
	job_t *j = (job_t *)arg;
	for (int i = j->tid; i < j->n; i += j->nt)
		j->n_kmers += vo_count_read(j->m, j->k, j->seq[i], j->len[i], j->simd, j->counts);
	return NULL;
}

int vo_count_file(const vo_map_t *m, int k, const char *fn, int simd, int n_threads,
                  int block_len, uint32_t *counts, vo_stats_t *st)
{
	orr_t *fx = orr_open(fn);
	size_t n_counts = 0, cap = 0;
	char **seq = NULL;
	int *len = NULL;
	uint32_t *priv = NULL;
	if (!fx) return -1;
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 256) n_threads = 256;
	for (uint32_t i = 0, nb = 1u << m->bits; i < nb; ++i)
		if (USED(m, i) && m->val[i] + 1 > n_counts) n_counts = m->val[i] + 1;
	n_counts = (n_counts + 1) & ~(size_t)1;
	if (n_threads > 1) priv = (uint32_t *)calloc((size_t)n_threads * n_counts, 4);
	int lives = 3; /* kt_pipeline(3, ...), vaf-counter.c:568 */
	for (;;) {
		/* one block: vaf-counter.c:486-517.  Any negative return (end of input or a bad
		 * FASTQ record) closes the block.  An empty block retires the pipeline worker that
		 * read it (kthread.c:97-125: a worker leaves when ITS step 0 returns NULL); the other
		 * workers go on calling step 0, in order, so the file ends with the third empty block. */
		size_t n = 0;
		long l, sum_len = 0;
		const char *s;
		while ((l = orr_next(fx, &s)) >= 0) {
			if (l < k) continue;
			if (n == cap) {
				size_t ncap = cap ? cap + (cap >> 1) : 1024;
				seq = (char **)realloc(seq, ncap * sizeof(char *));
				len = (int *)realloc(len, ncap * sizeof(int));
				memset(seq + cap, 0, (ncap - cap) * sizeof(char *));
				cap = ncap;
			}
			seq[n] = (char *)realloc(seq[n], (size_t)l);
			memcpy(seq[n], s, (size_t)l);
			len[n++] = (int)l;
			sum_len += l;
			if (st) st->n_reads++, st->n_bases += (uint64_t)l;
			if (sum_len >= block_len) break;
		}
		if (sum_len == 0) {
			if (--lives == 0) break;
			continue;
		}
		if (n_threads == 1) {
			for (size_t i = 0; i < n; ++i) {
				uint64_t nk = vo_count_read(m, k, seq[i], len[i], simd, counts);
				if (st) st->n_kmers += nk;
			}
		} else {
			pthread_t th[256];
			job_t jb[256];
			for (int t = 0; t < n_threads; ++t) {
				jb[t] = (job_t){m, k, simd, t, n_threads, (int)n, seq, len,
				                priv + (size_t)t * n_counts, 0};
				pthread_create(&th[t], NULL, job_run, &jb[t]);
			}
			for (int t = 0; t < n_threads; ++t) {
				pthread_join(th[t], NULL);
				if (st) st->n_kmers += jb[t].n_kmers;
			}
		}
	}
	if (priv) {
		for (int t = 0; t < n_threads; ++t)
			for (size_t i = 0; i < n_counts; ++i) counts[i] += priv[(size_t)t * n_counts + i];
		free(priv);
	}
	for (size_t i = 0; i < cap; ++i) free(seq[i]);
	free(seq);
	free(len);
	orr_close(fx);
	return 0;
}
