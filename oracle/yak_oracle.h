/*
 * yak_oracle.h -- CPU restatement of yak-count (k-mer counting with a blocked Bloom pre-filter).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product: only
 * tests/, __graft_entry__.smoke() and the cpu_baseline legs of the bench scripts may build,
 * link, import or execute it, and there only as the checker.
 *
 * Parity status: PINNED.  tests/test_yak_oracle.py compares the 1023 lines printed by this
 * restatement byte for byte with the output of the unmodified reference, compiled from
 * /root/reference/yak-count.c into oracle/_ref/yak-count by oracle/Makefile, on live inputs
 * (when oracle/_ref exists) and on the committed fixtures tests/golden/yak/ (made by that
 * binary, tests/golden/make_golden.sh), with and without the Bloom filter, one and two files.
 *
 * Each function cites the reference lines it restates (paths relative to the reference
 * checkout).  Written from the behaviour, not copied.
 */
#ifndef YAK_ORACLE_H
#define YAK_ORACLE_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yko yko_t;

/* k, -p prefix bits, -b Bloom bits (log2, 0 = none), -H hash functions (yak-count.c:312-321,463-470) */
yko_t *yko_create(int k, int pre, int bf_shift, int bf_n_hash);
void yko_destroy(yko_t *o);
/* one hashed k-mer through yak_ch_insert_list (yak-count.c:150-177): create_new = first pass */
void yko_add_hashed(yko_t *o, uint64_t h, int create_new);
/* one read: yak-count.c:345-361 (canonical k-mers, hash64) */
void yko_add_read(yko_t *o, const char *seq, long len, int create_new);
/* a whole file in blocks of chunk_size bases as yak-count.c:378-400 reads it; -1 if it cannot be opened */
int yko_add_file(yko_t *o, const char *fn, long chunk_size, int create_new);
/* between the passes: drop the Bloom filters, zero the counts (yak-count.c:139-148,189-203,451-452) */
void yko_second_pass(yko_t *o);
/* keep counts in [min, max] (yak-count.c:247-282); returns what is left */
uint64_t yko_shrink(yko_t *o, int min, int max);
/* yak_count_file (yak-count.c:445-456): both passes and the shrink; fn2 may be NULL */
int yko_count_files(yko_t *o, const char *fn1, const char *fn2, long chunk_size);
/* hist[c] = entries with count c (yak-count.c:205-239) */
void yko_hist(const yko_t *o, uint64_t hist[1024]);
uint64_t yko_distinct(const yko_t *o);
/* the 1023 lines yak-count prints (yak-count.c:503) */
void yko_print_hist(const uint64_t hist[1024], FILE *fp);
/* Bloom insert as the reference does it: how many of the n_hash bits were already set
 * (yak-count.c:86-104); exported for known-answer tests */
int yko_bf_insert(uint8_t *bits, int n_shift, int n_hash, uint64_t hash);

#ifdef __cplusplus
}
#endif
#endif
