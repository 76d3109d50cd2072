/*
 * kc_oracle.h -- CPU restatement of kc-c4's full k-mer counting path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product: only
 * tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of the bench
 * scripts may build, link, import or execute it, and there only as the checker.
 *
 * Parity status: PINNED.  tests/test_kc_oracle.py compares the histogram printed by this
 * restatement byte for byte with the output of the unmodified reference, compiled from
 * /root/reference/kc-c4.c into oracle/_ref/kc-c4 by oracle/Makefile, on live inputs (when
 * oracle/_ref exists) and on the committed fixtures tests/golden/kc_* (made by that binary,
 * script beside them).
 *
 * Each function cites the reference lines it restates (paths relative to the reference
 * checkout).  Written from the behaviour, not copied.
 */
#ifndef KC_ORACLE_H
#define KC_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kco kco_t;

/* byte -> {0,1,2,3,4}                                         (kc-c4.c:21-38)  */
int kco_nt4(uint8_t b);
/* invertible mix of a 2k-bit word                             (kc-c4.c:40-50)  */
uint64_t kco_hash64(uint64_t key, int k);
/* canonical k-mers of one read, hashed, in stream order       (kc-c4.c:74-90)
 * out must hold len - k + 1 words; returns how many were written */
long kco_hashed_kmers(const char *seq, long len, int k, uint64_t *out);

kco_t *kco_create(int k);
void kco_destroy(kco_t *o);
/* one read: kc-c4.c:141 (reads shorter than k are dropped), :74-90, :116-128 (count
 * saturates at 1023) */
void kco_add_read(kco_t *o, const char *seq, long len);
/* one hashed k-mer, as the insert step sees it                (kc-c4.c:116-128) */
void kco_add_hashed(kco_t *o, uint64_t h);
/* whole file through the FASTA/FASTQ reader, in blocks of block_len bases as kc-c4.c:133-153
 * reads it (that decides where a malformed FASTQ record ends the file); -1 if it cannot be
 * opened (kc-c4.c:166) */
int kco_add_file(kco_t *o, const char *fn, long block_len);
/* hist[c] = distinct k-mers seen min(c, 255) times            (kc-c4.c:186-215) */
void kco_hist(const kco_t *o, uint64_t hist[256]);
uint64_t kco_distinct(const kco_t *o);
uint64_t kco_instances(const kco_t *o);
/* the 255 lines kc-c4 prints                                  (kc-c4.c:232-233) */
void kco_print_hist(const uint64_t hist[256], FILE *fp);

#ifdef __cplusplus
}
#endif
#endif
