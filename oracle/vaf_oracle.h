/*
 * vaf_oracle.h -- CPU restatement of vaf-counter's k-mer extract-and-lookup path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, link, import or execute it, and there only as the checker.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_vs_ref.py,
 * tests/golden/) against outputs of the unmodified reference compiled from
 * /root/reference into oracle/_ref/ by oracle/Makefile.
 *
 * Each function cites the reference lines it restates (paths relative to the
 * reference checkout).  The code is written from the behaviour, not copied.
 */
#ifndef VAF_ORACLE_H
#define VAF_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VO_NO_KMER UINT64_MAX

/* byte -> {0,1,2,3,4}; strict 256-entry table          (vaf-counter.c:73-90)  */
int vo_nt4_strict(uint8_t b);
/* byte -> {0,1,2,3,4}; SSSE3 low-nibble lookup         (vaf-counter.c:272-276) */
int vo_nt4_nibble(uint8_t b);
/* code at offset i of a read of length len under the build the Makefile produces:
 * nibble rule for i < (len & ~15), strict rule for the tail (vaf-counter.c:278-290).
 * simd == 0 selects the scalar build (strict everywhere, vaf-counter.c:341-343). */
int vo_code_at(const char *seq, int len, int i, int simd);

uint64_t vo_encode_kmer(const char *s, int k);          /* vaf-counter.c:117-127 */
uint64_t vo_revcomp(uint64_t x, int k);                 /* vaf-counter.c:130-139 */
uint64_t vo_canonical(uint64_t x, int k);               /* vaf-counter.c:142-146 */
uint32_t vo_kmer_hash(uint64_t key);                    /* vaf-counter.c:56-63   */
uint32_t vo_h2b(uint32_t hash, uint32_t bits);          /* khashl.h:98           */

/* patterns.txt record                                   (vaf-counter.c:92-103)  */
typedef struct {
	char chr[256];
	int start, end;
	char rsid[256];
	char ref, alt;
	char ref_kmer[128], alt_kmer[128];
	uint32_t ref_count, alt_count;
} vo_pattern_t;

typedef struct {
	int n, m;
	vo_pattern_t *a;
} vo_patterns_t;

vo_patterns_t *vo_load_patterns(const char *fn);        /* vaf-counter.c:149-184 */
void vo_patterns_free(vo_patterns_t *db);

/* canonical k-mer -> (pattern index << 1 | is_alt); open addressing with the
 * reference's geometry and probe order (vaf-counter.c:198-252, khashl.h:137-221) */
typedef struct {
	uint32_t bits, count;
	uint32_t *used;       /* 1 bit per bucket */
	uint64_t *key;
	uint32_t *val;
	int n_collisions;     /* duplicates met while building (first insert wins) */
} vo_map_t;

vo_map_t *vo_map_build(const vo_patterns_t *db, int k);
void vo_map_free(vo_map_t *m);
/* returns bucket index or (1u<<bits) when absent       (khashl.h:137-150)       */
uint32_t vo_map_get(const vo_map_t *m, uint64_t key);

/* flat list of distinct canonical keys and their values in insertion order; this is
 * exactly what the C-ABI's vafgpu_create() takes.  Returns the entry count.        */
uint32_t vo_map_export(const vo_map_t *m, uint64_t *keys, uint32_t *vals);

/* rolling canonical k-mers of one read + lookup + count (vaf-counter.c:349-427,
 * 449-479).  counts[2*i] = ref, counts[2*i+1] = alt.  Returns k-mers emitted.      */
uint64_t vo_count_read(const vo_map_t *m, int k, const char *seq, int len,
                       int simd, uint32_t *counts);

/* same, but only emits the canonical k-mers (for unit comparison with the
 * reference's extractor); out may be NULL to just count.                            */
uint64_t vo_extract_read(int k, const char *seq, int len, int simd, uint64_t *out);

/* VAF writer                                            (vaf-counter.c:654-680)  */
int vo_write_vaf(FILE *fp, const vo_patterns_t *db, const uint32_t *counts);

/* whole-file driver used by the oracle CLI and the bench CPU leg: parses FASTA/FASTQ
 * (plain or gz) with the same record rules as kseq.h:192-232, skips reads shorter
 * than k (vaf-counter.c:494), forms blocks of >= block_len bases with the reference's
 * stop rules (vaf-counter.c:492,509,513) and adds into counts.  n_threads > 1 splits the
 * reads of a block across threads (counts are summed; the result does not depend on it).
 * Returns 0, or -1 if the file cannot be opened (the reference silently skips it,
 * vaf-counter.c:557).                                                                */
typedef struct {
	uint64_t n_reads, n_bases, n_kmers;
} vo_stats_t;
int vo_count_file(const vo_map_t *m, int k, const char *fn, int simd, int n_threads,
                  int block_len, uint32_t *counts, vo_stats_t *st);

#ifdef __cplusplus
}
#endif
#endif
