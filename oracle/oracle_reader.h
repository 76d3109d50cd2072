/*
 * oracle_reader.h -- TEST INFRASTRUCTURE ONLY: the checker's own FASTA/FASTQ record reader.
 *
 * The oracle must not share its parser with the product (kmer-cnt_b200/host/fastx.c), or
 * agreement between the two would say nothing about parsing.  This one restates the record
 * rules of kseq_read (kseq.h:181-232 of the reference) over a file that is decompressed into
 * memory first; tests/test_oracle.py pins it against the live reference binaries on awkward
 * files, the product's reader is pinned the same way in its own tests.
 */
#ifndef ORACLE_READER_H
#define ORACLE_READER_H

typedef struct orr orr_t;

orr_t *orr_open(const char *fn);          /* NULL if the file cannot be opened (gzopen, vaf-counter.c:557) */
void orr_close(orr_t *r);
/* next record: sequence length and *seq (valid until the next call), -1 at the end of the
 * input, -2 for a FASTQ record whose quality string is short or missing (kseq.h:187-191) */
long orr_next(orr_t *r, const char **seq);
const char *orr_name(const orr_t *r);     /* header of that record up to the first white space */

#endif
