/*
 * fastx.h -- streaming FASTA/FASTQ record reader for the host side of vaf-counter.
 *
 * Written from scratch; it reproduces the record rules of the reader the reference
 * uses (kseq.h:192-232 over gzread, vaf-counter.c:12,557-559): multi-line FASTA and
 * FASTQ, '>' / '@' headers, blank lines skipped, a trailing '\r' dropped per line,
 * plain or gzip input.  Only what the counting path consumes is kept: the sequence.
 */
#ifndef KMERCNT_FASTX_H
#define KMERCNT_FASTX_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fastx fastx_t;

/* NULL if the file cannot be opened (the reference then silently skips the file,
 * vaf-counter.c:557). */
fastx_t *fastx_open(const char *fn);
void fastx_close(fastx_t *fx);

/* Next record.  Returns the sequence length (>= 0) and points *seq at it (valid until
 * the next call, not NUL-terminated beyond len), or
 *   -1  end of input
 *   -2  FASTQ record whose quality string is missing or of a different length
 * exactly where kseq_read would (kseq.h:187-191,227-231). */
long fastx_next(fastx_t *fx, const char **seq);

/* Name of the record fastx_next returned last: the header up to the first white space
 * (kseq.h:199-204), NUL-terminated, valid until the next call. */
const char *fastx_name(const fastx_t *fx);

#ifdef __cplusplus
}
#endif
#endif
