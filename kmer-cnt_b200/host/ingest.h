/*
 * ingest.h -- parallel host ingest for vaf-counter: several reader threads, each feeding a
 * producer of its own (vafgpu_producer_*), in place of the reference's single kseq reader
 * (step 0 of worker_pipeline, vaf-counter.c:486-517, which bounds the reference end to end).
 *
 * Work is cut into units.  A gzip file, a FASTA file, a pipe, or anything that is not a plain
 * four-line FASTQ is one unit, parsed by the exact streaming reader (fastx.c: the reference's
 * record rules, kseq.h:192-232).  A plain four-line FASTQ file is memory-mapped and cut into
 * slices; every slice is a unit.  Slicing never changes what is counted:
 *   pass 1  every slice guesses its first record (a line that starts with '@' and is followed,
 *           two lines on, by a line that starts with '+'), walks strictly formed records
 *           ('@' line, one sequence line that does not start with '@', '+' or '>', '+' line,
 *           one quality line of the same length, no '\r', no blank lines) up to the first
 *           record that starts in the next slice, and reports where that is;
 *   check   slice 0 starts at offset 0 and every slice ends exactly where the next one
 *           guessed its start, the last one at end of file.  The records then form one chain
 *           from the first byte of the file, which is precisely what the sequential reader
 *           would have produced;
 *   pass 2  the slices are parsed again, this time handing the reads to the engine.
 * A file that fails the check in any way is read sequentially instead.
 */
#ifndef KMERCNT_INGEST_H
#define KMERCNT_INGEST_H

#include <stddef.h>
#include <stdint.h>

#include "../../include/vafgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	uint64_t seqs, bases; /* reads with len >= k and their bases, as vaf-counter.c:505-506 counts them */
	double seconds;       /* first unit started .. last unit finished */
	int opened;           /* 0: could not be opened (silently skipped, vaf-counter.c:557) */
	int sliced;           /* number of slices it was cut into (0 = read sequentially) */
} ingest_file_t;

/* Where the readers hand their reads: one producer per reader thread (vafgpu_producer_* for
 * vaf-counter, kcgpu_producer_* for kc-c4). */
typedef struct ingest_sink {
	void *engine;
	int (*producer_create)(void *engine, void **producer);
	int (*producer_add_read)(void *producer, const char *seq, size_t len);
	int (*producer_destroy)(void *producer);
	const char *(*error)(void *engine);
} ingest_sink_t;

int ingest_files_to(const ingest_sink_t *sink, int n_files, char **files, int k, int block_len, int n_threads,
                    ingest_file_t *out);

/* Count every file with n_threads reader threads.  block_len is -b (the sequential reader
 * closes a block when it holds that many bases, vaf-counter.c:509).  Returns 0, or -1 after
 * printing the engine's error. */
int ingest_files(vafgpu_ctx *ctx, int n_files, char **files, int k, int block_len, int n_threads,
                 ingest_file_t *out);

#ifdef __cplusplus
}
#endif
#endif
