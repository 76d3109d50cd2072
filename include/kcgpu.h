/*
 * kcgpu.h -- C ABI of the full k-mer counting mode (the kc-c4 path of gerbenvoshol/kmer-cnt)
 * on B200.  Part of libvafgpu.so.
 *
 * What it replaces, in kc-c4.c of the reference: count_seq_buf (kc-c4.c:74-90: rolling forward
 * and reverse-complement words over the strict base table kc-c4.c:21-38, canonical minimum,
 * the invertible hash64 of kc-c4.c:40-50), c4x_insert_buf / worker_for (kc-c4.c:64-72,116-128:
 * one khashl set per hash suffix, a 10-bit saturating count in the low bits of the key) and
 * worker_hist / print_hist (kc-c4.c:186-215: 256 bins of min(count, 255)).
 *
 * hash64 is a bijection on 2k-bit words, so the reference's tables hold exactly one entry per
 * distinct canonical k-mer; the histogram is a function of the multiset of canonical k-mers and
 * of nothing else.  Here one open-addressing table per GPU holds the reference's own slot word
 * (hash bits << 10 | count).  Like the reference, the scan does not touch the table: it files
 * every hashed k-mer under its region (hash suffix) in a list, and the lists are emptied into
 * the table region by region ("flush"), so that the slice of the table being filled stays in
 * L2.  With several GPUs a k-mer belongs to GPU hash64(k-mer) mod n (the reference's partition
 * by hash suffix, kc-c4.c:66, generalised from 2^p tables to n owners).  Two ways to get it there:
 *   fused    kcgpu_set_owners(): the counting kernel itself files every k-mer in its owner's
 *            lists, over NVLink peer memory for the other GPUs (no exchange buffers);
 *   staged   kcgpu_extract_device() files hashed k-mers per owner, the caller exchanges the
 *            lists (NCCL all-to-all), kcgpu_insert_device() adds what arrived.
 *
 * Plain C types only; 0 on success or a negative VAFGPU_E* code (vafgpu.h); no CPU fallback:
 * kcgpu_create fails with VAFGPU_ENOGPU without an sm_100 device.
 */
#ifndef KCGPU_H
#define KCGPU_H

#include <stddef.h>
#include <stdint.h>

#include "vafgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kcgpu_ctx kcgpu_ctx;
typedef struct kcgpu_producer kcgpu_producer;

#define KCGPU_MAX_OWNERS 16
#define KCGPU_IPC_HANDLE_BYTES 64
#define KCGPU_NO_LISTS UINT64_MAX /* list_slots: no region lists, every k-mer goes straight to the table */

typedef struct kcgpu_stats {
	uint64_t n_reads;     /* reads accepted by kcgpu_add_read (len >= k)                      */
	uint64_t n_bases;
	uint64_t n_blocks;    /* kernel launches over stream blocks                                */
	uint64_t n_kmers;     /* k-mer instances this context's kernels extracted or inserted      */
	uint64_t n_distinct;  /* slots this context's kernels claimed (in any owner's table)       */
	uint64_t n_overflow;  /* k-mers lost because a table region was full: counts incomplete    */
	uint64_t n_dropped;   /* k-mers that did not fit the lists of kcgpu_extract_device         */
	uint64_t n_direct;    /* k-mers that found their region list full and went straight to the table */
	uint64_t n_flushes;
	uint64_t table_slots;
	uint64_t list_slots;  /* capacity of the region lists, all regions together                */
	uint64_t flush_bytes; /* stream bytes a context takes between two flushes                  */
	double   kernel_ms;   /* kernels behind kcgpu_add_read (CUDA events on their streams)      */
	double   h2d_ms;
} kcgpu_stats;

/* sm_100 devices visible to this process (0 if there is none) */
int kcgpu_device_count(void);

/*
 * One context = one table and its region lists on one device.  table_slots is rounded up to a
 * power of two (at least 4096); 0 = the largest power of two that, with its lists, fits in 3/4
 * of the device's free memory (8 bytes per slot).  list_slots: capacity of the region lists
 * (8 bytes each; 0 = table_slots / 2; KCGPU_NO_LISTS = none).  block_bytes: size of each pinned
 * staging block behind kcgpu_add_read (0 = 16 MiB).
 */
int kcgpu_create(kcgpu_ctx **ctx, int k, uint64_t table_slots, uint64_t list_slots, size_t block_bytes, int device);

/*
 * The yak-count variant of the counting path (yak-count.c of the reference: the same tables and
 * the same 10-bit counts as kc-c4, plus a blocked Bloom filter in front of them and a second
 * pass, yak-count.c:71-104,150-177,445-456).
 *
 * kcgpu_create_filtered: a context with a Bloom filter of 2^bloom_bits bits (cut down to a
 * quarter of the device's memory; 0 = none) and bloom_hashes bits per k-mer (yak-count -b / -H).
 * kcgpu_set_pass says what the insert step does with the k-mers counted from then on (it
 * flushes what was filed before):
 *   KCGPU_PASS_COUNT   make an entry if there is none, count up to 1023      (kc-c4; yak-count -b 0)
 *   KCGPU_PASS_CLAIM   yak-count's first pass with a filter: an entry, with a count of 0, for every
 *                      k-mer the filter has seen before -- the second and later occurrences,
 *                      and false positives; k-mers seen once never reach the table
 *   KCGPU_PASS_LOOKUP  its second pass: count the k-mers that have an entry, skip the others
 * kcgpu_histogram1024 is yak-count's histogram (yak-count.c:205-239) after its shrink
 * (:247-282, 453): hist[c] = entries with a count of c for min_count <= c <= max_count, 0
 * elsewhere; the reference prints rows 1..1023.
 * The filter takes one 64-bit word and one atomic per k-mer instead of the reference's 512-bit
 * block walked bit by bit: with one file the result does not depend on the filter at all (every
 * k-mer seen twice gets an entry whatever the filter's false positives, and entries seen once
 * are dropped by the shrink); with a second file it can differ from the reference's by the k-mers
 * that are false positives of one filter and not of the other, as the reference's own result
 * differs between two values of -b.
 */
#define KCGPU_PASS_COUNT 0
#define KCGPU_PASS_CLAIM 1
#define KCGPU_PASS_LOOKUP 2
int kcgpu_create_filtered(kcgpu_ctx **ctx, int k, uint64_t table_slots, uint64_t list_slots, size_t block_bytes, int device,
                          int bloom_bits, int bloom_hashes);
int kcgpu_set_pass(kcgpu_ctx *ctx, int pass);
int kcgpu_histogram1024(kcgpu_ctx *ctx, uint64_t hist[1024], int min_count, int max_count, kcgpu_stats *stats);

/*
 * Hand one parsed read to the engine: replaces the per-read copy of step 0 and steps 1 and 2
 * (kc-c4.c:133-180).  Reads shorter than k are dropped (kc-c4.c:141).  Bytes are classified by
 * the reference's strict table (kc-c4.c:21-38): A C G T U in either case and the bytes 0..3
 * are bases, everything else ends a k-mer.  One thread at a time (more: kcgpu_producer_*).
 */
int kcgpu_add_read(kcgpu_ctx *ctx, const char *seq, size_t len);

/*
 * Several reader threads: one producer each.  A producer packs its reads into a staging block
 * of its own (creating it adds one to the context) and submits the block when it is full, when
 * it is flushed and when it is destroyed; kcgpu_add_read is the context's built-in producer.
 * Everything else of a context may be called from any one thread at a time.
 */
int kcgpu_producer_create(kcgpu_ctx *ctx, kcgpu_producer **producer);
int kcgpu_producer_add_read(kcgpu_producer *producer, const char *seq, size_t len);
int kcgpu_producer_flush(kcgpu_producer *producer);
int kcgpu_producer_destroy(kcgpu_producer *producer);

/*
 * Count a stream the caller has already packed in HOST memory: reads separated by '\n', bytes
 * other than A C G T U (either case) end a k-mer (no byte translation is done: hand parsed reads
 * to kcgpu_add_read if they may hold the bytes 0..3).  Page-locked memory is copied to the
 * device as it is, block by block, each copy overlapping the previous block's kernel; pageable
 * memory goes through the pinned staging blocks.  Returns when everything is submitted: a
 * page-locked buffer is still being read by the copy engines then, and must stay valid and
 * unmodified until kcgpu_sync, kcgpu_flush or kcgpu_histogram has returned.
 */
int kcgpu_submit_stream(kcgpu_ctx *ctx, const char *bytes, size_t n_bytes);

/*
 * Count a stream that is already resident on the context's device: reads separated by '\n',
 * 16-byte aligned, n_bytes a multiple of 16; bytes other than A C G T U (either case) end a
 * k-mer; k-mers do not span calls.  Every k-mer is filed with its owner as named by
 * kcgpu_set_owners (default: this context).  `stream` is a cudaStream_t (NULL = the context's
 * own).  Asynchronous, except that the call flushes first (and waits for it) when the lists
 * could not take n_bytes more k-mers; a context whose owners were set by kcgpu_set_owners never
 * flushes by itself (see there).
 */
int kcgpu_count_device(kcgpu_ctx *ctx, const void *d_bytes, size_t n_bytes, void *stream);

/*
 * The two halves of kcgpu_count_device, for an exchange by collective: extract hashes every
 * canonical k-mer of the stream and files it under its owner, hash mod n_parts:
 * d_keys[part * cap_per_part + i], i < d_part_counts[part] (uint32, the caller zeroes them;
 * entries beyond cap_per_part are dropped, the count still runs on, and kcgpu_stats.n_dropped
 * reports them).  After the exchange, insert adds n hashed k-mers, all owned by `my_part` of
 * `n_parts`, to the context's table.  Both asynchronous on `stream`.
 */
int kcgpu_extract_device(kcgpu_ctx *ctx, const void *d_bytes, size_t n_bytes, int n_parts,
                         uint64_t *d_keys, size_t cap_per_part, uint32_t *d_part_counts, void *stream);
int kcgpu_insert_device(kcgpu_ctx *ctx, const uint64_t *d_hashed_keys, size_t n, int n_parts, void *stream);

/*
 * Several GPUs, fused form.  The allocation of this context (table, lists, cursors) as a device
 * pointer and as a CUDA IPC handle (KCGPU_IPC_HANDLE_BYTES bytes) for another process;
 * kcgpu_ipc_open maps a peer's allocation into this process.  kcgpu_set_owners names the
 * allocation of every owner (all created with the same k, table_slots and list_slots as this
 * context; tables[my_part] may be NULL = this context's own); from then on this context's
 * kernels file a k-mer with owner hash mod n_parts (what it filed before is flushed first: name
 * the owners before counting, or between two barriers).  The owners may live in other processes, so
 * such a context never flushes by itself: the caller calls kcgpu_flush on every owner, between
 * two barriers (nobody may be filing while lists are emptied), at the latest when the contexts
 * together have taken n_parts * kcgpu_stats.flush_bytes of stream since the last flush.  Lists
 * that fill up earlier cost speed, not k-mers: the excess goes straight to the table.
 * kcgpu_link does all of it for contexts of one process (peer access enabled both ways),
 * flushes the group when it is due, and makes kcgpu_sync / kcgpu_flush / kcgpu_histogram of any
 * member act on all of them.
 */
int kcgpu_table(kcgpu_ctx *ctx, void **d_table, uint64_t *table_slots);
int kcgpu_ipc_export(kcgpu_ctx *ctx, void *handle);
int kcgpu_ipc_open(kcgpu_ctx *ctx, const void *handle, void **d_peer_table);
int kcgpu_set_owners(kcgpu_ctx *ctx, int n_parts, int my_part, void *const *tables);
int kcgpu_link(kcgpu_ctx *const *ctxs, int n);

/* Submit what kcgpu_add_read has staged and wait for every kernel of this context (and of the
 * contexts linked to it).  Across processes the caller adds its own barrier. */
int kcgpu_sync(kcgpu_ctx *ctx);

/* kcgpu_sync, then empty the region lists of this context (and of the contexts linked to it)
 * into the table(s), and wait for that. */
int kcgpu_flush(kcgpu_ctx *ctx);

/*
 * kcgpu_flush, then the histogram of kc-c4.c:186-215 over this context's table: hist[c] = number
 * of distinct canonical k-mers seen min(c, 255) times, counts saturating at 1023 first as in
 * kc-c4.c:125 (hist[0] is 0).  With several owners the per-context histograms are summed by
 * the caller (they partition the k-mers).  Counting can go on afterwards.
 */
int kcgpu_histogram(kcgpu_ctx *ctx, uint64_t hist[256], kcgpu_stats *stats);

/* Empty the table and the statistics (asynchronous on the context's stream). */
int kcgpu_reset(kcgpu_ctx *ctx);

void kcgpu_destroy(kcgpu_ctx *ctx);
const char *kcgpu_strerror(const kcgpu_ctx *ctx);

/* the reference's hash (kc-c4.c:40-50) on the host, exported for tests */
uint64_t kcgpu_hash64(uint64_t key, int k);

#ifdef __cplusplus
}
#endif
#endif
