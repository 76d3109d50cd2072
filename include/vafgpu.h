/*
 * vafgpu.h -- C ABI of the B200 k-mer extract-and-lookup engine behind vaf-counter.
 *
 * The reference (gerbenvoshol/kmer-cnt) has no plugin or FFI interface; the seam this
 * library replaces is internal to vaf-counter.c: everything between "a parsed read is
 * available" (step 0 of worker_pipeline, vaf-counter.c:486-517) and "the per-pattern
 * ref/alt counters are final" (vaf-counter.c:654).  Each entry point names the reference
 * lines it stands in for.  Plain C types only; no exceptions cross the boundary; every
 * function that can fail returns 0 on success or a negative VAFGPU_E* code, with a
 * message available from vafgpu_strerror().  There is no CPU fallback: without a usable
 * sm_100 device vafgpu_create() fails with VAFGPU_ENOGPU.
 *
 * Encoding of k-mer keys at this boundary is the reference's: A=0 C=1 G=2 T=3, first
 * base in the most significant position, canonical = min(forward, reverse complement)
 * (vaf-counter.c:117-146).
 */
#ifndef VAFGPU_H
#define VAFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAFGPU_OK       0
#define VAFGPU_EINVAL  -1   /* bad argument                                  */
#define VAFGPU_ENOGPU  -2   /* no usable CUDA device / not an sm_100 device  */
#define VAFGPU_ECUDA   -3   /* a CUDA runtime call failed                    */
#define VAFGPU_ENOMEM  -4   /* host or device allocation failed              */
#define VAFGPU_ENCCL   -5   /* NCCL could not be loaded or a call failed     */
#define VAFGPU_ESTATE  -6   /* call not valid in the current state           */

/* vafgpu_create flags */
#define VAFGPU_F_REFERENCE_RECIPE 1u /* run the literal recipe kernel (rolling forward and
                                        reverse-complement words, canonical minimum, the
                                        khashl hash and probe order of vaf-counter.c:56-63,
                                        349-427, khashl.h:98,137-150) instead of the
                                        anchor-filter kernel; used as the on-device
                                        verification mode */
#define VAFGPU_F_HOST_MERGE       2u /* merge per-device counters through host memory instead
                                        of over NVLink (peer access is not required then)  */
#define VAFGPU_F_STRICT_BYTES     4u /* vafgpu_add_read classifies every byte by the strict table
                                        (vaf-counter.c:73-90 = snp-pattern-gen.c:30-47): what a
                                        build without SSSE3 does, and what snp-pattern-gen's
                                        genome scan does (snp-pattern-gen.c:162-190) */

typedef struct vafgpu_ctx vafgpu_ctx;
typedef struct vafgpu_producer vafgpu_producer;

typedef struct vafgpu_stats {
	uint64_t n_reads;       /* reads accepted by vafgpu_add_read (len >= k)               */
	uint64_t n_bases;       /* their bases: "Bases processed" of vaf-counter.c:506,700     */
	uint64_t n_blocks;      /* device launches                                             */
	uint64_t n_bytes;       /* bytes scanned on the devices (bases + separators + padding) */
	uint64_t n_candidates;  /* anchors that passed the on-chip filter (0 in recipe mode)   */
	uint64_t n_hits;        /* k-mer occurrences counted (sum of all counters' increments) */
	uint64_t n_kmers;       /* valid k-mers seen; only maintained in recipe mode           */
	double   kernel_ms;     /* sum of kernel durations over all devices (CUDA events)      */
	double   h2d_ms;        /* sum of host-to-device copy durations                        */
	int      n_devices;
	int      anchor_stride; /* S: one anchor every S bases                                 */
	int      anchor_len;    /* L: anchor length in bases                                   */
	uint32_t filter_bytes;  /* shared-memory filter size                                   */
	uint32_t table_slots;   /* slots of the L2-resident exact table                        */
	/* which form of the anchor kernel this panel selected (all 0 in recipe mode) */
	int      filter_canon;    /* 1: strand-symmetric filter keys (large panels)              */
	int      lookup_deferred; /* 1: deferred two-level lookup (large panels, stride >= 4)    */
	int      kernel_threads;  /* CTA size of the instantiation                               */
	uint32_t filter2_bytes;   /* second filter level (L2-resident), 0 if not in use          */
} vafgpu_stats;

/*
 * Build an engine.  Replaces kt_pipeline()/kt_for() set-up in count_fastq_kmers
 * (vaf-counter.c:550-568) and takes the product of create_combined_kmer_map
 * (vaf-counter.c:198-252) as a flat list, so that the reference's first-insert-wins rule
 * lives in exactly one place, the caller:
 *   canon_keys[i]  canonical k-mer, vals[i] = (pattern index << 1) | is_alt, i < n_entries;
 *   if a key occurs twice the first occurrence is kept.
 *   n_patterns     the counter vector has 2*n_patterns words ([2i] = ref, [2i+1] = alt).
 *   block_bytes    size of one pinned staging block (the CLI passes -b); 0 = 16 MiB.
 *   n_buffers      staging blocks (and streams) per device, >= 2 for copy/compute overlap;
 *                  0 = 3 (the reference keeps at most three blocks in flight).
 *   n_devices      devices to use starting at the current CUDA device 0; 0 = all visible.
 * The tables are replicated on every device.
 */
int vafgpu_create(vafgpu_ctx **ctx, int k, const uint64_t *canon_keys, const uint32_t *vals,
                  uint32_t n_entries, uint32_t n_patterns, size_t block_bytes, int n_buffers,
                  int n_devices, unsigned flags);

/*
 * Hand one parsed read to the engine.  Replaces the per-read malloc+memcpy of step 0
 * (vaf-counter.c:494-507), the SSSE3/scalar encoder's byte classification
 * (vaf-counter.c:261-291,73-90) and, by submitting a staging block whenever it fills,
 * steps 1 and 2 (vaf-counter.c:519-544).  Reads shorter than k are dropped and not
 * counted in the statistics, as in vaf-counter.c:494.  Bytes are canonicalised to
 * {A,C,G,T,N} while they are copied: for offsets below (len & ~15) by the reference's
 * low-nibble rule, for the tail by its strict table, so that the device sees exactly the
 * bases the Makefile-built reference sees.  vafgpu_add_read and vafgpu_submit_stream share
 * one implicit producer: call them from one thread at a time; reader threads that run side
 * by side each use a producer of their own (below).
 */
int vafgpu_add_read(vafgpu_ctx *ctx, const char *seq, size_t len);

/*
 * Parallel ingest: one producer per reader thread (one per input file or file slice).
 * Replaces the single kseq reader of step 0 (vaf-counter.c:486-517; the reference parses
 * with one thread, which bounds it end to end).  A producer packs its reads into a staging
 * block of its own, taken from the context's pool (a producer waits while every block is
 * being filled or is in flight: that back-pressure stands in for kt_pipeline's "at most
 * three blocks", kthread.c:130-159); blocks go to the devices round-robin in the order
 * they are taken.  Counting is additive, so the result does not depend on how reads are
 * spread over producers.  Each producer is used by one thread at a time; different
 * producers may be used concurrently.  vafgpu_producer_flush submits the partly filled
 * block and adds the producer's read/base totals to the context's statistics; every
 * producer must be flushed (or destroyed) before vafgpu_finish is called.
 */
int vafgpu_producer_create(vafgpu_ctx *ctx, vafgpu_producer **producer);
int vafgpu_producer_add_read(vafgpu_producer *producer, const char *seq, size_t len);
int vafgpu_producer_flush(vafgpu_producer *producer);
int vafgpu_producer_destroy(vafgpu_producer *producer); /* flushes first */

/*
 * Submit a host buffer that is already in stream form: reads separated by '\n', bytes
 * other than A,C,G,T,U (either case) end a k-mer.  The buffer is copied to a staging
 * block (or used in place if it is pinned) and processed asynchronously; n_bytes may
 * exceed block_bytes, in which case it is cut at separators.  n_reads/n_bases are only
 * added to the statistics.  A page-locked buffer is read by the copy engines AFTER this
 * call returns: it must stay valid and unmodified until vafgpu_finish() (or vafgpu_reset())
 * has returned.  Pageable memory has been copied when the call returns.
 */
int vafgpu_submit_stream(vafgpu_ctx *ctx, const char *bytes, size_t n_bytes,
                         uint64_t n_reads, uint64_t n_bases);

/*
 * Scan a stream that is already resident in device memory (same format; d_bytes 16-byte
 * aligned, n_bytes a multiple of 16) on the given device ordinal, on `stream`
 * (a cudaStream_t passed as void*, NULL = the engine's own), adding into `d_counts`
 * (2*n_patterns uint32 on that device, NULL = the engine's own counters).  Asynchronous.
 * This is the kernel-only entry point bench.py times.
 */
int vafgpu_count_device(vafgpu_ctx *ctx, int device, const void *d_bytes, size_t n_bytes,
                        uint32_t *d_counts, void *stream);

/*
 * Drain all streams and copy the counters out.  With several devices there is nothing to
 * merge: every device's kernels add their hits straight into device 0's vector through
 * NVLink peer memory (hits are rare and the additions need no answer), which stands in
 * for the all-reduce of a one-vector-per-device design; devices without peer access or
 * native peer atomics (or VAFGPU_F_HOST_MERGE) keep a vector each, summed here.  Replaces
 * the end of kt_pipeline plus the shared-memory atomics of worker_lookup
 * (vaf-counter.c:473-477): counts[2i] / counts[2i+1] are what the reference leaves in
 * pattern_t.ref_count / alt_count.  Counters keep accumulating across calls (several
 * input files, vaf-counter.c:647-650) until vafgpu_reset().
 */
int vafgpu_finish(vafgpu_ctx *ctx, uint32_t *counts, vafgpu_stats *stats);

int vafgpu_reset(vafgpu_ctx *ctx);   /* zero counters and statistics */

/*
 * One counter vector for several PROCESSES (one process per GPU, e.g. under torchrun):
 * the owner exports a CUDA IPC handle of its vector (handle_bytes >= 64), the others attach
 * it, after which their kernels add into the owner's vector over NVLink and their own
 * vafgpu_finish() only drains (it returns zeros); the owner reads the totals with
 * vafgpu_finish() once every process has drained (a barrier of the caller's).  The devices
 * must be NVLink peers with native atomics.  Stands in for the all-reduce across ranks.
 */
int vafgpu_export_counters(vafgpu_ctx *ctx, void *handle, size_t handle_bytes);
int vafgpu_attach_counters(vafgpu_ctx *ctx, const void *handle, size_t handle_bytes);
void vafgpu_destroy(vafgpu_ctx *ctx);

/* message for the last error on ctx (ctx may be NULL: error of the last failed create) */
const char *vafgpu_strerror(const vafgpu_ctx *ctx);

/* anchor plan the engine uses for a given k: one anchor of *len bases every *stride bases */
int vafgpu_plan(int k, int *stride, int *len);

/* host-side helper used by vafgpu_add_read, exported for tests: canonicalise one read
 * (rules above) into out[0..len).  simd_rule = 0 applies the strict table everywhere
 * (the reference built without SSSE3, vaf-counter.c:341-343). */
void vafgpu_canonicalise_read(const char *seq, size_t len, char *out, int simd_rule);

const char *vafgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
