/*
 * seekmer_b200.h — C ABI of the B200-native Seekmer bulk-infer hot path.
 *
 * The reference (GuanLab/seekmer) has no FFI layer: its native code is Cython
 * compiled into the Python package, and the boundary of the hot path is the
 * Python-level API of `seekmer._mapper.ReadMapper`, `seekmer.mapper` and
 * `seekmer.infer`.  This header is what a ctypes (or Cython
 * `cdef extern`) binding on the reference side would bind instead; the stub is
 * shown in INTEGRATION.md.  Every entry point names the reference interface it
 * replaces (paths relative to /root/reference/seekmer/).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy / Python types.
 *   - every function returns 0 on success, a negative skm_status on failure;
 *     skm_last_error() returns a per-thread message for the last failure.
 *   - handles own their device memory; callers own every buffer they pass.
 *   - a handle is bound to one device and is not thread-safe; use one mapper
 *     per host thread / stream (the reference's one ReadMapper per thread,
 *     mapper.py:174-182).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Calls with host buffers synchronise the stream before returning; calls
 *     with device buffers only enqueue work.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with SKM_ERR_CUDA.
 */
#ifndef SEEKMER_B200_H
#define SEEKMER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKM_KMER_SIZE 25             /* _kmer.pxd:9-17 */
#define SKM_MAX_FRAGMENT_LENGTH 2000 /* _mapper.pyx:18 */

typedef enum skm_status {
    SKM_OK = 0,
    SKM_ERR_INVALID = -1,   /* bad argument */
    SKM_ERR_CUDA = -2,      /* CUDA runtime / no device */
    SKM_ERR_CAPACITY = -3,  /* class table, id pool or list arena exhausted */
    SKM_ERR_OOM = -4,       /* device allocation failed */
    SKM_ERR_COLLISION = -5  /* two different classes share a 128-bit dictionary key (p < 1e-24); nothing merged silently */
} skm_status;

typedef struct skm_index skm_index;
typedef struct skm_mapper skm_mapper;
typedef struct skm_em_plan skm_em_plan;

/* ---- index input contract: the arrays held by `KMerIndex` ------------------
 * (_common.pxd:15-35,57-66; dtypes as produced by `seekmer index`,
 *  _index_builder.pyx:129-141,551-571) */
typedef struct skm_kmer_slot {   /* `kmers`: open addressing, linear probing */
    uint64_t kmer;               /* contig-forward 2-bit 25-mer; all-ones = empty */
    int32_t entry;               /* contig id */
    int32_t offset;              /* k-mer position inside the contig */
} skm_kmer_slot;

typedef struct skm_contig_entry { /* `contigs` */
    int64_t offset;              /* into `sequences` */
    int64_t length;
    uint64_t first_kmer;
    uint64_t last_kmer;
    int64_t target_offset;       /* into `targets` */
    int64_t target_count;
} skm_contig_entry;

typedef struct skm_target {      /* `targets`: sorted by signed (entry, offset) per contig */
    int32_t entry;               /* transcript t, or ~t when the contig is reverse in t */
    int32_t offset;
} skm_target;

const char *skm_last_error(void);
int skm_device_count(void);
const char *skm_version(void);

/* Replaces: KMerIndex.__init__ / KMerIndex.load (_common.pyx:21-48,287-313).
 * Re-lays the index out for the GPU once: canonical-key open-addressing table
 * (load <= 0.25, 16-byte slots in 64-byte buckets), 128-byte contig records
 * (header, 8 inline targets, graph links), 2-bit packed contig sequences,
 * entry-only int32 target lists.  `inputs_on_device` != 0 means the
 * four arrays are device pointers on `device` (reference layout); they are
 * only read. */
int skm_index_create(const skm_kmer_slot *kmers, int64_t n_slots,
                     const skm_contig_entry *contigs, int64_t n_contigs,
                     const char *sequences, int64_t n_bases,
                     const skm_target *targets, int64_t n_targets,
                     int64_t n_transcripts, int device, int inputs_on_device,
                     void *stream, skm_index **out);
void skm_index_destroy(skm_index *index);

/* Replaces: KMerIndex.save / KMerIndex.load (_common.pyx:268-313) for the device-side form of
 * the index (the native GPU-layout index file): the re-laid-out table, the
 * contig records with their graph links, the 2-bit sequences and the target entries are written
 * as they lie in HBM and read back with plain copies - loading runs no relayout kernel and
 * none of the 8 table probes per contig.  `trailer` is caller-owned bytes stored after the
 * image (the Python layer keeps the transcript table there); skm_index_load reports where they
 * are.  A file written by another layout version is refused ("invalid index version.",
 * _common.pyx:303-304). */
int skm_index_save(const skm_index *index, const char *path, const void *trailer,
                   int64_t trailer_bytes, void *stream);
int skm_index_load(const char *path, int device, void *stream, skm_index **out,
                   int64_t *trailer_offset, int64_t *trailer_bytes);

/* info[0]=distinct k-mers, [1]=device table slots, [2]=max target_count,
 * [3]=device bytes held, [4]=n_contigs, [5]=n_targets, [6]=n_transcripts, [7]=device */
int skm_index_info(const skm_index *index, int64_t info[8]);

/* Replaces: KMerIndex.map_kmer (_common.pyx:54-97) for a vector of k-mers
 * (parity test (i) of SURVEY.md §8(c)).  Miss = {entry 0, offset -1}. */
int skm_map_kmers(const skm_index *index, const uint64_t *kmers, int64_t n,
                  int32_t *out_entry, int32_t *out_offset, int buffers_on_device,
                  void *stream);

/* Replaces: ReadMapper.__init__ + MapResult.__init__ state
 * (_mapper.pyx:39-53, mapper.py:43-58): a device-resident class dictionary
 * (ordered transcript-id tuple -> count, first-seen unit index), the FLD
 * histogram and the unaligned counter.  class_capacity = max distinct classes
 * (0 = default 1<<22), id_capacity = total ids over all classes (0 = 8x). */
int skm_mapper_create(skm_index *index, int64_t class_capacity, int64_t id_capacity,
                      skm_mapper **out);
void skm_mapper_destroy(skm_mapper *mapper);
/* Replaces: MapResult.clear + fresh FLD (mapper.py:143-145). */
int skm_mapper_reset(skm_mapper *mapper, void *stream);

/* Replaces: one iteration of ReadMapper.__call__ (_mapper.pyx:73-101):
 * map_read / map_read_pair on every unit of the batch, FLD update (:90-94),
 * MapResult.update (mapper.py:60-75) into the device dictionary.
 *
 *   bases         all reads concatenated, one ASCII byte per base (any byte
 *                 other than ACGTacgt encodes as 'A' for k-mers and is a
 *                 wildcard for the 8-base edge checks, _kmer.pxd:253-273,
 *                 _mapper.pyx:500-501)
 *   read_offsets  n_reads+1 int64 offsets into `bases`, or NULL when every read
 *                 has `fixed_read_len` bases
 *   n_units       reads (single-ended) or pairs; mates are interleaved
 *                 (2i, 2i+1) as the feeders produce them (common.py:189-190)
 *   first_unit    global index of unit 0 (keeps first-seen class order across
 *                 batches and shards)
 *   out_class     optional int32[n_units]: dictionary slot of the unit's class,
 *                 -1 = unaligned  (for -m readmap and per-read parity tests)
 *   out_length    optional int32[n_units]: span.end - span.begin + k (:90)
 * Reads shorter than k are undefined in the reference (_kmer.pxd:46-68 reads past
 * their end).  Here they are legal input (trimmed FASTQ): a unit with such a read is
 * reported unaligned with span length 0, and such units are counted (skm_classes_size).
 * Reads longer than 4096 bases are refused with SKM_ERR_INVALID.
 */
int skm_map_batch(skm_mapper *mapper, const uint8_t *bases, const int64_t *read_offsets,
                  int32_t fixed_read_len, int32_t max_read_len, int64_t n_units, int paired,
                  int64_t first_unit, int buffers_on_device, int32_t *out_class,
                  int32_t *out_length, void *stream);

/* Replaces: the line loops of common.feed_single_ended_reads / feed_pair_ended_reads
 * (common.py:126-197) plus skm_map_batch: maps the reads of raw FASTQ TEXT.  text1 (and text2
 * for mates; NULL = single-ended) must start at a record boundary; as many whole 4-line records
 * as both chunks hold are parsed on the device (newline scan; line 4r+1 = sequence of record r,
 * blanks stripped at both ends), packed straight out of the text and mapped.  consumedN = bytes
 * of textN that were used: the caller prepends the rest to the next chunk.  Chunks < 2 GiB.
 * out_class / out_length as in skm_map_batch (sized for the units the chunks can hold). */
int skm_map_fastq(skm_mapper *mapper, const uint8_t *text1, int64_t n1, const uint8_t *text2,
                  int64_t n2, int64_t first_unit, int buffers_on_device, int64_t *consumed1,
                  int64_t *consumed2, int64_t *n_units, int32_t *out_class, int32_t *out_length,
                  void *stream);

/* Measurement support: device durations (ms, CUDA events on the launch stream) of the three
 * kernels of the most recent chunk mapped by skm_map_batch: ms[0] = pack_reads_kernel,
 * ms[1] = map_reads_kernel, ms[2] = tally_units_kernel.  Blocks until that chunk is done. */
int skm_mapper_kernel_ms(skm_mapper *mapper, double ms[3]);

/* Diagnostics of map_reads_kernel, filled only by a library built with -DSKM_STATS=1
 * (otherwise SKM_ERR_INVALID): for each of the six phases [iterations, items claimed, of them
 * borrowed from another lane's column, clock cycles], then the number of idle polls.  No
 * reference counterpart; used by tools/sweep_map.py. */
int skm_debug_map_stats(uint64_t stats[32], int reset);

/* sizes[0]=n_classes, [1]=total ids, [2]=unaligned units, [3]=aligned units,
 * [4]=class slots capacity, [5]=status flags raised on device (0 = none),
 * [6]=units with a read shorter than k (they are part of [2]), [7]=id-pool cursor */
int skm_classes_size(skm_mapper *mapper, int64_t sizes[8], void *stream);

/* Replaces: reading MapResult.counter / fragment_length_counts
 * (mapper.py:54-58) — the dictionary compacted to CSR.  Classes come out in
 * table order; sort by first_unit for the reference's insertion order at
 * job_count=1.  Any pointer may be NULL.  fld has SKM_MAX_FRAGMENT_LENGTH entries. */
int skm_classes_export(skm_mapper *mapper, int64_t *key_offsets, int32_t *key_ids,
                       int64_t *counts, int64_t *first_unit, int32_t *slots, int64_t *fld,
                       int buffers_on_device, void *stream);

/* Replaces: the cross-thread merge under MapResult.lock (_mapper.pyx:100-105)
 * for the multi-GPU case: add externally produced classes (CSR, counts,
 * first_unit), FLD and unaligned count into this mapper's dictionary. */
int skm_classes_merge(skm_mapper *mapper, const int64_t *key_offsets, const int32_t *key_ids,
                      const int64_t *counts, const int64_t *first_unit, int64_t n_classes,
                      const int64_t *fld, int64_t unaligned, int buffers_on_device,
                      void *stream);

/* The same for every peer of a multi-GPU exchange at once (device buffer): `gathered` is the
 * all-gather of one packed export per rank, words_per_rank int64 words apart, each laid out as
 * [n_classes, n_ids, unaligned | fld[2000] | key_offsets[n+1] | counts[n] | first_unit[n] |
 * key_ids (int32)]; the block of `rank` itself is skipped. */
int skm_classes_merge_packed(skm_mapper *mapper, const int64_t *gathered, int64_t words_per_rank,
                             int world, int rank, void *stream);

/* Where the EM's scratch blocks come from.  on = 1: the virtual-memory-management API with
 * access for the owning device only - for a process that has peer access to other GPUs switched
 * on (NCCL, peer copies), where every cudaMalloc also maps the block into the peers (100+ ms per
 * GB); on = 0 (default): cudaMalloc.  No reference counterpart (the reference's scratch is numpy).
 * Returns the previous setting. */
int skm_scratch_local_only(int on);

/* The EM entry points keep their large scratch blocks cached per device between calls (a
 * cudaMalloc/cudaFree pair per call cost more than the EM); at most 8 GiB stay idle, and the
 * cache is emptied before an allocation is reported as failed.  This gives every idle cached
 * block back to the driver (e.g. before another library needs the memory); *freed_bytes may
 * be NULL.  No reference counterpart (the reference's scratch is malloc/free inside the call). */
int skm_release_cache(int device, int64_t *freed_bytes);

/* Replaces: MapResult.effective_lengths (mapper.py:134-141), fp64, same
 * accumulation order. */
int skm_effective_lengths(const int64_t *fld, const double *lengths, int64_t n_transcripts,
                          double *out, int buffers_on_device, int device, void *stream);

/* Replaces: infer.em (infer.py:133-168) for n_replicates independent count
 * vectors sharing one class structure (E2/E4 of SURVEY.md §8(a)).
 *   class_ptr[n_classes+1], class_tx[nnz]  CSR by class, ids in tuple order
 *   counts[n_replicates * n_classes]       fp64 class counts per replicate
 *   eff_len[n_transcripts]
 *   x0[n_replicates * n_transcripts]       initial guess (already normalised)
 *   out_x same shape; out_iters[n_replicates] = EM iterations executed
 * Convergence: max over {x_t > 1e-8} |x_t - old_t| / x_t <= 0.01 (infer.py:160);
 * a replicate with no x_t > 1e-8 stops (the reference raises ValueError). */
int skm_em(const int64_t *class_ptr, const int32_t *class_tx, int64_t n_classes, int64_t nnz,
           const double *counts, const double *eff_len, int64_t n_transcripts,
           const double *x0, int64_t n_replicates, int64_t max_iters, double *out_x,
           int32_t *out_iters, int buffers_on_device, int device, void *stream);

/* ---- EM plans: the class structure resident on one device, built once -----------------------
 * Replaces: the `class_map` / `class_count` arrays that SummarizedResult carries from
 * MapResult.summarize (mapper.py:77-104) into every quantify call (infer.py:88-130).  A plan
 * holds the class x transcript structure in both orders (CSR by class, CSC by transcript); the
 * main EM, the bootstrap replicates (infer.py:79-82) and any repeated call share it.
 *   skm_em_plan_create       from CSR arrays (host or device); `counts` (int64[n_classes], may be
 *                            NULL) makes the plan own the integer class counts
 *   skm_em_plan_from_mapper  from a mapper's dictionary where it lies in HBM: classes in
 *                            first-seen order (the Counter order at job_count=1), ids in tuple
 *                            order, counts included; n_transcripts <= 0 = the index's
 *   skm_em_plan_info         info[0]=n_classes [1]=nnz [2]=n_transcripts [3]=device [4]=owns counts
 *   skm_em_plan_run          skm_em on the plan; counts NULL = the plan's own (n_replicates 1)
 *   skm_em_plan_bootstrap    skm_em_bootstrap on the plan; counts NULL = the plan's own */
int skm_em_plan_create(const int64_t *class_ptr, const int32_t *class_tx, int64_t n_classes,
                       int64_t nnz, int64_t n_transcripts, const int64_t *counts,
                       int buffers_on_device, int device, void *stream, skm_em_plan **out);
int skm_em_plan_from_mapper(skm_mapper *mapper, int64_t n_transcripts, void *stream,
                            skm_em_plan **out);
int skm_em_plan_info(const skm_em_plan *plan, int64_t info[5]);
void skm_em_plan_destroy(skm_em_plan *plan);
int skm_em_plan_run(const skm_em_plan *plan, const double *counts, const double *eff_len,
                    const double *x0, int64_t n_replicates, int64_t max_iters, double *out_x,
                    int32_t *out_iters, int buffers_on_device, void *stream);
int skm_em_plan_bootstrap(const skm_em_plan *plan, const int64_t *counts, const double *eff_len,
                          const double *x0, int64_t n_replicates, int64_t first_replicate,
                          uint64_t seed, int method, int64_t max_iters, int tpm, double *out_x,
                          int32_t *out_iters, int buffers_on_device, void *stream);

/* Replaces: scipy.stats.multinomial(n, p).rvs() at infer.py:108-111 — resample
 * n = sum(counts) reads with replacement, n_replicates times; every replicate sums to n.
 * Counter-based (Philox4x32-10 keyed by seed), so replicate r is a function of (seed, r) alone:
 * replicate ids start at first_replicate and shards reproduce the single-call rows.  `method`:
 *   SKM_RESAMPLE_TREE   O(n_classes) per replicate whatever the read depth: a binary tree of
 *                       exact binomial splits (inversion / BTPE, csrc/binomial.cuh), one Philox
 *                       stream per (tree node, replicate).  What the bootstraps use.
 *   SKM_RESAMPLE_DRAWS  O(n) per replicate: n categorical draws with integer arithmetic only
 *                       (counter = (draw pair, replicate); one Philox block yields two 64-bit
 *                       draws); bit-identical to the numpy restatement in oracle/oracle.py.
 * out is int64[n_replicates * n_classes]. */
#define SKM_RESAMPLE_DRAWS 0
#define SKM_RESAMPLE_TREE 1
int skm_multinomial(const int64_t *counts, int64_t n_classes, int64_t n_replicates,
                    int64_t first_replicate, uint64_t seed, int method, int64_t *out,
                    int buffers_on_device, int device, void *stream);

/* Replaces: the bootstrap loop of infer.run (infer.py:79-82), i.e. n_replicates times
 * quantify(results, x0=main, bootstrap=True): resample the integer class counts (as
 * skm_multinomial, same counters), run the EM for every replicate from the one normalised
 * guess x0[n_transcripts] and, when `tpm` != 0, apply the TPM post-processing of
 * infer.py:127-129 - all on the device; the resampled counts never cross PCIe.
 * out_x is [n_replicates][n_transcripts]. */
int skm_em_bootstrap(const int64_t *class_ptr, const int32_t *class_tx, int64_t n_classes,
                     int64_t nnz, const int64_t *counts, const double *eff_len,
                     int64_t n_transcripts, const double *x0, int64_t n_replicates,
                     int64_t first_replicate, uint64_t seed, int method, int64_t max_iters, int tpm,
                     double *out_x, int32_t *out_iters, int buffers_on_device, int device,
                     void *stream);

/* Replaces: `[infer.quantify(r) for r in map_results]` of the single-cell workflow
 * (impute.py:101, mapper.py:196-234 samples) - n_samples independent EMs (infer.py:133-168),
 * each with its OWN class structure, in one set of launches.  The samples' structures are
 * laid end to end:
 *   class_ptr[n_classes+1], class_tx[nnz]   CSR by class over all samples; class_tx holds the
 *                                            transcript index within the class's own sample
 *   sample_class_ptr[n_samples+1]           first class of every sample (0 ... n_classes)
 *   counts[n_classes]                       fp64 class counts
 *   eff_len, x0, out_x                      [n_samples][n_transcripts]
 *   out_iters[n_samples]                    EM iterations each sample executed
 * Every sample stops by its own convergence test and is then carried through unchanged;
 * results are bit-identical to one skm_em call per sample. */
int skm_em_samples(const int64_t *class_ptr, const int32_t *class_tx,
                   const int64_t *sample_class_ptr, int64_t n_samples, int64_t n_classes,
                   int64_t nnz, const double *counts, const double *eff_len,
                   int64_t n_transcripts, const double *x0, int64_t max_iters, double *out_x,
                   int32_t *out_iters, int buffers_on_device, int device, void *stream);

/* The same for samples whose class structures are already on the device as PLANS (one per
 * sample, skm_em_plan_from_mapper: the mapper.py:196-234 / impute.py:101 flow without the host
 * round trip of class_map): n_plans independent EMs in one set of launches.  All plans sit on
 * one device, share n_transcripts and own their counts.  eff_len, x0, out_x: [n_plans][T];
 * out_iters[n_plans].  Bit-identical to one skm_em_plan_run per plan. */
int skm_em_plans_run(const skm_em_plan *const *plans, int64_t n_plans, const double *eff_len,
                     const double *x0, int64_t max_iters, double *out_x, int32_t *out_iters,
                     int buffers_on_device, void *stream);

/* Workload generation twin of seekmer_b200/synth.py (bench/test support, not
 * part of the reference surface): fills `bases` (device) with ASCII reads for
 * global units [first_unit, first_unit+n_units). */
int skm_synth_reads(const uint8_t *tx_codes, const int64_t *tx_offsets, int64_t n_transcripts,
                    const uint64_t *cum_weights, uint64_t total_weight, int32_t read_len,
                    int32_t frag_mean, int32_t frag_sd, int32_t sub_thresh, int32_t n_thresh,
                    int32_t random_pct, uint64_t seed, int paired, int64_t first_unit,
                    int64_t n_units, uint8_t *bases, int device, void *stream);

/* Index-construction support (device buffers only): place n contig-forward k-mers with
 * their positions into `table` (n_slots, power of two) in the reference layout — home slot
 * from the reference hash of the canonical k-mer (_kmer.pxd:174-219), linear probing
 * (_index_builder.pyx:313-342).  Used by seekmer_b200/index_build.py. */
int skm_build_kmer_table(const uint64_t *kmers, const int32_t *entry, const int32_t *offset,
                         int64_t n, skm_kmer_slot *table, int64_t n_slots, int device,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SEEKMER_B200_H */
