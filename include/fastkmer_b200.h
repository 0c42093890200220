/* fastkmer_b200.h — C ABI of the B200-native exact k-mer counting path.
 *
 * This is the drop-in boundary for fastkmer's hot path.  The reference has no
 * FFI; its seam is the single call
 *     SparkBinKmerCounter.executeJob(spark, configuration)
 *         src/main/scala/skc/SparkBinKmerCounter.scala:989   (SBKC:989)
 * made by skc.test.LocalTestKmerCounter (LTKC:76) and skc.test.TestKmerCounter
 * (TKC:74).  A JVM host replaces the body of executeJob by one JNI call into
 * fkm_execute_job (see INTEGRATION.md for the Scala/JNI stub).
 *
 * Conventions: every function returns FKM_OK (0) or a negative FKM_E* code and
 * never throws or aborts across the ABI; fkm_last_error() gives the message of
 * the calling thread's last failure.  No torch / C++ types in any signature.
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with FKM_ECUDA.
 *
 * Shorthands in the citations below:
 *   SBKC = src/main/scala/skc/SparkBinKmerCounter.scala
 *   UTIL = src/main/scala/skc/package.scala
 *   TCFG = src/main/scala/skc/test/package.scala
 *   LTKC = src/main/scala/skc/test/LocalTestKmerCounter.scala
 */
#ifndef FASTKMER_B200_H
#define FASTKMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FKM_OK        0
#define FKM_EINVAL   -1   /* bad argument / unsupported configuration            */
#define FKM_ECUDA    -2   /* CUDA runtime failure or no device                   */
#define FKM_EIO      -3   /* file could not be read / written                    */
#define FKM_ENOMEM   -4   /* host or device allocation failed                    */
#define FKM_EOVERFLOW -5  /* a 32-bit count overflowed (reference counts are Int, SBKC:562,676) */

typedef struct fkm_ctx fkm_ctx;        /* one per (process, GPU): device, stream, scratch */
typedef struct fkm_result fkm_result;  /* per-bin (canonical k-mer, count) arrays          */

/* Mirrors skc.test.testutil.TestConfiguration (TCFG:16-30); field meaning and
 * the derived values are the reference's.  Strings are borrowed for the call. */
typedef struct fkm_config {
    int32_t k;                       /* k-mer length, m <= k <= 64                      */
    int32_t m;                       /* signature length, 3 <= m <= 15 (UTIL:78 Int)    */
    int32_t x;                       /* (k,x)-mer compression; must be >= 1 when use_ht=0 (SBKC:495 crashes on 0) */
    int32_t max_b;                   /* requested bins; b = min(4^m, max_b) (TCFG:32)   */
    int32_t sequence_type;           /* 0 short reads, 1 long sequence (SBKC:1010)      */
    int32_t use_ht;                  /* 1: hash-table count (SBKC:664), 0: sort count (SBKC:428) */
    int32_t write;                   /* 1: write <outputDir>/bin<id> files              */
    int32_t use_kryo_serializer;     /* accepted, ignored (JVM serialisation only)      */
    int32_t use_custom_partitioner;  /* accepted; bin->GPU ownership is always exact-histogram based */
    int32_t num_partition_tasks;
    const char* dataset;             /* FASTA path (fkm_execute_job only)               */
    const char* output_directory;
    const char* prefix;
} fkm_config;

/* Counters of one job.  n_* are exact.  *_ref are measured under the GPU's own
 * super-k-mer chunking (the reference's cutting rule is not observable).       */
typedef struct fkm_stats {
    uint64_t n_positions;        /* bases + one separator per record, as laid out on device */
    uint64_t n_bases;            /* sequence bytes of all records (invalid ones included)   */
    uint64_t n_kmers;            /* valid k-windows                                          */
    uint64_t n_superkmers;       /* super-k-mer records scattered into bins                  */
    uint64_t superkmer_bytes;    /* bytes of that intermediate                               */
    uint64_t n_distinct;         /* distinct (bin, canonical k-mer)                          */
    uint64_t total_count;        /* sum of counts == n_kmers                                 */
    uint64_t digest_sum;         /* order-independent digests of (bin, k-mer, count)         */
    uint64_t digest_xor;
    uint64_t n_nonempty_bins;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t gpu_launches;       /* kernels launched by this job                             */
    uint64_t n_batches;
    uint64_t n_fallbacks;        /* times the asynchronous count phase had to be redone synchronously */
    double   ms_total;           /* host wall time of the call                               */
    double   ms_stage[8];        /* 0 parse/pack+H2D 1 histogram 2 scatter 3 count 4 compact/reduce 5 digest 6 D2H/write 7 whole device pipeline (CUDA events) */
    uint64_t n_folded_records;   /* records left after folding identical ones (knob fold_records); 0 = not folded */
    double   ms_fold;            /* device time of the folding stage (between scatter and count)       */
    uint64_t n_mid_bins;         /* shared-memory count path: tables (mid bins) the bins were cut into; 0 = global-table path */
    uint64_t n_slow_bins;        /* ... of which overflowed shared memory and were redone in a global-memory table */
    double   ms_partition;       /* partitioned count path (count_mode 2): device time of the k-mer partition (k_sub_hist/scan/scatter); ms_stage[3] is then k_count_keys alone */
} fkm_stats;

const char* fkm_last_error(void);

/* device < 0: current device.  stream: a cudaStream_t (as void*) to launch on,
 * or NULL for a stream the context creates.                                    */
int  fkm_ctx_create(int device, void* stream, fkm_ctx** out);
void fkm_ctx_destroy(fkm_ctx* ctx);
int  fkm_ctx_sync(fkm_ctx* ctx);
/* tuning knobs (optional).
 * "count_mode" (use_ht = 1): 2 = the canonical k-mers of a bin are expanded once, hash-partitioned into sub-buckets and counted in
 *   tables in shared memory (csrc/fkm_part.cuh; the default: DRAM sees the records once, the keys twice and the result once, and
 *   the stage does not depend on the input size); 1 = dual-minimizer mid bins, tables in shared memory (csrc/fkm_smem.cuh);
 *   0 = tables in global memory with record folding (round 1; also the fallback of mode 2).
 * "sort_partition" (use_ht = 0): 1 = k-mers expanded once, sub-buckets by their top bits, chunk sort in shared memory (default);
 *   2 = ordered shared-memory tables (k_count_keys_ordered, opt-in); 0 = the round-1 passes.
 * "bin_split": internal bins per bin on the hash path (0 = chosen from the input size; a power of two otherwise, the same on every
 *   rank of a multi-GPU job), see fkm_job_bins.  "speculative_scatter" (default 1): the FASTA front end scatters every scanned
 *   chunk under the PCIe copy into bin regions sized from a forecast.  "part_fill" (0.45), "part_budget_keys" (2^30),
 *   "part_max_subs" (768), "part_two_ctas" (0): planning of the partitioned stage.
 * "smem_table_slots" / "smem_slow_slots" / "debug_rho_scale" / "debug_event_scale" / "debug_force_lsd": test hooks (smaller tables
 *   force the slow path / the global-table fallback, wrong estimates force the overflow paths).
 * Also {"table_budget_bytes","async_table_bytes","sort_budget_keys","load_factor","ingest_chunk_bytes","smem_fill",
 * "fold_records" (hash path in global tables with k <= 32: fold identical super-k-mer records into one weighted record before
 * counting, unless the sampled first bins show more than "fold_max_ratio" (0.6) distinct records; 0 = never), "fold_table_bytes",
 * "fold_pool", "cas_first" (experiment: probe with the CAS itself; slower, see profiles/README.md)} */
int  fkm_ctx_set(fkm_ctx* ctx, const char* name, double value);

/* b = min(4^m, max_b) and outputDir = outputDirectory + prefix + "k"+k+"_m"+m+"_x"+x+"_b"+b+"_s"+sequenceType
 * (TCFG:32-33).  out_dir may be NULL.                                          */
int  fkm_derive(const fkm_config* cfg, int32_t* b, char* out_dir, size_t out_dir_cap);

/* ---- the drop-in call: replaces the body of executeJob (SBKC:989-1046) ------
 * Reads cfg->dataset (FASTA), counts on the GPU, writes the bin files when
 * cfg->write (layout SBKC:550-606 sort path incl. "EOF" trailer, SBKC:715-734 HT
 * path).  stats may be NULL.                                                   */
int  fkm_execute_job(fkm_ctx* ctx, const fkm_config* cfg, fkm_stats* stats);

/* The same job on several GPUs of one node (no Python, no NCCL): the FASTA text is cut at record boundaries into one
 * byte range per GPU, every GPU scans its range, bins are assigned to GPUs by LPT over the exact histogram (the role of
 * MultiprocessorSchedulingPartitioner.scala:35-69), the records move GPU to GPU (cudaMemcpyPeerAsync over NVLink: Spark's
 * shuffle, SBKC:1035,1042), every GPU counts and writes the bins it owns.  devices: n_devices CUDA device ids.  The bin
 * files are the single-GPU job's (byte for byte for use_ht = 0; the same lines in another order for use_ht = 1, whose
 * order is unspecified in the reference too, SBKC:723).  A single long record is not cut: it is counted by one GPU.    */
int  fkm_execute_job_multi(const int32_t* devices, int32_t n_devices, const fkm_config* cfg, fkm_stats* stats);

/* test hook, needs no GPU: the bin owners and the exchange plan fkm_execute_job_multi makes for rank `rank` of n GPUs from the ranks'
 * histograms h_rec / h_kmer [n][bins * split] (the layout of fastkmer_b200.multigpu.plan_exchange, which it must equal).           */
int  fkm_debug_multi_plan(int32_t n, int32_t bins, int32_t split, const uint64_t* h_rec, const uint64_t* h_kmer, int32_t rank,
                          int32_t* owner, uint64_t* send_base, uint64_t* send_off, uint64_t* recv_off, uint64_t* bin_rec, uint64_t* bin_kmer,
                          uint64_t* seg_src, uint64_t* seg_dst, uint64_t seg_cap, uint64_t* n_seg);

/* ---- in-memory variants (same path, no file I/O) ---------------------------- */
/* FASTA text in host memory (pinned if it came from fkm_host_alloc).  The raw
 * text is copied to the GPU and parsed there (record split of SURVEY App. A.1). */
int  fkm_count_fasta(fkm_ctx* ctx, const fkm_config* cfg, const uint8_t* fasta, uint64_t n_bytes,
                     fkm_result** out, fkm_stats* stats);

/* 2-bit packed input, HOST memory.  bases: 32 positions per uint64, first
 * position in the two MSBs (A=0 C=1 G=2 T=3, UTIL:19-22); invalid: 32 positions
 * per uint32, first position in the MSB, set for every non-ACGT byte and for the
 * one separator position that follows each record.                             */
int  fkm_count_packed_host(fkm_ctx* ctx, const fkm_config* cfg, const uint64_t* bases, const uint32_t* invalid,
                           uint64_t n_positions, fkm_result** out, fkm_stats* stats);
/* Same layout, DEVICE memory already resident (what `value` in bench.py times). */
int  fkm_count_packed_device(fkm_ctx* ctx, const fkm_config* cfg, const void* d_bases, const void* d_invalid,
                             uint64_t n_positions, fkm_result** out, fkm_stats* stats);

/* Host-side FASTA -> packed layout (record split of SURVEY App. A.1: '\n' is
 * stripped, every other non-ACGT byte is an invalid position).  Returns the
 * number of positions; call with bases==NULL to size (words = (n+31)/32).      */
int  fkm_pack_fasta(const uint8_t* fasta, uint64_t n_bytes, uint64_t* bases, uint32_t* invalid,
                    uint64_t cap_positions, uint64_t* n_positions, uint64_t* n_bases);
/* the same with `threads` host threads (<= 0: one per hardware thread; 1: the plain loop); bit-identical output.  fkm_pack_fasta
 * is fkm_pack_fasta_mt(…, 0).  Measured: profiles/README.md (the packer does not beat the PCIe bus on 16 threads, so the
 * end-to-end path ships the text and parses it on the device).                                                             */
int  fkm_pack_fasta_mt(const uint8_t* fasta, uint64_t n_bytes, uint64_t* bases, uint32_t* invalid,
                       uint64_t cap_positions, uint64_t* n_positions, uint64_t* n_bases, int32_t threads);

/* pinned host memory for inputs (so H2D runs at PCIe speed) */
int  fkm_host_alloc(size_t bytes, void** out);
void fkm_host_free(void* p);

/* ---- results -----------------------------------------------------------------
 * The device arrays of a result live in the context's job arena: they stay valid
 * until the NEXT job on the same context (or fkm_ctx_destroy).  Copy or write a
 * result before counting again; a stale result fails with FKM_EINVAL.           */
uint64_t fkm_result_size(const fkm_result* r);                   /* distinct entries            */
int32_t  fkm_result_num_bins(const fkm_result* r);               /* b                            */
int32_t  fkm_result_sorted(const fkm_result* r);                 /* 1 if k-mers ascend inside each bin (use_ht=0) */
/* per-bin start offsets into the entry arrays, b+1 values */
int  fkm_result_bin_offsets(const fkm_result* r, uint64_t* offsets);
/* entries in bin-major order; key = 2k-bit integer, first base most significant,
 * hi = bits 127..64 (0 when k <= 32).  Any of the pointers may be NULL.        */
int  fkm_result_copy(const fkm_result* r, int32_t* bin, uint64_t* key_hi, uint64_t* key_lo, uint32_t* count);
/* writes <out_dir>/bin<id> for every non-empty bin (SBKC:550-606 / 715-734)    */
int  fkm_result_write(const fkm_result* r, const char* out_dir);
void fkm_result_free(fkm_result* r);
/* A copy in plain device memory that stays valid across later jobs (free it with fkm_result_free).           */
int  fkm_result_clone(const fkm_result* r, fkm_result** out);
/* sum over the (bin, k-mer) pairs present in both of count_a * count_b; a and b are clones of sorted
 * (use_ht = 0) results of one configuration.  Building block of the multi-sample distances.                  */
int  fkm_result_dot(fkm_ctx* ctx, const fkm_result* a, const fkm_result* b, uint64_t* dot);

/* ---- synthetic inputs of SURVEY §8(d) (counter-based splitmix64) -------------- */
typedef struct fkm_synth {
    uint64_t seed_genome, seed_reads, seed_errors;
    uint64_t genome_len, n_reads, read_len;
    uint64_t first_read;          /* global index of this shard's first read (multi-GPU sharding) */
} fkm_synth;
/* FASTA text ('>r<i>\n<seq>\n') into host memory; out==NULL sizes. */
int  fkm_synth_fasta_host(const fkm_synth* s, uint8_t* out, uint64_t cap, uint64_t* n_bytes);
/* packed layout generated directly in device memory (owned by ctx until
 * fkm_device_free).  d_bases / d_invalid receive device pointers.              */
int  fkm_synth_packed_device(fkm_ctx* ctx, const fkm_synth* s, void** d_bases, void** d_invalid, uint64_t* n_positions);
/* One long sequence (BASELINE config 3): bases [first_pos, first_pos + n_bases) of a synthetic genome with
 * 5 % 5-kb repeats and 0.5 % 1-kb N runs, as a single record ('>chr\n' + 70-column lines on the host).      */
typedef struct fkm_synth_long { uint64_t seed_genome, seed_repeats, seed_n, first_pos, n_bases; } fkm_synth_long;
int  fkm_synth_long_fasta_host(const fkm_synth_long* s, uint8_t* out, uint64_t cap, uint64_t* n_bytes);
int  fkm_synth_long_packed_device(fkm_ctx* ctx, const fkm_synth_long* s, void** d_bases, void** d_invalid, uint64_t* n_positions);
int  fkm_device_free(fkm_ctx* ctx, void* d_ptr);

/* ---- staged entry points: one process per GPU, bins owned per GPU ---------------
 * The stages of fkm_count_packed_device, cut where Spark's shuffle sits
 * (reduceByKey, SBKC:1035,1042), so that the host can exchange bins between GPUs:
 *   fkm_mg_scan     scan this rank's shard; hist_rec / hist_kmer [b] receive the records and
 *                   k-mers this shard puts into every bin (the role of getBinsEstimateSizes,
 *                   SBKC:172-288, but exact).  Starts a job on the context.
 *   fkm_mg_scatter  write the shard's super-k-mer records to d_send (caller-allocated,
 *                   sum(hist_rec) * fkm_record_bytes bytes) at record offset bin_base[bin]
 *                   (b+1 values; owner-major so that each peer's bins are contiguous).
 *   (the host runs one variable-size all-to-all on the send buffers)
 *   fkm_mg_regroup  received records are source-major; segment i = [seg_src[i], seg_src[i+1])
 *                   moves to record offset seg_dst[i] of a bin-major buffer (returned in d_out,
 *                   owned by the context until its next job).
 *   fkm_mg_count    count the bins this rank owns: bin_rec / bin_kmer [b] = records and k-mers
 *                   of bin b in d_records (0 for bins owned elsewhere).
 * "b" above is the job's number of INTERNAL bins, fkm_job_bins(): for deep inputs the hash path cuts every bin of the
 * configuration into 2^j internal bins by a second hash of the signature (all internal bins of a bin are consecutive
 * and must have one owner); knob "bin_split" fixes 2^j, and every rank of a job must use the same value.            */
int32_t fkm_record_bytes(const fkm_config* cfg);
int  fkm_job_bins(fkm_ctx* ctx, const fkm_config* cfg, uint64_t n_positions, int32_t* bins);
int  fkm_mg_scan(fkm_ctx* ctx, const fkm_config* cfg, const void* d_bases, const void* d_invalid, uint64_t n_positions,
                 uint64_t* hist_rec, uint64_t* hist_kmer);
/* fkm_mg_scan on FASTA text in host memory (copied to the GPU and parsed there) */
int  fkm_mg_scan_fasta(fkm_ctx* ctx, const fkm_config* cfg, const uint8_t* fasta, uint64_t n_bytes,
                       uint64_t* hist_rec, uint64_t* hist_kmer, uint64_t* n_bases);
int  fkm_mg_scatter(fkm_ctx* ctx, const uint64_t* bin_base, void* d_send);
int  fkm_mg_regroup(fkm_ctx* ctx, const fkm_config* cfg, const void* d_recv, uint64_t n_records,
                    const uint64_t* seg_src, const uint64_t* seg_dst, uint64_t n_seg, void** d_out);
int  fkm_mg_count(fkm_ctx* ctx, const fkm_config* cfg, const void* d_records, const uint64_t* bin_rec, const uint64_t* bin_kmer,
                  fkm_result** out, fkm_stats* stats);

/* ---- multi-sample distances (skc.multisequence) -----------------
 * Reads carry a sample tag = the leading run of [A-Za-z0-9_] of their header.  For every pair of samples
 * dist[a*max_samples + b] = sum over distinct canonical k-mers of (c_a - c_b)^2 (squared euclidean distance of the
 * per-sample count vectors: SparkMultiSequenceKmerCounter.scala:458-520, multiseq/SquaredEuclidean.java:19-32).
 * names receives max_samples x 64 bytes (NUL-padded tags in order of first appearance) or may be NULL.  *merged, when
 * requested, holds (k-mer, sum of counts) per bin, ascending, written without the "EOF" trailer (MSKC:487,524).  */
int  fkm_multiseq_fasta(fkm_ctx* ctx, const fkm_config* cfg, const uint8_t* fasta, uint64_t n_bytes, int32_t max_samples,
                        int32_t* n_samples, char* names, double* dist, fkm_result** merged, fkm_stats* stats);

/* ---- test hooks (stage-level parity against the oracle) ----------------------- */
/* bin of every window start (−1 where the k-window holds an invalid position);
 * bins_out has n_positions entries.                                             */
int  fkm_debug_window_bins(fkm_ctx* ctx, const fkm_config* cfg, const uint64_t* bases, const uint32_t* invalid,
                           uint64_t n_positions, int32_t* bins_out);
/* the device ingest's packed arrays (must equal fkm_pack_fasta bit for bit)     */
int  fkm_debug_pack_fasta_device(fkm_ctx* ctx, const uint8_t* fasta, uint64_t n_bytes, uint64_t* bases, uint32_t* invalid,
                                 uint64_t cap_positions, uint64_t* n_positions, uint64_t* n_bases);
uint64_t fkm_total_launches(void);    /* kernels launched by this process so far */

#ifdef __cplusplus
}
#endif
#endif /* FASTKMER_B200_H */
