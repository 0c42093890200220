// fkm_lib.cu — host pipeline and C ABI (include/fastkmer_b200.h) of the B200 path.
//
// Replaces everything behind SparkBinKmerCounter.executeJob (SBKC:989-1046):
//   mapPartitions(getSuperKmers)        -> k_scan<MODE 0> (exact bin histogram) + k_scan<MODE 1> (scatter)
//   reduceByKey(_ ++ _)  (the shuffle)  -> bin-major record buffer sized by the exact histogram
//   foreachPartition(extractKXmersHT)   -> k_count_ht + k_compact_ht over batches of bins
//   foreachPartition(extractKXmers)     -> k_expand + segmented radix sort + k_rle
// There is no CPU fallback: every entry point needs a CUDA device.
#include "fkm_kernels.cuh"
#include "fkm_smem.cuh"
#include "fkm_part.cuh"
#include "fkm_ingest.cuh"
#include "fkm_host.h"
#include "../../include/fastkmer_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace fkm;

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

int fkm_set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fkm_set_error(e_ == cudaErrorMemoryAllocation ? FKM_ENOMEM : FKM_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CKL() do { g_launches++; ctx->job_launches++; CK(cudaGetLastError()); } while (0)

// Bump arena for everything a job allocates on the device.  Slabs are only ever added
// inside a job; when the next job begins, several slabs are merged into one of the
// peak size, so a steady stream of similar jobs does no cudaMalloc / cudaFree at all.
struct Arena {
    struct Slab { char* p; size_t cap, used; };
    std::vector<Slab> slabs;
    struct Mark { size_t slab, used; };
    static size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }
    cudaError_t alloc(void** out, size_t n) {
        n = up(n ? n : 16, 512);
        if (!slabs.empty() && slabs.back().used + n <= slabs.back().cap) {
            *out = slabs.back().p + slabs.back().used; slabs.back().used += n; return cudaSuccess;
        }
        Slab sl; sl.cap = std::max<size_t>(n, (size_t)64 << 20); sl.used = n;
        cudaError_t e = cudaMalloc((void**)&sl.p, sl.cap);
        if (e != cudaSuccess) return e;
        slabs.push_back(sl); *out = sl.p;
        return cudaSuccess;
    }
    Mark mark() const { return slabs.empty() ? Mark{0, 0} : Mark{slabs.size() - 1, slabs.back().used}; }
    void release(Mark m) {                       // stack discipline: drop everything allocated after mark()
        if (slabs.empty()) return;
        for (size_t i = m.slab + 1; i < slabs.size(); i++) slabs[i].used = 0;
        // a slab that did not exist at mark() time is simply emptied
        if (m.slab < slabs.size()) slabs[m.slab].used = std::min(slabs[m.slab].used, m.used);
    }
    size_t peak = 0;
    void note_peak() { size_t t = 0; for (auto& sl : slabs) t += sl.cap == 0 ? 0 : std::max(sl.used, (size_t)0); peak = std::max(peak, t); }
    cudaError_t reset() {                        // caller guarantees the device is idle
        size_t total = 0; for (auto& sl : slabs) total += sl.cap;
        if (slabs.size() > 1) {
            for (auto& sl : slabs) cudaFree(sl.p);
            slabs.clear();
            Slab sl; sl.cap = up(total + total / 50, (size_t)64 << 20); sl.used = 0;
            cudaError_t e = cudaMalloc((void**)&sl.p, sl.cap);
            if (e != cudaSuccess) return e;
            slabs.push_back(sl);
        }
        for (auto& sl : slabs) sl.used = 0;
        return cudaSuccess;
    }
    void destroy() { for (auto& sl : slabs) cudaFree(sl.p); slabs.clear(); }
};

struct fkm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int n_sm = 0;
    size_t smem_optin = 0;
    double table_budget_bytes = 4.0 * (1ull << 30);
    double sort_budget_keys = 256.0 * (1 << 20);
    double load_factor = 0.6;
    double ingest_chunk_bytes = 256.0 * (1 << 20);   // FASTA text is streamed to the GPU in chunks of about this size
    double debug_event_scale = 1.0;   // test hook: scales the run-event list capacity (forces the second-scan fallback)
    double fold_records = 1.0;        // hash path, k <= 32: 1 = fold identical super-k-mer records into one weighted record before counting
    double fold_table_bytes = 1024.0 * 1048576.0;   // ... in batches of bins whose record tables (32-byte slots) fit this
    double fold_pool = 128.0;         // ... records per warp in k_fold_insert (64, 128 or 256)
    double fold_max_ratio = 0.6;      // ... unless the first batch shows that more than this share of the records is distinct
    double cas_first = 0.0;           // hash path: 1 = probe with the CAS itself instead of a read followed by a CAS
    double debug_force_lsd = 0.0;     // sort path: 0 = auto (MSD + shared-memory chunk sort for 64-bit keys, LSD passes for 128-bit), 1 = LSD, 2 = MSD
    double sort_partition = 1.0;      // sort path: 1 = expand once, sub-buckets by the top key bits, shared-memory chunk sort (k_radix_local) + run-length count; 2 = the same sub-buckets + ordered shared-memory tables (k_count_keys_ordered: no sort at all; opt-in: the top bits of the k-mers of a minimizer bin are too unevenly spread for it); 0 = the round-1 passes
    double debug_rho_scale = 1.0;     // test hook: scales the learnt distinct/k-mer ratio (forces the overflow fallback)
    double count_mode = 2.0;          // hash path: 2 = k-mers hash-partitioned into sub-buckets, tables in shared memory (fkm_part.cuh); 1 = dual-minimizer mid bins, tables in shared memory (fkm_smem.cuh); 0 = tables in global memory
    double smem_table_slots = 0.0;    // test hook: slots of the shared-memory table (0 = as many as fit)
    double smem_slow_slots = 1048576.0;   // slots of every CTA's private global table (slow path of k_count_smem)
    double smem_fill = 0.6;           // share of the table's capacity the planner aims at
    double part_fill = 0.45;          // partitioned count path (count_mode 2): distinct k-mers per sub-bucket as a share of the table's slots
    double bin_split = 0.0;           // internal bins per bin (hash path, count_mode 0 or 2): 0 = chosen from the input size, else a power of two (a multi-GPU job sets the same value on every rank)
    int job_split = 0;                // log2 of the internal bins per bin of the job in progress (fkm_common.h split_bin)
    double speculative_scatter = 1.0; // FASTA front end, hash path: scatter the scanned chunks under the PCIe copy into bin regions sized from a forecast
    double part_two_ctas = 0.0;       // ... 1: k_count_keys as two 512-thread CTAs per SM with half-size tables (experiment)
    double part_max_subs = 768.0;     // ... sub-buckets a bin may need before the job is left to the global-table pipeline
    double part_budget_keys = 1024.0 * 1048576.0;   // ... k-mers per batch of bins (the key buffer holds one batch)
    std::vector<cudaEvent_t> evpool;  // ... per-batch timing events
    double async_table_bytes = 1024.0 * (1 << 20); // tables of one asynchronous batch; 0 disables the asynchronous phase (0.25-16 GB all within 8 %, profiles/r1_table_sweep.txt)
    uint64_t job_launches = 0;
    uint64_t gen = 0;                 // job generation: results of older jobs are invalid
    Arena arena;
    struct ScanState* mg_scan = nullptr;   // state between fkm_mg_scan and fkm_mg_scatter
    cudaEvent_t ev[10];
    cudaEvent_t evs[24];              // sampled per-kernel timings inside the asynchronous phase
    cudaStream_t copy_stream = nullptr;   // H2D of FASTA chunks, overlapped with parsing / scanning on `stream`
    cudaEvent_t copied[2];
};

// job-lifetime device memory (see Arena); "free" is a no-op, the arena is reset by the next job
static inline cudaError_t dmalloc(fkm_ctx* ctx, void** p, size_t n) { return ctx->arena.alloc(p, n); }
template <typename T> static inline cudaError_t dmalloc(fkm_ctx* ctx, T** p, size_t n) { return dmalloc(ctx, (void**)p, n); }
static inline void dfree(fkm_ctx*, void*) {}

struct Chunk { void* keys = nullptr; uint32_t* cnt = nullptr; uint64_t n = 0; };
struct fkm_result {
    bool owned = false;                           // clone: arrays in plain device memory, valid until fkm_result_free
    unsigned long long* d_base = nullptr;         // clone: device copy of out_base (for k_sparse_dot)
    bool eof_trailer = true;                      // sorted files end with "EOF" (SBKC:606); the multisequence writer has none (MSKC:524-526)
    fkm_ctx* ctx = nullptr; uint64_t gen = 0;     // arrays live in ctx's arena until its next job
    int device = 0;
    int32_t B = 0, k = 0; bool wide = false; bool sorted = false;
    std::vector<Chunk> chunks;
    std::vector<uint64_t> out_base;      // B+1
    uint64_t total = 0;
};

struct ScanState;
static void free_scan_state(ScanState* p);      // defined after ScanState

extern "C" const char* fkm_last_error(void) { return g_err.c_str(); }
extern "C" uint64_t fkm_total_launches(void) { return g_launches.load(); }

extern "C" int fkm_ctx_create(int device, void* stream, fkm_ctx** out) {
    if (!out) return fkm_set_error(FKM_EINVAL, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fkm_set_error(FKM_ECUDA, "no CUDA device (%s); fastkmer_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= n) return fkm_set_error(FKM_EINVAL, "device %d out of range (%d devices)", device, n);
    CK(cudaSetDevice(device));
    fkm_ctx* c = new fkm_ctx();
    c->device = device;
    if (stream) c->stream = (cudaStream_t)stream;
    else { CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, device));
    c->n_sm = p.multiProcessorCount; c->smem_optin = p.sharedMemPerBlockOptin;
    for (auto& ev : c->ev) CK(cudaEventCreate(&ev));
    for (auto& ev : c->evs) CK(cudaEventCreate(&ev));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto& ev : c->copied) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    *out = c;
    return FKM_OK;
}
extern "C" void fkm_ctx_destroy(fkm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& ev : c->ev) cudaEventDestroy(ev);
    for (auto& ev : c->evs) cudaEventDestroy(ev);
    for (auto& ev : c->evpool) cudaEventDestroy(ev);
    for (auto& ev : c->copied) cudaEventDestroy(ev);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaStreamSynchronize(c->stream);
    c->arena.destroy();
    free_scan_state(c->mg_scan);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}
extern "C" int fkm_ctx_sync(fkm_ctx* c) { CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->stream)); return FKM_OK; }
extern "C" int fkm_ctx_set(fkm_ctx* c, const char* name, double v) {
    if (!c || !name) return fkm_set_error(FKM_EINVAL, "null argument");
    if (!strcmp(name, "table_budget_bytes")) c->table_budget_bytes = v;
    else if (!strcmp(name, "sort_budget_keys")) c->sort_budget_keys = v;
    else if (!strcmp(name, "load_factor")) c->load_factor = v;
    else if (!strcmp(name, "async_table_bytes") || !strcmp(name, "l2_table_bytes")) c->async_table_bytes = v;
    else if (!strcmp(name, "debug_rho_scale")) c->debug_rho_scale = v;
    else if (!strcmp(name, "count_mode")) c->count_mode = v;
    else if (!strcmp(name, "smem_table_slots")) c->smem_table_slots = v;
    else if (!strcmp(name, "smem_slow_slots")) c->smem_slow_slots = v;
    else if (!strcmp(name, "smem_fill")) c->smem_fill = v;
    else if (!strcmp(name, "part_fill")) c->part_fill = v;
    else if (!strcmp(name, "speculative_scatter")) c->speculative_scatter = v;
    else if (!strcmp(name, "part_two_ctas")) c->part_two_ctas = v;
    else if (!strcmp(name, "bin_split")) c->bin_split = v;
    else if (!strcmp(name, "part_max_subs")) c->part_max_subs = v;
    else if (!strcmp(name, "sort_partition")) c->sort_partition = v;
    else if (!strcmp(name, "part_budget_keys")) c->part_budget_keys = v;
    else if (!strcmp(name, "fold_records")) c->fold_records = v;
    else if (!strcmp(name, "fold_max_ratio")) c->fold_max_ratio = v;
    else if (!strcmp(name, "fold_pool")) c->fold_pool = v;
    else if (!strcmp(name, "fold_table_bytes")) c->fold_table_bytes = v;
    else if (!strcmp(name, "cas_first")) c->cas_first = v;
    else if (!strcmp(name, "debug_force_lsd")) c->debug_force_lsd = v;
    else if (!strcmp(name, "debug_event_scale")) c->debug_event_scale = v;
    else if (!strcmp(name, "ingest_chunk_bytes")) c->ingest_chunk_bytes = v;
    else return fkm_set_error(FKM_EINVAL, "unknown knob %s", name);
    return FKM_OK;
}

// start of a job on this context: device idle, arena rewound, older results invalidated
static int job_begin(fkm_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(ctx->arena.reset());
    ctx->gen++;
    ctx->job_launches = 0;
    ctx->job_split = 0;
    return FKM_OK;
}

static int validate(const fkm_config* cfg, int32_t* B) {
    if (!cfg) return fkm_set_error(FKM_EINVAL, "cfg is NULL");
    if (cfg->m < 3 || cfg->m > 15) return fkm_set_error(FKM_EINVAL, "m=%d unsupported (3 <= m <= 15; UTIL:78 overflows Int beyond 15)", cfg->m);
    if (cfg->k < cfg->m || cfg->k > 64) return fkm_set_error(FKM_EINVAL, "k=%d unsupported (m <= k <= 64)", cfg->k);
    if (cfg->max_b < 1) return fkm_set_error(FKM_EINVAL, "B=%d must be >= 1", cfg->max_b);
    if (!cfg->use_ht && cfg->x < 1) return fkm_set_error(FKM_EINVAL, "x=%d: the reference's sort path fails for x < 1 (SBKC:495-508)", cfg->x);
    int64_t b = std::min<int64_t>((int64_t)1 << (2 * cfg->m), (int64_t)cfg->max_b);     // TCFG:32
    *B = (int32_t)b;
    return FKM_OK;
}

// log2 of the internal bins every bin is cut into for this job (fkm_common.h split_bin).  Only the hash path with its own
// partition levels uses them (the sort path's output is ordered inside a bin; the dual-minimizer mode has its cells).
static int choose_split(const fkm_ctx* ctx, const fkm_config* cfg, int32_t B, uint64_t n_pos_hint) {
    if (!cfg->use_ht || (ctx->count_mode >= 1.0 && ctx->count_mode < 2.0)) return 0;
    int s = 0;
    if (ctx->bin_split >= 1.0) { while ((1 << (s + 1)) <= (int)ctx->bin_split && s < 6) s++; }
    else { const double per_bin = (double)n_pos_hint / (double)B; while (per_bin / (double)(1 << s) > 6.0e6 && s < 6) s++; }
    while (s > 0 && ((int64_t)B << s) > 65536) s--;
    return s;
}
extern "C" int fkm_job_bins(fkm_ctx* ctx, const fkm_config* cfg, uint64_t n_positions, int32_t* bins) {
    if (!ctx || !bins) return fkm_set_error(FKM_EINVAL, "null argument");
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    *bins = B << choose_split(ctx, cfg, B, n_positions);
    return FKM_OK;
}

extern "C" int fkm_derive(const fkm_config* cfg, int32_t* b, char* out_dir, size_t cap) {
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    if (b) *b = B;
    if (out_dir) {
        std::string s = std::string(cfg->output_directory ? cfg->output_directory : "") + (cfg->prefix ? cfg->prefix : "") +
                        "k" + std::to_string(cfg->k) + "_m" + std::to_string(cfg->m) + "_x" + std::to_string(cfg->x) +
                        "_b" + std::to_string(B) + "_s" + std::to_string(cfg->sequence_type);       // TCFG:33
        if (s.size() + 1 > cap) return fkm_set_error(FKM_EINVAL, "out_dir buffer too small");
        memcpy(out_dir, s.c_str(), s.size() + 1);
    }
    return FKM_OK;
}

extern "C" int fkm_host_alloc(size_t bytes, void** out) { CK(cudaMallocHost(out, bytes)); return FKM_OK; }
extern "C" void fkm_host_free(void* p) { if (p) cudaFreeHost(p); }

// whole file into pinned host memory (so the H2D copy runs at PCIe speed)
static int fkm_read_file_pinned(const char* path, uint8_t** out, uint64_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) return fkm_set_error(FKM_EIO, "cannot open %s", path);
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); return fkm_set_error(FKM_EIO, "cannot size %s", path); }
    cudaError_t e = cudaMallocHost((void**)out, std::max<long>(sz, 16));
    if (e != cudaSuccess) { fclose(f); return fkm_set_error(FKM_ENOMEM, "pinned alloc of %ld bytes: %s", sz, cudaGetErrorString(e)); }
    size_t got = sz ? fread(*out, 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) { cudaFreeHost(*out); return fkm_set_error(FKM_EIO, "short read on %s", path); }
    *n = (uint64_t)sz;
    return FKM_OK;
}

// ------------------------------------------------------------------ scan setup
typedef void (*ScanKernel)(const ScanParams);
template <int MODE, bool DUAL> static ScanKernel scan_kernel(int NL) {
    switch (NL) {
        case 1: return k_scan<MODE, 1, DUAL>; case 2: return k_scan<MODE, 2, DUAL>; case 3: return k_scan<MODE, 3, DUAL>;
        case 4: return k_scan<MODE, 4, DUAL>; case 5: return k_scan<MODE, 5, DUAL>; default: return k_scan<MODE, 6, DUAL>;
    }
}
// doubling shifts of a sliding minimum over w values: 1, 2, 4, ... then the remainder; w = 1 + sum.  Returns the levels used.
static int fill_shifts(int w, int sh[6]) {
    int n = 0, have = 1;
    for (int j = 0; j < 6; j++) sh[j] = 0;
    while (have * 2 <= w) { sh[n++] = have; have *= 2; }
    if (w > have) sh[n++] = w - have;
    return n;
}
// length of the second (hash-ordered) minimizer of the shared-memory count path: long enough to spread a bin over many
// cells, short enough to leave runs of several windows (only the partition depends on it, never a count)
static int second_minimizer_len(int k) { return k >= 23 ? 15 : std::max(std::min(k, 10), k - 8); }

struct ScanSetup { ScanParams P; size_t smem; int grid; ScanKernel fn; };
template <int MODE, bool DUAL>
static int scan_setup(fkm_ctx* ctx, const fkm_config* cfg, int32_t B, const void* d_bases, const void* d_inv, uint64_t n_pos, ScanSetup* S) {
    ScanParams& P = S->P;
    memset(&P, 0, sizeof P);
    P.bases = (const uint64_t*)d_bases; P.inv = (const uint32_t*)d_inv;
    P.n_pos = n_pos; P.n_words = (n_pos + 31) / 32;
    P.k = cfg->k; P.m = cfg->m; P.w = cfg->k - cfg->m + 1;
    int nl = std::max(1, fill_shifts(P.w, P.sh1));
    if (DUAL) {
        P.m2 = second_minimizer_len(cfg->k);
        P.mask2 = (P.m2 >= 16) ? 0xFFFFFFFFu : ((1u << (2 * P.m2)) - 1u);
        nl = std::max(nl, fill_shifts(cfg->k - P.m2 + 1, P.sh2));
    }
    P.seg_len = 4096;
    P.e_total = (MODE == 2) ? n_pos + (uint64_t)cfg->k - 1 : n_pos;
    P.n_seg = (P.e_total + P.seg_len - 1) / P.seg_len;
    P.Bi = (uint32_t)B; P.split = ctx->job_split; P.B = (uint32_t)B >> ctx->job_split;      // B counts internal bins here
    P.wide = cfg->k > 32;
    P.cap = (cfg->k > 32) ? (125 - cfg->k) : (61 - cfg->k);
    P.smem_hist = (MODE == 0 && !DUAL && B <= kSmemHistMaxB) ? 1 : 0;
    S->smem = P.smem_hist ? (size_t)B * 8 : 0;
    S->fn = scan_kernel<MODE, DUAL>(nl);
    CK(cudaFuncSetAttribute(S->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S->smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, S->fn, kScanThreads, S->smem));
    if (occ < 1) occ = 1;
    const uint64_t ctas = (P.n_seg + kScanThreads / 32 - 1) / (kScanThreads / 32);
    S->grid = (int)std::min<uint64_t>(ctas, (uint64_t)ctx->n_sm * occ);
    if (S->grid < 1) S->grid = 1;
    return FKM_OK;
}

// FKM_TRACE=1: host-side timeline of one job on stderr (where the host blocks between launches)
struct Trace {
    bool on; std::chrono::steady_clock::time_point t0, last;
    Trace() : on(getenv("FKM_TRACE") != nullptr) { t0 = last = std::chrono::steady_clock::now(); }
    void mark(const char* what, long long a = -1) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[fkm trace] %9.3f ms (+%8.3f) %s %lld\n", std::chrono::duration<double, std::milli>(now - t0).count(),
                std::chrono::duration<double, std::milli>(now - last).count(), what, a);
        last = now;
    }
};

static inline uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ the pipeline
// What the scan stage leaves behind for the scatter stage (kept in the context between the
// staged multi-GPU entry points; a local inside the single-GPU pipeline).  The input may be
// scanned in several chunks (FASTA text streamed over PCIe, split at record boundaries):
// every chunk has its own packed arrays and run-event list, the bin histogram is shared.
struct ChunkScan {
    const void* d_bases = nullptr; const void* d_inv = nullptr; uint64_t n_pos = 0;
    ulonglong2* d_events = nullptr; uint64_t ev_cap = 0;
    unsigned long long* d_count = nullptr;      // [0] events written ([0..1]: the two lists of a dual scan)
    int* d_ovf = nullptr;
    unsigned long long n_events = 0; int ev_ovf = 0;
    ulonglong2* d_events2[2] = {nullptr, nullptr}; uint64_t ev_cap2[2] = {0, 0}; unsigned long long n_events2[2] = {0, 0};   // dual scan
};
struct ScanState {
    fkm_config cfg; int32_t B = 0;
    unsigned long long *d_hist_rec = nullptr, *d_hist_kmer = nullptr;
    std::vector<ChunkScan> chunks;
    std::vector<unsigned long long> h_rec, h_kmer;
    uint64_t n_pos_total = 0;
    bool valid = false;
    uint64_t gen = 0;                           // job generation of the context the state belongs to
    // speculative scatter (FASTA front end, hash path): while the text is still on the PCIe bus the chunks already scanned are
    // scattered into bin regions sized from a forecast of the histogram; h_rec (exact, known at the end) are the bins' record counts
    struct Spec {
        bool on = false, failed = false;
        void* d_records = nullptr; unsigned long long *d_bin_base = nullptr, *d_cursor = nullptr; int cursor_shift = 0; int* d_ovf = nullptr;
        size_t next_chunk = 0; std::vector<unsigned long long> base;      // [B+1] region offsets
    } spec;
    // dual scan (shared-memory count path): runs cut by the signature and a second minimizer, histogram per (bin, cell)
    bool dual = false; int cell_bits = 0; uint32_t sample_hi = 0;
    unsigned long long *d_cell_rec = nullptr, *d_cell_kmer = nullptr;
};

static void free_scan_state(ScanState* p) { delete p; }

// n_pos_hint: positions the whole job will scan (sizes the cells of a dual scan)
static int scan_begin(fkm_ctx* ctx, const fkm_config* cfg, int32_t B, ScanState* S, bool dual = false, uint64_t n_pos_hint = 0) {
    const size_t bB = (size_t)B * 8;
    S->cfg = *cfg; S->B = B; S->chunks.clear(); S->n_pos_total = 0; S->valid = false; S->gen = ctx->gen; S->spec = ScanState::Spec();
    S->h_rec.assign((size_t)B, 0); S->h_kmer.assign((size_t)B, 0);
    CK(dmalloc(ctx, &S->d_hist_rec, bB)); CK(dmalloc(ctx, &S->d_hist_kmer, bB));
    CK(cudaMemsetAsync(S->d_hist_rec, 0, bB, ctx->stream)); CK(cudaMemsetAsync(S->d_hist_kmer, 0, bB, ctx->stream));
    S->dual = dual; S->cell_bits = 0; S->sample_hi = 0; S->d_cell_rec = S->d_cell_kmer = nullptr;
    if (dual) {
        // about 2048 k-windows per cell: fine enough to pack mid bins of a few thousand k-mers evenly
        const uint64_t per_bin = n_pos_hint / (uint64_t)B;
        while (S->cell_bits < 12 && (2048ull << S->cell_bits) < per_bin) S->cell_bits++;
        while (S->cell_bits > 0 && ((uint64_t)B << S->cell_bits) > (1ull << 26)) S->cell_bits--;      // at most 64 M cells
        S->sample_hi = (uint32_t)std::max(1, B / 64);
        const size_t cells = (size_t)B << S->cell_bits;
        CK(dmalloc(ctx, &S->d_cell_rec, cells * 8)); CK(dmalloc(ctx, &S->d_cell_kmer, cells * 8));
        CK(cudaMemsetAsync(S->d_cell_rec, 0, cells * 8, ctx->stream)); CK(cudaMemsetAsync(S->d_cell_kmer, 0, cells * 8, ctx->stream));
    }
    return FKM_OK;
}
// one scan launch (asynchronous): bin histogram += this chunk, run events of this chunk
static int scan_chunk(fkm_ctx* ctx, ScanState* S, const void* d_bases, const void* d_inv, uint64_t n_pos) {
    cudaStream_t s = ctx->stream;
    const fkm_config* cfg = &S->cfg;
    ChunkScan C;
    C.d_bases = d_bases; C.d_inv = d_inv; C.n_pos = n_pos;
    const int w_mm = cfg->k - cfg->m + 1;
    // runs average ~(w+1)/2 windows on random sequence; leave generous head-room, an overflow falls back to a second scan
    C.ev_cap = (uint64_t)(((double)n_pos * std::min(0.5, 2.4 / (double)(w_mm + 1)) + (double)(1u << 20)) * ctx->debug_event_scale) + 64;
    CK(dmalloc(ctx, &C.d_count, 16)); CK(dmalloc(ctx, &C.d_ovf, 8));
    CK(cudaMemsetAsync(C.d_count, 0, 16, s)); CK(cudaMemsetAsync(C.d_ovf, 0, 8, s));
    ScanSetup Q; int rc;
    if (!S->dual) {
        CK(dmalloc(ctx, &C.d_events, (size_t)C.ev_cap * 16));
        rc = scan_setup<0, false>(ctx, cfg, S->B, d_bases, d_inv, n_pos, &Q); if (rc) return rc;
        Q.P.events = C.d_events; Q.P.ev_cap = C.ev_cap;
    } else {
        // two cut rules: about 2/(w+1) + 2/(w2+1) runs per position; the sample bins' events go to their own (short) list
        const int w2 = cfg->k - second_minimizer_len(cfg->k) + 1;
        const uint64_t total = (uint64_t)(((double)n_pos * std::min(0.7, 2.4 / (double)(w_mm + 1) + 2.4 / (double)(w2 + 1)) + (double)(1u << 20)) * ctx->debug_event_scale) + 64;
        const double frac = (double)S->sample_hi / (double)S->B;
        C.ev_cap2[0] = std::min<uint64_t>(total, (uint64_t)((double)total * frac * 3.0) + (1u << 18));     // the sample bins' events
        C.ev_cap2[1] = total;                                                                                    // every event
        CK(dmalloc(ctx, &C.d_events2[0], (size_t)C.ev_cap2[0] * 16)); CK(dmalloc(ctx, &C.d_events2[1], (size_t)C.ev_cap2[1] * 16));
        rc = scan_setup<0, true>(ctx, cfg, S->B, d_bases, d_inv, n_pos, &Q); if (rc) return rc;
        Q.P.events2[0] = C.d_events2[0]; Q.P.events2[1] = C.d_events2[1]; Q.P.ev_cap2[0] = C.ev_cap2[0]; Q.P.ev_cap2[1] = C.ev_cap2[1];
        Q.P.cell_bits = S->cell_bits; Q.P.sample_hi = S->sample_hi; Q.P.cell_rec = S->d_cell_rec; Q.P.cell_kmer = S->d_cell_kmer;
    }
    Q.P.hist_rec = S->d_hist_rec; Q.P.hist_kmer = S->d_hist_kmer;
    Q.P.ev_count = C.d_count; Q.P.ev_overflow = C.d_ovf;
    if (n_pos) { Q.fn<<<Q.grid, kScanThreads, Q.smem, s>>>(Q.P); CKL(); }
    S->chunks.push_back(C);
    S->n_pos_total += n_pos;
    return FKM_OK;
}
static int scan_end(fkm_ctx* ctx, ScanState* S, fkm_stats* st) {
    cudaStream_t s = ctx->stream;
    const size_t bB = (size_t)S->B * 8;
    if (S->dual) {      // the dual scan keeps its histogram per (bin, cell) only
        k_cells_bin_totals<<<(unsigned)(((uint64_t)S->B * 32 + 255) / 256), 256, 0, s>>>(S->d_cell_rec, S->d_cell_kmer, S->cell_bits, S->B, S->d_hist_rec, S->d_hist_kmer); CKL();
    }
    CK(cudaMemcpyAsync(S->h_rec.data(), S->d_hist_rec, bB, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(S->h_kmer.data(), S->d_hist_kmer, bB, cudaMemcpyDeviceToHost, s));
    for (ChunkScan& C : S->chunks) {
        CK(cudaMemcpyAsync(C.n_events2, C.d_count, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&C.ev_ovf, C.d_ovf, 4, cudaMemcpyDeviceToHost, s));
    }
    int spec_ovf = 0;
    if (S->spec.on) CK(cudaMemcpyAsync(&spec_ovf, S->spec.d_ovf, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (S->spec.on && (spec_ovf || S->spec.next_chunk != S->chunks.size())) S->spec.failed = true;
    for (ChunkScan& C : S->chunks) if (!S->dual) C.n_events = C.n_events2[0];
    st->d2h_bytes += 2 * bB + 12 * S->chunks.size();
    S->valid = true;
    return FKM_OK;
}

// k_compact_ht gives every CTA an equal, contiguous share of the tiles: whole waves only (4 per launch), so
// no SM idles behind a partial last wave
template <bool WIDE>
static uint64_t compact_grid(fkm_ctx* ctx) {
    static int per_sm = 0;
    if (!per_sm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_compact_ht<WIDE>, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    }
    return (uint64_t)ctx->n_sm * (uint64_t)per_sm * 4ull;
}

// stage 1 on input that is already resident: one chunk
static int stage_scan(fkm_ctx* ctx, const fkm_config* cfg, int32_t B, const void* d_bases, const void* d_inv, uint64_t n_pos,
                      ScanState* S, fkm_stats* st, bool dual = false) {
    int rc = scan_begin(ctx, cfg, B, S, dual, n_pos); if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = scan_chunk(ctx, S, d_bases, d_inv, n_pos); if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    rc = scan_end(ctx, S, st); if (rc) return rc;
    float ms; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); st->ms_stage[1] = ms;
    return FKM_OK;
}

// stage 2: run events -> super-k-mer records at d_records[bin_base[bin] + ...] (the "shuffle").  bin_base is any
// per-bin record offset table: bin-major on one GPU, owner-major for the multi-GPU send buffer.
// the write cursors of the scatter stage: one per bin, 1 << shift words apart
static int cursor_shift_for(int32_t B) { int sh = 5; while (sh > 0 && ((uint64_t)B << sh) > (1ull << 20)) sh--; return sh; }
// one chunk's run events -> records (asynchronous)
static int scatter_chunk(fkm_ctx* ctx, ScanState* S, const ChunkScan& C, const unsigned long long* d_bin_base, unsigned long long* d_cursor,
                         int cursor_shift, void* d_records, int* d_overflow) {
    const fkm_config* cfg = &S->cfg;
    const bool wide = cfg->k > 32;
    ScatterParams Q;
    Q.events = C.d_events; Q.n_events = C.n_events; Q.bases = (const uint64_t*)C.d_bases; Q.n_words = (C.n_pos + 31) / 32;
    Q.B = (uint32_t)S->B >> ctx->job_split; Q.split = ctx->job_split; Q.cap = wide ? (125 - cfg->k) : (61 - cfg->k); Q.k = cfg->k;
    Q.bin_base = d_bin_base; Q.cursor = d_cursor; Q.cursor_shift = cursor_shift; Q.records = d_records; Q.overflow = d_overflow;
    if (C.n_events) {
        const unsigned grid = (unsigned)((C.n_events + 255) / 256);
        if (wide) k_scatter_events<true><<<grid, 256, 0, ctx->stream>>>(Q); else k_scatter_events<false><<<grid, 256, 0, ctx->stream>>>(Q);
        CKL();
    }
    return FKM_OK;
}

static int stage_scatter(fkm_ctx* ctx, ScanState* S, const unsigned long long* d_bin_base,
                         void* d_records, uint64_t n_rec, fkm_stats* st) {
    cudaStream_t s = ctx->stream;
    const fkm_config* cfg = &S->cfg;
    // One write cursor per bin, 256 bytes apart.  The stage does one atomic per run event on these B words; packed
    // densely (16 KB at B = 2048) they are 8 of the 2-KB grains by which addresses are dealt to the two dies and the
    // L2 slices, and the stage took 17, 19 or 23 ms for the same kernel depending on where the allocation had landed.
    // Spread out, the hot words cover all slices whatever the placement.
    const int cursor_shift = cursor_shift_for(S->B);
    unsigned long long* d_cursor = nullptr;
    CK(dmalloc(ctx, &d_cursor, ((size_t)S->B << cursor_shift) * 8));
    CK(cudaMemsetAsync(d_cursor, 0, ((size_t)S->B << cursor_shift) * 8, s));
    CK(cudaEventRecord(ctx->ev[1], s));
    for (ChunkScan& C : S->chunks) {
        if (!C.ev_ovf) {
            int rc = scatter_chunk(ctx, S, C, d_bin_base, d_cursor, cursor_shift, d_records, nullptr); if (rc) return rc;
        } else {
            // the event list was too small for this chunk: scan it again, writing the records directly
            st->n_fallbacks++;
            ScanSetup Q; int rc = scan_setup<1, false>(ctx, cfg, S->B, C.d_bases, C.d_inv, C.n_pos, &Q); if (rc) return rc;
            Q.P.bin_base = d_bin_base; Q.P.cursor = d_cursor; Q.P.cursor_shift = cursor_shift; Q.P.records = d_records;
            if (n_rec && C.n_pos) { Q.fn<<<Q.grid, kScanThreads, Q.smem, s>>>(Q.P); CKL(); }
        }
    }
    CK(cudaEventRecord(ctx->ev[2], s));
    return FKM_OK;
}

static constexpr int kRetryGlobal = 1;          // internal: the job must be redone by the global-table pipeline
static constexpr int kBinTooBig = 2;            // internal: the sort path met a bin of 2^32 k-mers

// ------------------------------------------------------------------ the partitioned count stage (fkm_part.cuh)
// Hash path with the tables in shared memory: bin-major records -> canonical k-mers, hash-partitioned into sub-buckets of a
// few thousand distinct k-mers (k_expand_hist / k_sub_scan / k_place_keys) -> k_count_keys.  Batches of bins whose k-mers
// fit the key buffer; the first B/64 bins are a synchronous sample sized for all-distinct k-mers, the distinct / k-mer
// ratio they show sizes the sub-buckets and the output regions of the rest, which is queued without a host sync.
// Returns kRetryGlobal when the input does not fit the scheme (a table or region overflowed, a bin of 2^32 k-mers).
template <bool WIDE>
static int count_partitioned(fkm_ctx* ctx, const fkm_config* cfg, int32_t B, const void* d_records, const unsigned long long* d_bin_base,
                             const std::vector<unsigned long long>& h_rec, const std::vector<unsigned long long>& h_kmer,
                             fkm_result* res, fkm_stats* st, unsigned long long* d_acc, uint64_t* out_total_p, float* ms_part_p, float* ms_count_p, const bool ordered,
                             const unsigned long long* d_bin_cnt = nullptr) {
    typedef typename Traits<WIDE>::Key Key;
    cudaStream_t s = ctx->stream;
    constexpr uint64_t TR = PartGeom<WIDE>::kTileRecs;
    // ordered (sort path): the sub-buckets of a bin are 2^j key ranges and k_count_keys_ordered's table keeps the k-mers in order
    auto retry = [&](const char* why, long long a = 0, long long b2 = 0) -> int {
        if (getenv("FKM_TRACE")) fprintf(stderr, "[fkm trace] partitioned count stage gives up: %s (%lld, %lld)\n", why, a, b2);
        return kRetryGlobal;
    };
    for (int b = 0; b < B; b++) if (h_kmer[(size_t)b] >= 0xFFFFFFF0ull) return retry("a bin of 2^32 k-mers", b);
    // geometry of k_count_keys's table
    // part_two_ctas: two CTAs of 512 threads per SM, each with a table of half the slots (experiment)
    const bool two = !ordered && ctx->part_two_ctas >= 1.0;
    uint32_t cap = (WIDE ? 8192u : 16384u) >> (two ? 1 : 0);
    while (cap > 64u && (size_t)cap * (sizeof(Key) + 6) + 2048 > ctx->smem_optin) cap >>= 1;
    if (ctx->smem_table_slots >= 64.0) while (cap > 64u && (double)cap > ctx->smem_table_slots) cap >>= 1;
    const uint32_t tail = ordered ? 512u : 0u;
    const size_t kc_smem = ordered ? (size_t)(cap + tail) * (sizeof(Key) + 4) : (size_t)cap * (sizeof(Key) + 6);
    const size_t sc_smem = (size_t)PartGeom<WIDE>::kBufKeys * (sizeof(Key) + 2);
    CK(cudaFuncSetAttribute(k_count_keys<WIDE, 1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)cap * (sizeof(Key) + 6))));
    CK(cudaFuncSetAttribute(k_count_keys<WIDE, 512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)cap * (sizeof(Key) + 6))));
    CK(cudaFuncSetAttribute(k_count_keys_ordered<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)(cap + 512u) * (sizeof(Key) + 4))));
    CK(cudaFuncSetAttribute(k_place_keys<WIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
    CK(cudaFuncSetAttribute(k_place_keys<WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
    static int occ_hist[2] = {0, 0}, occ_scat[2] = {0, 0};
    if (!occ_hist[WIDE]) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_hist[WIDE], k_expand_hist<WIDE, false>, PartGeom<WIDE>::kThreads, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_scat[WIDE], k_place_keys<WIDE, false>, PartGeom<WIDE>::kThreads, sc_smem));
        if (occ_hist[WIDE] < 1) occ_hist[WIDE] = 1;
        if (occ_scat[WIDE] < 1) occ_scat[WIDE] = 1;
    }
    // distinct k-mers a sub-bucket is planned for (key ranges are up to twice as full as the average: canonical k-mers crowd the low end of the key space)
    const double d_target = std::max(16.0, (double)cap * ctx->part_fill * (ordered ? 0.5 : 1.0));
    const uint64_t budget_keys = std::max<uint64_t>(1u << 16, (uint64_t)ctx->part_budget_keys);
    const unsigned grid = (unsigned)ctx->n_sm * (two ? 2u : 1u);
    const size_t bB = (size_t)B * 8;

    struct Batch { int lo, hi; size_t o32, o64h, o64k; uint64_t n_tiles, n_sub, hist_elems, n_keys, region_cap; double rho; };
    std::vector<Batch> batches;
    std::vector<uint32_t> h32;                  // per batch: tile_first[nb+1] | sub_first[nb+1]
    std::vector<unsigned long long> h64;        // per batch: hist_off[nb] | key_base[nb+1]
    auto plan = [&](int lo0, int hi_end, double rho_plan, double rho_out, bool one_batch) {
        for (int lo = lo0; lo < hi_end;) {
            Batch bt; bt.lo = lo; bt.n_tiles = bt.n_sub = bt.hist_elems = bt.n_keys = 0; bt.rho = rho_out;
            int hi = lo;
            while (hi < hi_end && (one_batch || hi == lo || bt.n_keys + h_kmer[(size_t)hi] <= budget_keys)) { bt.n_keys += h_kmer[(size_t)hi]; hi++; }
            bt.hi = hi;
            const int nb = hi - lo;
            bt.o32 = h32.size(); h32.resize(h32.size() + 2 * (size_t)(nb + 1));
            bt.o64h = h64.size(); h64.resize(h64.size() + (size_t)nb); bt.o64k = h64.size(); h64.resize(h64.size() + (size_t)nb + 1);
            uint32_t* tf = h32.data() + bt.o32; uint32_t* sf = tf + nb + 1;
            unsigned long long* ho = h64.data() + bt.o64h; unsigned long long* kb = h64.data() + bt.o64k;
            uint64_t keys = 0;
            for (int b = lo; b < hi; b++) {
                const uint64_t km = h_kmer[(size_t)b], nr = h_rec[(size_t)b];
                const uint64_t tiles = km ? (nr + TR - 1) / TR : 0;
                uint64_t subs = km ? (uint64_t)std::ceil((double)km * rho_plan / d_target) : 0;
                subs = km ? std::min<uint64_t>(std::max<uint64_t>(subs, 1), kPartMaxSubs) : 0;
                if (ordered && km) { uint64_t p2 = 1; while (p2 < subs) p2 <<= 1; subs = std::min<uint64_t>(p2, kPartMaxSubs); }
                tf[b - lo] = (uint32_t)bt.n_tiles; sf[b - lo] = (uint32_t)bt.n_sub; ho[b - lo] = bt.hist_elems; kb[b - lo] = keys;
                bt.n_tiles += tiles; bt.n_sub += subs; bt.hist_elems += tiles * subs; keys += km;
            }
            tf[nb] = (uint32_t)bt.n_tiles; sf[nb] = (uint32_t)bt.n_sub; kb[nb] = keys;
            // every CTA counts an equal share of the k-mers (plus at most one sub-bucket) into its own output region
            const uint64_t share = bt.n_keys / grid + 1;
            bt.region_cap = std::min<uint64_t>(share, (uint64_t)((double)share * std::min(1.0, rho_out * 1.25 + 0.01))) + 65536 + 4 * (uint64_t)cap;
            batches.push_back(bt);
            lo = hi;
            if (one_batch) break;
        }
    };

    unsigned long long *d_counters = nullptr, *d_cta_total = nullptr, *d_bin_off = nullptr; uint32_t* d_bin_cta = nullptr; int* d_flags = nullptr;
    void* d_slow_keys = nullptr; uint32_t* d_slow_cnt = nullptr;
    const uint64_t slow_slots = std::max<uint64_t>(1024, (uint64_t)ctx->smem_slow_slots);
    CK(dmalloc(ctx, &d_counters, 64)); CK(dmalloc(ctx, &d_flags, 16)); CK(dmalloc(ctx, &d_bin_cta, (size_t)B * 4)); CK(dmalloc(ctx, &d_bin_off, bB));
    CK(dmalloc(ctx, &d_slow_keys, (size_t)grid * slow_slots * sizeof(Key))); CK(dmalloc(ctx, &d_slow_cnt, (size_t)grid * slow_slots * 4));
    CK(cudaMemsetAsync(d_counters, 0, 64, s)); CK(cudaMemsetAsync(d_flags, 0, 16, s));

    uint64_t out_total = 0, n_sub_total = 0;
    float ms_part = 0, ms_count = 0;
    std::vector<unsigned long long> h_cta_total, h_bin_off((size_t)B);
    std::vector<uint32_t> h_bin_cta((size_t)B);
    // queues batches [b0, b1) of `batches` and waits for them; fills res->chunks / res->out_base
    auto run = [&](size_t b0, size_t b1) -> int {
        if (b0 >= b1) return FKM_OK;
        const size_t nbt = b1 - b0;
        uint64_t max_keys = 1, max_hist = 1, max_sub = 1;
        size_t lo32 = batches[b0].o32, lo64 = batches[b0].o64h;
        for (size_t i = b0; i < b1; i++) { max_keys = std::max(max_keys, batches[i].n_keys); max_hist = std::max(max_hist, batches[i].hist_elems); max_sub = std::max(max_sub, batches[i].n_sub); }
        uint32_t *d32 = nullptr, *d_hist = nullptr, *d_base = nullptr, *d_mid_bin = nullptr; unsigned long long *d64 = nullptr, *d_mid_key = nullptr; void* d_keys = nullptr;
        CK(dmalloc(ctx, &d32, (h32.size() - lo32) * 4 + 16)); CK(dmalloc(ctx, &d64, (h64.size() - lo64) * 8 + 16));
        CK(dmalloc(ctx, &d_hist, max_hist * 4)); CK(dmalloc(ctx, &d_base, max_hist * 4));
        CK(dmalloc(ctx, &d_keys, max_keys * sizeof(Key) + 64)); CK(dmalloc(ctx, &d_mid_key, (max_sub + 1) * 8)); CK(dmalloc(ctx, &d_mid_bin, max_sub * 4));
        void* d_keys_lin = nullptr; unsigned long long* d_tile_off = nullptr; uint32_t *d_tile_nk = nullptr, *d_kcur = nullptr;
        uint64_t max_tiles = 1, max_nb = 1;
        for (size_t i = b0; i < b1; i++) { max_tiles = std::max(max_tiles, batches[i].n_tiles); max_nb = std::max<uint64_t>(max_nb, (uint64_t)(batches[i].hi - batches[i].lo)); }
        CK(dmalloc(ctx, &d_keys_lin, (max_keys + max_tiles + max_nb + 8) * sizeof(Key))); CK(dmalloc(ctx, &d_tile_off, max_tiles * 8)); CK(dmalloc(ctx, &d_tile_nk, max_tiles * 4));
        CK(dmalloc(ctx, &d_kcur, max_nb * 4));
        CK(dmalloc(ctx, &d_cta_total, nbt * grid * 8));
        CK(cudaMemcpyAsync(d32, h32.data() + lo32, (h32.size() - lo32) * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d64, h64.data() + lo64, (h64.size() - lo64) * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(d_cta_total, 0, nbt * grid * 8, s));
        st->h2d_bytes += (h32.size() - lo32) * 4 + (h64.size() - lo64) * 8;
        while (ctx->evpool.size() < 3 * nbt) { cudaEvent_t e; CK(cudaEventCreate(&e)); ctx->evpool.push_back(e); }
        std::vector<void*> ok(nbt); std::vector<uint32_t*> oc(nbt);
        for (size_t i = b0; i < b1; i++) {
            const Batch& bt = batches[i];
            const int nb = bt.hi - bt.lo;
            CK(dmalloc(ctx, &ok[i - b0], (size_t)grid * bt.region_cap * sizeof(Key))); CK(dmalloc(ctx, &oc[i - b0], (size_t)grid * bt.region_cap * 4));
            CK(cudaEventRecord(ctx->evpool[3 * (i - b0)], s));
            if (bt.n_keys) {
                PartParams P;
                P.records = d_records; P.bin_rec_base = d_bin_base; P.bin_rec_cnt = d_bin_cnt; P.bin_lo = bt.lo; P.bin_hi = bt.hi;
                P.tile_first = d32 + (bt.o32 - lo32); P.sub_first = P.tile_first + nb + 1;
                P.hist_off = d64 + (bt.o64h - lo64); P.key_base = d64 + (bt.o64k - lo64);
                P.n_tiles = (uint32_t)bt.n_tiles; P.n_sub = (uint32_t)bt.n_sub; P.tile_hist = d_hist; P.tile_base = d_base; P.keys = d_keys;
                P.mid_key_base = d_mid_key; P.mid_bin = d_mid_bin; P.k = cfg->k;
                P.keys_lin = d_keys_lin; P.tile_key_off = d_tile_off; P.tile_nkeys = d_tile_nk; P.bin_key_cursor = d_kcur;
                CK(cudaMemsetAsync(d_kcur, 0, (size_t)nb * 4, s));
                const unsigned g1 = (unsigned)std::min<uint64_t>(bt.n_tiles, (uint64_t)ctx->n_sm * occ_hist[WIDE]);
                const unsigned g3 = (unsigned)std::min<uint64_t>(bt.n_tiles, (uint64_t)ctx->n_sm * occ_scat[WIDE]);
                if (ordered) k_expand_hist<WIDE, true><<<g1, PartGeom<WIDE>::kThreads, 0, s>>>(P); else k_expand_hist<WIDE, false><<<g1, PartGeom<WIDE>::kThreads, 0, s>>>(P);
                CKL();
                k_sub_scan<<<(unsigned)nb, 256, 0, s>>>(P); CKL();
                if (ordered) k_place_keys<WIDE, true><<<g3, PartGeom<WIDE>::kThreads, sc_smem, s>>>(P); else k_place_keys<WIDE, false><<<g3, PartGeom<WIDE>::kThreads, sc_smem, s>>>(P);
                CKL();
                CK(cudaEventRecord(ctx->evpool[3 * (i - b0) + 1], s));
                KeyCountParams Q;
                Q.keys = d_keys; Q.mid_key_base = d_mid_key; Q.mid_bin = d_mid_bin; Q.sub_first = P.sub_first; Q.bin_lo = bt.lo; Q.n_sub = (uint32_t)bt.n_sub;
                Q.out_keys = ok[i - b0]; Q.out_cnt = oc[i - b0]; Q.region_cap = bt.region_cap; Q.cta_total = d_cta_total + (i - b0) * grid;
                Q.bin_cta = d_bin_cta; Q.bin_off = d_bin_off; Q.acc = d_acc; Q.cap_slots = cap;
                Q.max_fill = cap > 4096u ? cap - 2048u : cap > 64u ? cap / 2 : 0u;        // (tiny test tables: the threads' claims in flight can exceed the table, see k_count_keys)
                Q.slow_keys = d_slow_keys; Q.slow_cnt = d_slow_cnt; Q.slow_slots = slow_slots; Q.slow_max_fill = slow_slots * 7 / 10; Q.flags = d_flags; Q.counters = d_counters;
                Q.k = cfg->k; Q.tail_slots = tail; Q.bin_shift = ctx->job_split;
                if (ordered) k_count_keys_ordered<WIDE><<<grid, kKcThreads, kc_smem, s>>>(Q);
                else if (cap < 4096u) k_count_keys<WIDE, 1024, true><<<grid, 1024, kc_smem, s>>>(Q);      // (test tables)
                else if (two) k_count_keys<WIDE, 512, false><<<grid, 512, kc_smem, s>>>(Q);
                else k_count_keys<WIDE, 1024, false><<<grid, 1024, kc_smem, s>>>(Q);
                CKL();
            } else CK(cudaEventRecord(ctx->evpool[3 * (i - b0) + 1], s));
            CK(cudaEventRecord(ctx->evpool[3 * (i - b0) + 2], s));
        }
        int flags[4] = {0, 0, 0, 0};
        h_cta_total.resize(nbt * grid);
        const int lo = batches[b0].lo, hi = batches[b1 - 1].hi;
        CK(cudaMemcpyAsync(h_cta_total.data(), d_cta_total, nbt * grid * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_bin_cta.data() + lo, d_bin_cta + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_bin_off.data() + lo, d_bin_off + lo, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(flags, d_flags, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        st->d2h_bytes += 16 + nbt * grid * 8 + (size_t)(hi - lo) * 12;
        for (size_t i = 0; i < nbt; i++) {
            float a = 0, c = 0;
            cudaEventElapsedTime(&a, ctx->evpool[3 * i], ctx->evpool[3 * i + 1]); cudaEventElapsedTime(&c, ctx->evpool[3 * i + 1], ctx->evpool[3 * i + 2]);
            ms_part += a; ms_count += c;
        }
        if (flags[2]) return fkm_set_error(FKM_EOVERFLOW, "a k-mer occurs more than 2^32-1 times: 32-bit counts overflow (the reference counts in Int, SBKC:562,676)");
        if (flags[0] || flags[1]) return retry(flags[0] ? "a table overflowed on the slow path too" : "an output region was too small", flags[0], flags[1]);
        // one result chunk per CTA region; a bin's entries begin in the region of the CTA that counted its first sub-bucket
        for (size_t i = b0; i < b1; i++) {
            const Batch& bt = batches[i];
            const uint32_t* sf = h32.data() + bt.o32 + (size_t)(bt.hi - bt.lo) + 1;
            std::vector<unsigned long long> cta_start((size_t)grid + 1);
            cta_start[0] = out_total;
            for (unsigned c = 0; c < grid; c++) {
                const unsigned long long n = h_cta_total[(i - b0) * grid + c];
                cta_start[(size_t)c + 1] = cta_start[(size_t)c] + n;
                if (n) {
                    Chunk ch; ch.keys = (char*)ok[i - b0] + (size_t)c * bt.region_cap * sizeof(Key); ch.cnt = oc[i - b0] + (size_t)c * bt.region_cap; ch.n = n;
                    res->chunks.push_back(ch);
                }
            }
            unsigned long long next = cta_start[(size_t)grid];
            for (int b = bt.hi - 1; b >= bt.lo; b--) {
                if (sf[b - bt.lo + 1] > sf[b - bt.lo]) next = cta_start[(size_t)h_bin_cta[(size_t)b]] + h_bin_off[(size_t)b];
                res->out_base[(size_t)b] = next;
            }
            out_total = cta_start[(size_t)grid];
            n_sub_total += bt.n_sub;
            st->n_batches++;
        }
        return FKM_OK;
    };

    // phase A: the sample, sized for all-distinct k-mers
    uint64_t km_all = 0; for (int b = 0; b < B; b++) km_all += h_kmer[(size_t)b];
    int s_hi = 0; uint64_t km_a = 0;            // (a prefix of the bins with 1/64 of the k-mers: a rank of a multi-GPU job owns only some of the bins)
    while (s_hi < B && (s_hi < 1 || km_a < km_all / 64)) km_a += h_kmer[(size_t)s_hi++];
    plan(0, s_hi, 1.0, 1.0, true);
    int rc = run(0, batches.size()); if (rc) return rc;
    if (s_hi < B) {
        double rho = 1.0;
        if (km_a > 100000) rho = std::min(1.0, (double)out_total / (double)km_a);
        rho *= ctx->debug_rho_scale;
        const size_t b0 = batches.size();
        const double rho_plan = std::min(1.0, rho * 1.1 + 0.01);
        // A tile's k-mers spread over all sub-buckets of its bin: beyond a few hundred sub-buckets the runs k_place_keys writes
        // shrink to a few keys and the per-(tile, sub-bucket) matrices outgrow the keys (8-GPU weak scaling, 23 M k-mers per
        // bin: partition 186 ms instead of 57).  Such bins are left to the global-table pipeline.
        for (int b = s_hi; b < B; b++)
            if ((double)h_kmer[(size_t)b] * rho_plan / d_target > ctx->part_max_subs) return retry("a bin needs too many sub-buckets", b, (long long)h_kmer[(size_t)b]);
        plan(s_hi, B, rho_plan, rho, false);
        rc = run(b0, batches.size()); if (rc) return rc;
    }
    res->out_base[(size_t)B] = out_total;
    unsigned long long h_counters[8];
    CK(cudaMemcpyAsync(h_counters, d_counters, 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    st->d2h_bytes += 64;
    st->n_mid_bins = n_sub_total; st->n_slow_bins = h_counters[0];
    *out_total_p = out_total; *ms_part_p = ms_part; *ms_count_p = ms_count;
    return FKM_OK;
}

// records that are already bin-major on this device (multi-GPU: after the exchange)
struct PreScattered { const void* d_records; const uint64_t* bin_rec; const uint64_t* bin_kmer; };

template <bool WIDE>
static int run_pipeline(fkm_ctx* ctx, const fkm_config* cfg, int32_t B, const void* d_bases, const void* d_inv,
                        uint64_t n_pos, fkm_result* res, fkm_stats* st, const PreScattered* pre, ScanState* scanned) {
    typedef typename Traits<WIDE>::Key Key;
    typedef typename Traits<WIDE>::Slot Slot;
    cudaStream_t s = ctx->stream;
    const int rec_bytes = Traits<WIDE>::kRecWords * 8;
    res->ctx = ctx; res->gen = ctx->gen;
    res->B = B; res->k = cfg->k; res->wide = WIDE; res->sorted = !cfg->use_ht; res->device = ctx->device;
    res->out_base.assign((size_t)B + 1, 0);

    unsigned long long *d_bin_base = nullptr, *d_cursor = nullptr,
                       *d_distinct = nullptr, *d_out_base = nullptr, *d_small = nullptr, *d_tbl_base = nullptr;
    void* d_records = nullptr; void* d_table = nullptr; int* d_ovf = nullptr; unsigned long long* d_acc = nullptr;
    void *d_keysA = nullptr, *d_keysB = nullptr; unsigned int *d_tile_seg = nullptr, *d_seg_tile0 = nullptr, *d_tile_hist = nullptr, *d_tile_heads = nullptr;
    unsigned long long* d_first = nullptr;
    int rc = FKM_OK;
    // everything below frees through this lambda
    auto cleanup = [&]() {
        dfree(ctx, d_bin_base); dfree(ctx, d_cursor); dfree(ctx, d_distinct);
        dfree(ctx, d_out_base); dfree(ctx, d_small); dfree(ctx, d_tbl_base); dfree(ctx, d_records); dfree(ctx, d_table); dfree(ctx, d_ovf);
        dfree(ctx, d_keysA); dfree(ctx, d_keysB); dfree(ctx, d_tile_seg); dfree(ctx, d_seg_tile0); dfree(ctx, d_tile_hist); dfree(ctx, d_tile_heads); dfree(ctx, d_first);
    };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
    return fkm_set_error(e_ == cudaErrorMemoryAllocation ? FKM_ENOMEM : FKM_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } } while (0)
#define CKLC() do { g_launches++; ctx->job_launches++; CKC(cudaGetLastError()); } while (0)

    Trace tr;
    const size_t bB = (size_t)B * 8;
    CKC(dmalloc(ctx, &d_bin_base, bB + 8));
    CKC(dmalloc(ctx, &d_cursor, bB)); CKC(dmalloc(ctx, &d_distinct, bB)); CKC(dmalloc(ctx, &d_out_base, bB + 8));
    CKC(dmalloc(ctx, &d_small, 64)); CKC(dmalloc(ctx, &d_ovf, 8)); CKC(dmalloc(ctx, &d_tbl_base, bB + 8));
    CKC(dmalloc(ctx, &d_acc, 192 * 8));
    CKC(cudaMemsetAsync(d_acc, 0, 192 * 8, s));
    CKC(cudaMemsetAsync(d_cursor, 0, bB, s)); CKC(cudaMemsetAsync(d_distinct, 0, bB, s));
    CKC(cudaMemsetAsync(d_out_base, 0, bB + 8, s)); CKC(cudaMemsetAsync(d_small, 0, 64, s));

    std::vector<unsigned long long> h_rec, h_kmer, h_base((size_t)B + 1);
    ScanState local_scan;
    ScanState& scan = scanned ? *scanned : local_scan;
    if (!pre) {
        if (!scanned) { rc = stage_scan(ctx, cfg, B, d_bases, d_inv, n_pos, &scan, st); if (rc) return rc; }
        else CKC(cudaEventRecord(ctx->ev[0], s));
        h_rec = scan.h_rec; h_kmer = scan.h_kmer;
        tr.mark("histogram done");
    } else {
        h_rec.assign(pre->bin_rec, pre->bin_rec + B); h_kmer.assign(pre->bin_kmer, pre->bin_kmer + B);
        CKC(cudaEventRecord(ctx->ev[0], s));
    }
    uint64_t n_rec = 0, n_kmers = 0, nonempty = 0;
    for (int b = 0; b < B; b++) { h_base[(size_t)b] = n_rec; n_rec += h_rec[(size_t)b]; n_kmers += h_kmer[(size_t)b]; nonempty += h_rec[(size_t)b] ? 1 : 0; }
    h_base[(size_t)B] = n_rec;
    st->n_kmers = n_kmers; st->n_superkmers = n_rec; st->superkmer_bytes = n_rec * rec_bytes; st->n_nonempty_bins = nonempty;
    CKC(cudaMemcpyAsync(d_bin_base, h_base.data(), bB + 8, cudaMemcpyHostToDevice, s));
    st->h2d_bytes += bB + 8;
    uint64_t out_total = 0;
    float ms_count = 0, ms_compact = 0, ms_part = 0;
    bool part_done = false, scattered = false;
    // ---- the records were scattered while the input was still arriving (speculative scatter of the FASTA front end): the bins'
    // regions have gaps, which only the partitioned count stage reads around; anything else scatters again below
    if (!pre && scanned && scanned->spec.on && !scanned->spec.failed && cfg->use_ht && ctx->count_mode >= 2.0 && n_rec) {
        unsigned long long* d_bin_cnt = nullptr;
        CKC(dmalloc(ctx, &d_bin_cnt, bB));
        CKC(cudaMemcpyAsync(d_bin_cnt, h_rec.data(), bB, cudaMemcpyHostToDevice, s));
        CKC(cudaEventRecord(ctx->ev[1], s)); CKC(cudaEventRecord(ctx->ev[2], s));
        const Arena::Mark mk = ctx->arena.mark();
        rc = count_partitioned<WIDE>(ctx, cfg, B, scanned->spec.d_records, scanned->spec.d_bin_base, h_rec, h_kmer, res, st, d_acc, &out_total, &ms_part, &ms_count,
                                     false, d_bin_cnt);
        if (rc == FKM_OK) { part_done = true; scattered = true; st->superkmer_bytes = scanned->spec.base[(size_t)B] * rec_bytes; }
        else if (rc != kRetryGlobal) { cleanup(); return rc; }
        else {
            rc = FKM_OK;
            ctx->arena.release(mk);
            st->n_batches = 0; st->n_mid_bins = 0; st->n_slow_bins = 0;
            res->chunks.clear(); out_total = 0; ms_count = 0; ms_part = 0;
            CKC(cudaMemsetAsync(d_acc, 0, 192 * 8, s));
        }
    }
    if (scattered) {
    } else if (!pre) {
        CKC(dmalloc(ctx, &d_records, std::max<size_t>(16, (size_t)n_rec * rec_bytes)));
        tr.mark("records allocated", (long long)n_rec);
        rc = stage_scatter(ctx, &scan, d_bin_base, d_records, n_rec, st); if (rc) return rc;
    } else {
        d_records = const_cast<void*>(pre->d_records);
        CKC(cudaEventRecord(ctx->ev[2], s));
    }

    // ---- hash path, count_mode 2: the partitioned count stage with its tables in shared memory (fkm_part.cuh)
    if (!part_done && (cfg->use_ht ? ctx->count_mode >= 2.0 : (ctx->sort_partition >= 2.0 && ctx->debug_force_lsd < 1.0)) && n_rec) {
        const Arena::Mark mk = ctx->arena.mark();
        rc = count_partitioned<WIDE>(ctx, cfg, B, d_records, d_bin_base, h_rec, h_kmer, res, st, d_acc, &out_total, &ms_part, &ms_count, !cfg->use_ht);
        if (rc == FKM_OK) part_done = true;
        else if (rc != kRetryGlobal) { cleanup(); return rc; }
        else {      // the global-table pipeline below redoes the count
            rc = FKM_OK;
            ctx->arena.release(mk);
            st->n_fallbacks++; st->n_batches = 0; st->n_mid_bins = 0; st->n_slow_bins = 0;
            res->chunks.clear(); out_total = 0; ms_count = 0; ms_part = 0;
            CKC(cudaMemsetAsync(d_acc, 0, 192 * 8, s));
        }
    }
    st->ms_partition = ms_part;
    const bool want_fold = !part_done && !WIDE && cfg->use_ht && ctx->fold_records >= 1.0 && n_rec >= 4096;

    // ---- stage 2b (hash path, 64-bit k-mers, optional): fold identical records.  Reads of a deeply sequenced
    // genome repeat its super-k-mers once per covering read (either strand); every k-mer of a repeated record costs
    // one atomic in k_count_ht.  Folding costs one 128-bit insertion per RECORD and lets the count stage add the
    // multiplicity with one atomic per k-mer of a DISTINCT record.  Batches of bins like the count stage; the first
    // batch is synchronous and decides whether the input repeats enough for the rest to pay.
    const void* d_count_records = d_records;      // what the count stage reads
    const uint32_t* d_weights = nullptr;
    float ms_fold = 0;
    if constexpr (!WIDE) {
        if (want_fold) {
            const uint64_t budget_slots = std::max<uint64_t>(1024, (uint64_t)(ctx->fold_table_bytes / sizeof(SlotW)));
            struct FoldBatch { int lo, hi; size_t tb_idx; uint64_t slots; };
            std::vector<FoldBatch> fbs; std::vector<unsigned long long> ftb;
            // bins [lo0, B) -> batches; a bin's table holds share * records / 0.6 slots.  The sampled first batch is
            // small and sized for all-distinct records (share 1); the rest is sized from the share it observed.
            auto plan = [&](int lo0, double share, bool sample_only) -> uint64_t {
                uint64_t max_slots = 1024;
                for (int lo = lo0; lo < B;) {
                    FoldBatch fb; fb.lo = lo; fb.tb_idx = ftb.size(); ftb.push_back(0);
                    int hi = lo; uint64_t slots = 0;
                    while (hi < B) {
                        const uint64_t c = h_rec[(size_t)hi];
                        const uint64_t sz = c ? round_up(std::max<uint64_t>((uint64_t)((double)c * share / 0.6) + 1, 1024), 1024) : 0;
                        if (hi > lo && slots + sz > budget_slots) break;
                        slots += sz; ftb.push_back(slots); hi++;
                        if (sample_only && hi - lo >= std::max(1, B / 64) && slots >= budget_slots / 32) break;
                    }
                    fb.hi = hi; fb.slots = slots; fbs.push_back(fb); max_slots = std::max(max_slots, slots); lo = hi;
                    if (sample_only) break;
                }
                return max_slots;
            };
            unsigned long long *d_ftb = nullptr, *d_fdistinct = nullptr, *d_fbase = nullptr, *d_fcursor = nullptr;
            void* d_frec = nullptr; uint32_t* d_fwt = nullptr;
            CKC(dmalloc(ctx, &d_ftb, ((size_t)2 * B + 8) * 8)); CKC(dmalloc(ctx, &d_fdistinct, bB)); CKC(dmalloc(ctx, &d_fbase, bB + 8)); CKC(dmalloc(ctx, &d_fcursor, bB));
            CKC(dmalloc(ctx, &d_frec, (size_t)n_rec * 16)); CKC(dmalloc(ctx, &d_fwt, (size_t)n_rec * 4));
            CKC(cudaMemsetAsync(d_fdistinct, 0, bB, s)); CKC(cudaMemsetAsync(d_fbase, 0, bB + 8, s)); CKC(cudaMemsetAsync(d_fcursor, 0, bB, s));
            CKC(cudaMemsetAsync(d_ovf, 0, 8, s));
            CKC(cudaEventRecord(ctx->ev[6], s));
            bool keep = true;
            size_t done = 0;                               // batches already queued
            auto run_batches = [&](void* d_ftab) -> int {
                for (; done < fbs.size(); done++) {
                    const FoldBatch& fb = fbs[done];
                    FoldParams F;
                    F.records = d_records; F.rec_lo = h_base[(size_t)fb.lo]; F.rec_hi = h_base[(size_t)fb.hi];
                    F.bin_base = d_bin_base; F.bin_lo = fb.lo; F.bin_hi = fb.hi; F.table = d_ftab; F.tbl_base = d_ftb + fb.tb_idx;
                    F.bin_distinct = d_fdistinct; F.overflow = d_ovf; F.max_probe = 512; F.k = cfg->k;
                    const uint64_t nr = F.rec_hi - F.rec_lo;
                    if (nr) {
                        const int pool = ctx->fold_pool >= 256.0 ? 256 : ctx->fold_pool >= 128.0 ? 128 : 64;   // records per warp
                        const unsigned grid = (unsigned)((nr + 8 * (uint64_t)pool - 1) / (8 * (uint64_t)pool));
                        if (pool == 256) k_fold_insert<256><<<grid, 256, 0, s>>>(F);
                        else if (pool == 128) k_fold_insert<128><<<grid, 256, 0, s>>>(F);
                        else k_fold_insert<64><<<grid, 256, 0, s>>>(F);
                        CKLC();
                    }
                    k_bin_offsets<<<1, 256, 0, s>>>(d_fdistinct, d_fbase, fb.lo, fb.hi, d_small); CKLC();
                    if (fb.slots) {
                        CompactParams Q;
                        Q.table = d_ftab; Q.n_slots = fb.slots; Q.tbl_base = d_ftb + fb.tb_idx; Q.n_bins = fb.hi - fb.lo; Q.bin_lo = fb.lo;
                        Q.out_base = d_fbase; Q.out_cursor = d_fcursor; Q.out_origin = 0; Q.out_cap = n_rec;
                        Q.out_keys = d_frec; Q.out_cnt = d_fwt; Q.clear = 1; Q.cap_overflow = d_ovf + 1; Q.acc = nullptr; Q.bin_shift = 0;
                        const unsigned grid = (unsigned)std::min<uint64_t>((fb.slots + 1023) / 1024, compact_grid<true>(ctx));
                        k_compact_ht<true><<<grid, 256, 0, s>>>(Q); CKLC();
                    }
                }
                return FKM_OK;
            };
            {   // the sample
                const uint64_t slots0 = plan(0, 1.0, true);
                void* d_ftab0 = nullptr;
                CKC(dmalloc(ctx, &d_ftab0, (size_t)slots0 * sizeof(SlotW)));
                CKC(cudaMemcpyAsync(d_ftb, ftb.data(), ftb.size() * 8, cudaMemcpyHostToDevice, s));
                st->h2d_bytes += ftb.size() * 8;
                k_fill_table<true><<<ctx->n_sm * 8, 256, 0, s>>>(d_ftab0, slots0); CKLC();
                rc = run_batches(d_ftab0); if (rc) { cleanup(); return rc; }
            }
            const int lo1 = fbs[0].hi;
            if (lo1 < B) {
                unsigned long long folded = 0;
                CKC(cudaMemcpyAsync(&folded, d_small, 8, cudaMemcpyDeviceToHost, s));
                CKC(cudaStreamSynchronize(s));
                st->d2h_bytes += 8;
                const uint64_t nr0 = h_base[(size_t)lo1] - h_base[0];
                const double share0 = nr0 ? (double)folded / (double)nr0 : 1.0;
                if (share0 > ctx->fold_max_ratio) keep = false;
                else {
                    const size_t first_new = ftb.size();
                    const uint64_t slots1 = plan(lo1, std::min(1.0, share0 * 1.3 + 0.03), false);
                    void* d_ftab1 = nullptr;
                    CKC(dmalloc(ctx, &d_ftab1, (size_t)slots1 * sizeof(SlotW)));
                    CKC(cudaMemcpyAsync(d_ftb + first_new, ftb.data() + first_new, (ftb.size() - first_new) * 8, cudaMemcpyHostToDevice, s));
                    st->h2d_bytes += (ftb.size() - first_new) * 8;
                    k_fill_table<true><<<ctx->n_sm * 8, 256, 0, s>>>(d_ftab1, slots1); CKLC();
                    rc = run_batches(d_ftab1); if (rc) { cleanup(); return rc; }
                }
            }
            std::vector<unsigned long long> h_fbase((size_t)B + 1);
            int flags[2] = {0, 0};
            if (keep) {
                CKC(cudaMemcpyAsync(h_fbase.data(), d_fbase, bB + 8, cudaMemcpyDeviceToHost, s));
                CKC(cudaMemcpyAsync(flags, d_ovf, 8, cudaMemcpyDeviceToHost, s));
            }
            CKC(cudaEventRecord(ctx->ev[7], s));
            CKC(cudaStreamSynchronize(s));
            cudaEventElapsedTime(&ms_fold, ctx->ev[6], ctx->ev[7]);
            if (keep && !flags[0] && !flags[1]) {
                st->d2h_bytes += bB + 16;
                h_base = h_fbase;                          // the count stage now sees the folded records
                CKC(cudaMemcpyAsync(d_bin_base, d_fbase, bB + 8, cudaMemcpyDeviceToDevice, s));
                d_count_records = d_frec; d_weights = d_fwt;
                st->n_folded_records = h_fbase[(size_t)B];
            }
            CKC(cudaMemsetAsync(d_ovf, 0, 8, s));
            CKC(cudaMemsetAsync(d_small, 0, 64, s));
        }
    }
    st->ms_fold = ms_fold;

    // ---- stage 3/4: per-bin exact count, batches of consecutive bins
    if (part_done) {
    } else if (cfg->use_ht) {
        double rho = 1.0;                         // sizing estimate of distinct / k-mers (with margin), learnt from the first batch
        double rho_obs = 1.0;                     // the ratio actually observed
        uint64_t table_cap = 0;
        std::vector<unsigned long long> tb;

        // ---- synchronous batches (DRAM-sized tables, exact output allocation, retry on overflow).
        // Used for the first batch (which measures rho) and as the fallback of the fast path.
        auto safe_batches = [&](int lo, const int hi_end, const bool one_batch, int* next_lo) -> int {
            const uint64_t budget_slots = std::max<uint64_t>(1024, (uint64_t)(ctx->table_budget_bytes / sizeof(Slot)));
            while (lo < hi_end) {
                bool retried = false;
                double use_rho = rho;
                int hi;
                for (;;) {
                    tb.clear(); tb.push_back(0);
                    hi = lo; uint64_t slots = 0;
                    while (hi < hi_end) {
                        uint64_t want = (uint64_t)((double)h_kmer[(size_t)hi] * use_rho / ctx->load_factor) + 1;
                        uint64_t sz = h_kmer[(size_t)hi] ? round_up(std::max<uint64_t>(want, 1024), 1024) : 0;
                        if (hi > lo && slots + sz > budget_slots) break;
                        slots += sz; tb.push_back(slots); hi++;
                        if (one_batch && hi - lo >= std::max(1, B / 64) && slots * sizeof(Slot) > (64ull << 20)) break;   // small first batch to learn rho
                    }
                    if (slots > table_cap) { CKC(dmalloc(ctx, &d_table, (size_t)slots * sizeof(Slot))); table_cap = slots; }
                    CKC(cudaEventRecord(ctx->ev[6], s));
                    if (slots) { k_fill_table<WIDE><<<ctx->n_sm * 8, 256, 0, s>>>(d_table, slots); CKLC(); }
                    CKC(cudaMemsetAsync(d_ovf, 0, 8, s));
                    CKC(cudaMemcpyAsync(d_tbl_base, tb.data(), tb.size() * 8, cudaMemcpyHostToDevice, s));
                    st->h2d_bytes += tb.size() * 8;
                    CountParams C;
                    C.records = d_count_records; C.weights = d_weights; C.rec_lo = h_base[(size_t)lo]; C.rec_hi = h_base[(size_t)hi];
                    C.bin_base = d_bin_base; C.bin_lo = lo; C.bin_hi = hi; C.table = d_table; C.tbl_base = d_tbl_base;
                    C.bin_distinct = d_distinct; C.overflow = d_ovf; C.k = cfg->k; C.max_probe = 512; C.first_state = ctx->cas_first >= 1.0 ? 1 : 0;
                    const uint64_t nr = C.rec_hi - C.rec_lo;
                    if (nr) { k_count_ht<WIDE><<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(C); CKLC(); }
                    k_bin_offsets<<<1, 256, 0, s>>>(d_distinct, d_out_base, lo, hi, d_small); CKLC();
                    CKC(cudaEventRecord(ctx->ev[7], s));
                    unsigned long long batch_total = 0; int ovf = 0;
                    CKC(cudaMemcpyAsync(&batch_total, d_small, 8, cudaMemcpyDeviceToHost, s));
                    CKC(cudaMemcpyAsync(&ovf, d_ovf, 4, cudaMemcpyDeviceToHost, s));
                    CKC(cudaStreamSynchronize(s));
                    st->d2h_bytes += 12;
                    { float ms; cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]); ms_count += ms; }
                    if (ovf) {
                        if (retried && use_rho >= 1.0) return fkm_set_error(FKM_ECUDA, "hash table overflow at full size (bins %d..%d)", lo, hi);
                        CKC(cudaMemsetAsync(d_distinct + lo, 0, (size_t)(hi - lo) * 8, s));
                        use_rho = retried ? 1.0 : std::min(1.0, use_rho * 2.0); retried = true;
                        continue;
                    }
                    uint64_t nk = 0; for (int b = lo; b < hi; b++) nk += h_kmer[(size_t)b];
                    if (nk > 100000) { rho_obs = (double)batch_total / (double)nk; rho = std::min(1.0, rho_obs * 1.25 + 0.02); }
                    Chunk ch; ch.n = batch_total;
                    if (batch_total) {
                        CKC(dmalloc(ctx, &ch.keys, (size_t)batch_total * sizeof(Key)));
                        CKC(dmalloc(ctx, &ch.cnt, (size_t)batch_total * 4));
                        res->chunks.push_back(ch);
                        CompactParams Q;
                        Q.table = d_table; Q.n_slots = slots; Q.tbl_base = d_tbl_base; Q.n_bins = hi - lo; Q.bin_lo = lo;
                        Q.out_base = d_out_base; Q.out_cursor = d_cursor; Q.out_origin = out_total; Q.out_cap = batch_total;
                        Q.out_keys = ch.keys; Q.out_cnt = ch.cnt; Q.clear = 0; Q.cap_overflow = d_ovf + 1; Q.acc = d_acc; Q.bin_shift = ctx->job_split;
                        CKC(cudaMemsetAsync(d_cursor + lo, 0, (size_t)(hi - lo) * 8, s));
                        CKC(cudaEventRecord(ctx->ev[6], s));
                        const unsigned grid = (unsigned)std::min<uint64_t>((slots + 1023) / 1024, compact_grid<WIDE>(ctx));
                        k_compact_ht<WIDE><<<grid, 256, 0, s>>>(Q); CKLC();
                        CKC(cudaEventRecord(ctx->ev[7], s));
                        CKC(cudaStreamSynchronize(s));
                        { float ms; cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]); ms_compact += ms; }
                    }
                    out_total += batch_total;
                    st->n_batches++;
                    break;
                }
                lo = hi;
                if (one_batch) break;
            }
            *next_lo = lo;
            return FKM_OK;
        };

        // ---- asynchronous batches: all launches are queued without a
        // host sync; tables are sized from rho, the output arrays from rho_obs; the compaction
        // kernel clears the slots it reads, so one fill serves every batch.  Any overflow
        // (table or output) is caught by flags read once at the end; the caller then redoes
        // the bins with safe_batches.
        auto fast_batches = [&](const int lo0, bool* ok) -> int {
            *ok = false;
            const uint64_t budget_slots = std::max<uint64_t>(1024, (uint64_t)(ctx->async_table_bytes / sizeof(Slot)));
            struct Batch { int lo, hi; size_t tb_idx; uint64_t slots; };
            std::vector<Batch> batches; std::vector<unsigned long long> tb_all;
            uint64_t max_slots = 0, nk_rest = 0;
            for (int lo = lo0; lo < B;) {
                Batch bt; bt.lo = lo; bt.tb_idx = tb_all.size(); tb_all.push_back(0);
                int hi = lo; uint64_t slots = 0;
                while (hi < B) {
                    uint64_t want = (uint64_t)((double)h_kmer[(size_t)hi] * rho / ctx->load_factor) + 1;
                    uint64_t sz = h_kmer[(size_t)hi] ? round_up(std::max<uint64_t>(want, 1024), 1024) : 0;
                    if (hi > lo && slots + sz > budget_slots) break;
                    slots += sz; tb_all.push_back(slots); nk_rest += h_kmer[(size_t)hi]; hi++;
                }
                bt.hi = hi; bt.slots = slots; batches.push_back(bt);
                max_slots = std::max(max_slots, slots);
                lo = hi;
            }
            const uint64_t out_cap = std::min<uint64_t>(nk_rest, (uint64_t)((double)nk_rest * rho_obs * 1.3) + (1u << 20));
            unsigned long long* d_tb_all = nullptr; void* d_tab = nullptr;
            Chunk ch;
            CKC(dmalloc(ctx, &d_tb_all, tb_all.size() * 8));
            CKC(dmalloc(ctx, &d_tab, (size_t)std::max<uint64_t>(max_slots, 1024) * sizeof(Slot)));
            CKC(dmalloc(ctx, &ch.keys, (size_t)std::max<uint64_t>(out_cap, 1) * sizeof(Key)));
            CKC(dmalloc(ctx, &ch.cnt, (size_t)std::max<uint64_t>(out_cap, 1) * 4));
            CKC(cudaMemcpyAsync(d_tb_all, tb_all.data(), tb_all.size() * 8, cudaMemcpyHostToDevice, s));
            st->h2d_bytes += tb_all.size() * 8;
            k_fill_table<WIDE><<<ctx->n_sm * 8, 256, 0, s>>>(d_tab, std::max<uint64_t>(max_slots, 1024)); CKLC();
            CKC(cudaMemsetAsync(d_ovf, 0, 8, s));
            CKC(cudaMemsetAsync(d_cursor + lo0, 0, (size_t)(B - lo0) * 8, s));
            CKC(cudaEventRecord(ctx->ev[6], s));
            const int n_samples = 8; int sampled = 0;
            const size_t stride = std::max<size_t>(1, batches.size() / n_samples);
            for (size_t bi = 0; bi < batches.size(); bi++) {
                const Batch& bt = batches[bi];
                const bool sample = (bi % stride == stride / 2) && sampled < n_samples;
                CountParams C;
                C.records = d_count_records; C.weights = d_weights; C.rec_lo = h_base[(size_t)bt.lo]; C.rec_hi = h_base[(size_t)bt.hi];
                C.bin_base = d_bin_base; C.bin_lo = bt.lo; C.bin_hi = bt.hi; C.table = d_tab; C.tbl_base = d_tb_all + bt.tb_idx;
                C.bin_distinct = d_distinct; C.overflow = d_ovf; C.k = cfg->k; C.max_probe = 512; C.first_state = ctx->cas_first >= 1.0 ? 1 : 0;
                const uint64_t nr = C.rec_hi - C.rec_lo;
                if (sample) CKC(cudaEventRecord(ctx->evs[3 * sampled], s));
                if (nr) { k_count_ht<WIDE><<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(C); CKLC(); }
                if (sample) CKC(cudaEventRecord(ctx->evs[3 * sampled + 1], s));
                k_bin_offsets<<<1, 256, 0, s>>>(d_distinct, d_out_base, bt.lo, bt.hi, d_small); CKLC();
                if (bt.slots) {
                    CompactParams Q;
                    Q.table = d_tab; Q.n_slots = bt.slots; Q.tbl_base = d_tb_all + bt.tb_idx; Q.n_bins = bt.hi - bt.lo; Q.bin_lo = bt.lo;
                    Q.out_base = d_out_base; Q.out_cursor = d_cursor; Q.out_origin = out_total; Q.out_cap = out_cap;
                    Q.out_keys = ch.keys; Q.out_cnt = ch.cnt; Q.clear = 1; Q.cap_overflow = d_ovf + 1; Q.acc = d_acc; Q.bin_shift = ctx->job_split;
                    const unsigned grid = (unsigned)std::min<uint64_t>((bt.slots + 1023) / 1024, compact_grid<WIDE>(ctx));
                    k_compact_ht<WIDE><<<grid, 256, 0, s>>>(Q); CKLC();
                }
                if (sample) { CKC(cudaEventRecord(ctx->evs[3 * sampled + 2], s)); sampled++; }
            }
            CKC(cudaEventRecord(ctx->ev[7], s));
            int flags[2] = {0, 0}; unsigned long long end_off = 0;
            CKC(cudaMemcpyAsync(flags, d_ovf, 8, cudaMemcpyDeviceToHost, s));
            CKC(cudaMemcpyAsync(&end_off, d_out_base + B, 8, cudaMemcpyDeviceToHost, s));
            CKC(cudaStreamSynchronize(s));
            st->d2h_bytes += 16;
            float ms_all = 0; cudaEventElapsedTime(&ms_all, ctx->ev[6], ctx->ev[7]);
            float sc = 0, sp = 0;
            for (int i = 0; i < sampled; i++) {
                float a1 = 0, a2 = 0;
                cudaEventElapsedTime(&a1, ctx->evs[3 * i], ctx->evs[3 * i + 1]); cudaEventElapsedTime(&a2, ctx->evs[3 * i + 1], ctx->evs[3 * i + 2]);
                sc += a1; sp += a2;
            }
            const float share = (sc + sp) > 0 ? sc / (sc + sp) : 1.0f;      // split of the phase between count and offsets+compact, from the sampled batches
            ms_count += ms_all * share; ms_compact += ms_all * (1.0f - share);
            if (flags[0] || flags[1]) {            // undo: the caller redoes [lo0, B) synchronously
                CKC(cudaMemsetAsync(d_distinct + lo0, 0, (size_t)(B - lo0) * 8, s));
                CKC(cudaMemsetAsync(d_cursor + lo0, 0, (size_t)(B - lo0) * 8, s));
                st->n_fallbacks++;
                return FKM_OK;
            }
            ch.n = end_off - out_total;
            res->chunks.push_back(ch);
            out_total = end_off;
            st->n_batches += batches.size();
            *ok = true;
            return FKM_OK;
        };

        int lo = 0;
        rc = safe_batches(0, B, true, &lo); if (rc) { cleanup(); return rc; }
        bool fast_ok = false;
        const double rho_keep = rho, rho_obs_keep = rho_obs;
        rho *= ctx->debug_rho_scale; rho_obs *= ctx->debug_rho_scale;
        if (lo < B && ctx->async_table_bytes >= 1.0) {
            // the fast path folds its digest into the accumulators: keep a copy to roll back on fallback
            unsigned long long* d_acc_keep = nullptr;
            CKC(dmalloc(ctx, &d_acc_keep, 192 * 8));
            CKC(cudaMemcpyAsync(d_acc_keep, d_acc, 192 * 8, cudaMemcpyDeviceToDevice, s));
            rc = fast_batches(lo, &fast_ok); if (rc) { cleanup(); return rc; }
            if (!fast_ok) CKC(cudaMemcpyAsync(d_acc, d_acc_keep, 192 * 8, cudaMemcpyDeviceToDevice, s));
        }
        rho = rho_keep; rho_obs = rho_obs_keep;
        if (!fast_ok && lo < B) { rc = safe_batches(lo, B, false, &lo); if (rc) { cleanup(); return rc; } }
    } else {
        // The sort kernels index a bin's keys with 32 bits (as the reference indexes its arrays with Int: a bin of 2^31 (k,x)-mers
        // cannot exist there, SBKC:484-542): count_device counts such an input with the hash path instead, to tell a real 32-bit
        // count overflow (FKM_EOVERFLOW, as on every other path) from a configuration that only needs more bins.
        for (int b = 0; b < B; b++)
            if (h_kmer[(size_t)b] >= 0xFFFFFFF0ull) { cleanup(); return kBinTooBig; }
        const uint64_t budget_keys = std::max<uint64_t>(kSortTile, (uint64_t)ctx->sort_budget_keys);
        const int n_pass = (2 * cfg->k + 7) / 8;
        const uint64_t local_cap = LocalSort<WIDE>::kCap;
        const size_t local_smem = (size_t)2 * local_cap * sizeof(Key) + local_cap * 2 + 16 * 256 * 4;
        CKC(cudaFuncSetAttribute(k_radix_local<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)local_smem));
        // plan every batch first, so that each scratch array is allocated once at its largest size
        struct SortBatch { int lo, hi; uint64_t nk, n_tiles, max_bin; int bits; };   // bits: 0 no partition, 8/10 MSD partition, -1 LSD passes
        std::vector<SortBatch> plan;
        uint64_t max_nk = 0, max_tiles = 0, max_hist = 0, max_sdt = 1, max_chunks = 1;
        for (int lo = 0; lo < B;) {
            SortBatch sb; sb.lo = lo; sb.nk = 0; sb.n_tiles = 0; sb.max_bin = 0;
            int hi = lo;
            while (hi < B) {
                const uint64_t c = h_kmer[(size_t)hi];
                if (hi > lo && sb.nk + c > budget_keys) break;
                sb.nk += c; sb.n_tiles += (c + kSortTile - 1) / kSortTile; sb.max_bin = std::max(sb.max_bin, c);
                hi++;
            }
            sb.hi = hi;
            const uint64_t need = (sb.max_bin * 4 + local_cap - 1) / local_cap;      // sub-buckets wanted (4x head-room for skew)
            // measured (profiles/README.md): the shared-memory chunk sort wins for 64-bit keys (config 3: 189 -> 158 ms),
            // not for 128-bit keys (config 4 shape: 1258 -> 1328 ms), which keep the LSD passes.  debug_force_lsd: 1 = LSD, 2 = MSD.
            const bool want_msd = ctx->debug_force_lsd >= 2.0 ? true : ctx->debug_force_lsd >= 1.0 ? false : !WIDE;
            sb.bits = !want_msd ? -1 : need <= 1 ? 0 : need <= 256 ? 8 : need <= 1024 ? 10 : -1;
            if (sb.bits > 2 * cfg->k) sb.bits = -1;
            const uint64_t nd = sb.bits > 0 ? (1ull << sb.bits) : 256;
            max_nk = std::max(max_nk, sb.nk); max_tiles = std::max(max_tiles, sb.n_tiles);
            max_hist = std::max(max_hist, sb.n_tiles * nd);
            if (sb.bits >= 0) {
                max_sdt = std::max<uint64_t>(max_sdt, (uint64_t)(hi - lo) * (sb.bits ? nd : 1));
                max_chunks = std::max<uint64_t>(max_chunks, 2 * sb.nk / local_cap + (uint64_t)(hi - lo) + 16);
            }
            plan.push_back(sb);
            lo = hi;
        }
        // Partitioned sort (fkm_part.cuh kernels in ORDERED mode): the k-mers are expanded once with coalesced stores
        // (k_expand_hist), cut into sub-buckets by their top bits (k_sub_scan + k_place_keys: one read + one write), and
        // runs of consecutive sub-buckets are sorted completely in shared memory (k_radix_local), for both key widths.
        // A bin needs 3 * k-mers / chunk capacity sub-buckets (canonical k-mers crowd the low end of the key space: at most
        // 2x the uniform density), rounded up to a power of two; above kPartMaxSubs the batch keeps the passes below.
        constexpr uint64_t PTR = PartGeom<WIDE>::kTileRecs;
        auto part_subs = [&](uint64_t c) -> uint64_t { uint64_t need = (c * 3 + local_cap - 1) / local_cap, p2 = 1; while (p2 < need) p2 <<= 1; return c ? p2 : 0; };
        const bool want_part = ctx->debug_force_lsd < 1.0 && ctx->sort_partition >= 1.0;
        uint64_t max_ptiles = 1, max_phist = 1, max_psub = 1; int max_pnb = 1;
        std::vector<char> batch_part(plan.size(), 0);
        for (size_t bi = 0; bi < plan.size() && want_part; bi++) {
            const SortBatch& sb = plan[bi];
            bool ok = sb.nk > 0; uint64_t tiles = 0, hist = 0, subs = 0;
            for (int b = sb.lo; b < sb.hi && ok; b++) {
                const uint64_t c = h_kmer[(size_t)b], ps = part_subs(c), t = c ? (h_rec[(size_t)b] + PTR - 1) / PTR : 0;
                if (ps > (uint64_t)kPartMaxSubs || c >= 0xFFFFFFF0ull || 2 * cfg->k < 11) ok = false;
                tiles += t; hist += t * ps; subs += ps;
            }
            if (!ok || tiles >= 0xFFFFFFFFull || subs >= 0xFFFFFFFFull) continue;
            batch_part[bi] = 1;
            max_ptiles = std::max(max_ptiles, tiles); max_phist = std::max(max_phist, hist); max_psub = std::max(max_psub, subs); max_pnb = std::max(max_pnb, sb.hi - sb.lo);
            max_chunks = std::max<uint64_t>(max_chunks, 2 * sb.nk / local_cap + subs / 2 + (uint64_t)(sb.hi - sb.lo) + 16);
        }
        unsigned int* d_sdt = nullptr; ChunkDesc* d_chunks = nullptr;
        uint32_t *d_p32 = nullptr, *d_phist = nullptr, *d_pbase = nullptr, *d_ptnk = nullptr, *d_pkcur = nullptr, *d_pmidbin = nullptr;
        unsigned long long *d_p64 = nullptr, *d_ptoff = nullptr, *d_pmid = nullptr;
        const size_t sc_smem = (size_t)PartGeom<WIDE>::kBufKeys * (sizeof(Key) + 2);
        int occ_p1 = 1, occ_p3 = 1;
        if (want_part) {
            CKC(cudaFuncSetAttribute(k_place_keys<WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
            CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p1, k_expand_hist<WIDE, true>, PartGeom<WIDE>::kThreads, 0));
            CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p3, k_place_keys<WIDE, true>, PartGeom<WIDE>::kThreads, sc_smem));
            occ_p1 = std::max(occ_p1, 1); occ_p3 = std::max(occ_p3, 1);
            CKC(dmalloc(ctx, &d_p32, ((size_t)max_pnb + 1) * 8 + 16)); CKC(dmalloc(ctx, &d_p64, ((size_t)max_pnb + 1) * 16 + 16));
            CKC(dmalloc(ctx, &d_phist, max_phist * 4)); CKC(dmalloc(ctx, &d_pbase, max_phist * 4));
            CKC(dmalloc(ctx, &d_ptoff, max_ptiles * 8)); CKC(dmalloc(ctx, &d_ptnk, max_ptiles * 4)); CKC(dmalloc(ctx, &d_pkcur, (size_t)max_pnb * 4));
            CKC(dmalloc(ctx, &d_pmid, (max_psub + 1) * 8)); CKC(dmalloc(ctx, &d_pmidbin, max_psub * 4));
        }
        // (the first key array also serves as the tile-by-tile expansion buffer: every tile and bin may be padded by one key)
        CKC(dmalloc(ctx, &d_keysA, (std::max<uint64_t>(max_nk, 1) + max_ptiles + (uint64_t)max_pnb + 8) * sizeof(Key))); CKC(dmalloc(ctx, &d_keysB, std::max<uint64_t>(max_nk, 1) * sizeof(Key)));
        CKC(dmalloc(ctx, &d_tile_seg, std::max<uint64_t>(max_tiles, 1) * 4)); CKC(dmalloc(ctx, &d_tile_hist, std::max<uint64_t>(max_hist, 1) * 4));
        CKC(dmalloc(ctx, &d_tile_heads, (max_tiles + 1) * 4)); CKC(dmalloc(ctx, &d_seg_tile0, ((size_t)B + 1) * 4));
        CKC(dmalloc(ctx, &d_sdt, max_sdt * 4)); CKC(dmalloc(ctx, &d_chunks, max_chunks * sizeof(ChunkDesc)));
        CKC(dmalloc(ctx, &d_first, (max_nk + 1) * 8));          // run heads of one batch (at most one per key)
        std::vector<unsigned long long> kb; std::vector<unsigned int> tseg, st0, sdt0;
        std::vector<uint32_t> p32; std::vector<unsigned long long> p64;
        for (size_t sbi = 0; sbi < plan.size(); sbi++) {
            const SortBatch& sb = plan[sbi];
            const int lo = sb.lo, hi = sb.hi;
            const uint64_t nk = sb.nk, n_tiles = sb.n_tiles;
            const int n_seg = hi - lo;
            if (n_tiles == 0) {          // only empty bins: just carry the output offset forward
                k_bin_offsets<<<1, 256, 0, s>>>(d_distinct, d_out_base, lo, hi, d_small); CKLC();
                continue;
            }
            kb.clear(); kb.push_back(0); tseg.clear(); st0.clear(); st0.push_back(0); sdt0.clear();
            for (int b = lo; b < hi; b++) {
                const uint64_t c = h_kmer[(size_t)b];
                kb.push_back(kb.back() + c); sdt0.push_back((unsigned)c);
                for (uint64_t t = 0; t < (c + kSortTile - 1) / kSortTile; t++) tseg.push_back((unsigned)(b - lo));
                st0.push_back((unsigned)tseg.size());
            }
            CKC(cudaMemcpyAsync(d_tbl_base, kb.data(), kb.size() * 8, cudaMemcpyHostToDevice, s));
            CKC(cudaMemcpyAsync(d_tile_seg, tseg.data(), (size_t)n_tiles * 4, cudaMemcpyHostToDevice, s));
            CKC(cudaMemcpyAsync(d_seg_tile0, st0.data(), st0.size() * 4, cudaMemcpyHostToDevice, s));
            st->h2d_bytes += kb.size() * 8 + n_tiles * 4 + st0.size() * 4;
            CKC(cudaEventRecord(ctx->ev[6], s));
            SortParams Q;
            Q.seg_base = d_tbl_base; Q.tile_seg = d_tile_seg; Q.seg_tile0 = d_seg_tile0; Q.tile_hist = d_tile_hist;
            Q.n_tiles = (unsigned)n_tiles; Q.n_seg = n_seg; Q.seg_digit_tot = nullptr;
            void* in = d_keysA; void* out = d_keysB;
            bool sorted_done = false, expanded = false;
            if (batch_part[sbi]) {
                const int nbb = hi - lo;
                p32.assign(2 * (size_t)(nbb + 1), 0); p64.assign(2 * (size_t)(nbb + 1), 0);
                uint32_t* tf = p32.data(); uint32_t* sf = tf + nbb + 1; unsigned long long* ho = p64.data(); unsigned long long* kbp = ho + nbb + 1;
                uint64_t tiles = 0, subs = 0, hist = 0, keys = 0;
                for (int b = lo; b < hi; b++) {
                    const uint64_t c = h_kmer[(size_t)b], ps = part_subs(c), t = c ? (h_rec[(size_t)b] + PTR - 1) / PTR : 0;
                    tf[b - lo] = (uint32_t)tiles; sf[b - lo] = (uint32_t)subs; ho[b - lo] = hist; kbp[b - lo] = keys;
                    tiles += t; subs += ps; hist += t * ps; keys += c;
                }
                tf[nbb] = (uint32_t)tiles; sf[nbb] = (uint32_t)subs; kbp[nbb] = keys;
                CKC(cudaMemcpyAsync(d_p32, p32.data(), p32.size() * 4, cudaMemcpyHostToDevice, s));
                CKC(cudaMemcpyAsync(d_p64, p64.data(), p64.size() * 8, cudaMemcpyHostToDevice, s));
                CKC(cudaMemsetAsync(d_pkcur, 0, (size_t)nbb * 4, s));
                CKC(cudaMemsetAsync(d_ovf, 0, 8, s));
                st->h2d_bytes += p32.size() * 4 + p64.size() * 8;
                PartParams PP;
                PP.records = d_records; PP.bin_rec_base = d_bin_base; PP.bin_rec_cnt = nullptr; PP.bin_lo = lo; PP.bin_hi = hi;
                PP.tile_first = d_p32; PP.sub_first = d_p32 + nbb + 1; PP.hist_off = d_p64; PP.key_base = d_p64 + nbb + 1;
                PP.n_tiles = (uint32_t)tiles; PP.n_sub = (uint32_t)subs; PP.tile_hist = d_phist; PP.tile_base = d_pbase;
                PP.bin_key_cursor = d_pkcur; PP.tile_key_off = d_ptoff; PP.tile_nkeys = d_ptnk; PP.keys_lin = d_keysA; PP.keys = d_keysB;
                PP.mid_key_base = d_pmid; PP.mid_bin = d_pmidbin; PP.k = cfg->k;
                k_expand_hist<WIDE, true><<<(unsigned)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * occ_p1), PartGeom<WIDE>::kThreads, 0, s>>>(PP); CKLC();
                k_sub_scan<<<(unsigned)nbb, 256, 0, s>>>(PP); CKLC();
                k_place_keys<WIDE, true><<<(unsigned)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * occ_p3), PartGeom<WIDE>::kThreads, sc_smem, s>>>(PP); CKLC();
                const uint64_t cap_chunks = 2 * nk / local_cap + subs / 2 + (uint64_t)nbb + 16;
                k_form_chunks_sub<<<(nbb + 127) / 128, 128, 0, s>>>(d_pmid, PP.sub_first, nbb, (unsigned)local_cap, d_chunks, (unsigned)cap_chunks,
                                                                    (unsigned int*)(d_ovf + 1), d_ovf); CKLC();
                int flags[2] = {0, 0};
                CKC(cudaMemcpyAsync(flags, d_ovf, 8, cudaMemcpyDeviceToHost, s));
                CKC(cudaStreamSynchronize(s));
                st->d2h_bytes += 8;
                expanded = true;
                if (!flags[0]) {
                    const unsigned n_chunks = (unsigned)flags[1];
                    const unsigned grid = std::min<unsigned>(n_chunks, (unsigned)ctx->n_sm * 4);
                    if (n_chunks) { k_radix_local<WIDE><<<grid, 512, local_smem, s>>>(d_keysB, d_keysA, d_chunks, n_chunks, n_pass); CKLC(); }
                    in = d_keysA;
                    sorted_done = true;
                } else {
                    st->n_fallbacks++;           // a sub-bucket larger than a chunk: LSD passes over the k-mers as k_place_keys left them (bin-major)
                    in = d_keysB; out = d_keysA;
                }
            }
            if (!expanded) {
                ExpandParams E;
                E.records = d_records; E.rec_lo = h_base[(size_t)lo]; E.rec_hi = h_base[(size_t)hi];
                E.bin_base = d_bin_base; E.bin_lo = lo; E.bin_hi = hi; E.key_base = d_tbl_base; E.key_cursor = d_cursor; E.keys = d_keysA; E.k = cfg->k;
                CKC(cudaMemsetAsync(d_cursor + lo, 0, (size_t)(hi - lo) * 8, s));
                const uint64_t nr = E.rec_hi - E.rec_lo;
                k_expand<WIDE><<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(E); CKLC();
            }
            if (!expanded && sb.bits >= 0) {
                // MSD partition on the top bits (skipped when every bin already fits one chunk) ...
                int nd = 1;
                if (sb.bits > 0) {
                    nd = 1 << sb.bits;
                    Q.in = in; Q.out = out; Q.shift = 2 * cfg->k - sb.bits; Q.nd = nd; Q.seg_digit_tot = d_sdt;
                    k_radix_hist<WIDE><<<(unsigned)n_tiles, 256, 0, s>>>(Q); CKLC();
                    k_radix_scan<<<n_seg, 256, 0, s>>>(Q); CKLC();
                    k_radix_scatter<WIDE><<<(unsigned)n_tiles, 256, 0, s>>>(Q); CKLC();
                    Q.seg_digit_tot = nullptr;
                } else {
                    CKC(cudaMemcpyAsync(d_sdt, sdt0.data(), sdt0.size() * 4, cudaMemcpyHostToDevice, s));
                }
                // ... then chunks of <= local_cap keys, each sorted completely in shared memory
                const uint64_t cap_chunks = 2 * nk / local_cap + (uint64_t)n_seg + 16;
                CKC(cudaMemsetAsync(d_ovf, 0, 8, s));
                k_form_chunks<<<(n_seg + 127) / 128, 128, 0, s>>>(d_tbl_base, d_sdt, n_seg, nd, (unsigned)local_cap, d_chunks,
                                                                  (unsigned)cap_chunks, (unsigned int*)(d_ovf + 1), d_ovf); CKLC();
                int flags[2] = {0, 0};
                CKC(cudaMemcpyAsync(flags, d_ovf, 8, cudaMemcpyDeviceToHost, s));
                CKC(cudaStreamSynchronize(s));
                st->d2h_bytes += 8;
                if (!flags[0]) {
                    const void* src = sb.bits > 0 ? out : in;
                    void* dst = sb.bits > 0 ? in : out;
                    const unsigned n_chunks = (unsigned)flags[1];
                    const unsigned grid = std::min<unsigned>(n_chunks, (unsigned)ctx->n_sm * 4);
                    if (n_chunks) { k_radix_local<WIDE><<<grid, 512, local_smem, s>>>(src, dst, d_chunks, n_chunks, n_pass); CKLC(); }
                    in = dst;
                    sorted_done = true;
                } else {
                    st->n_fallbacks++;           // a sub-bucket larger than a chunk: LSD passes over the (untouched) expanded keys
                }
            }
            if (!sorted_done) {
                if (!expanded) { in = d_keysA; out = d_keysB; }
                for (int p = 0; p < n_pass; p++) {
                    Q.in = in; Q.out = out; Q.shift = 8 * p; Q.nd = 256;
                    k_radix_hist<WIDE><<<(unsigned)n_tiles, 256, 0, s>>>(Q); CKLC();
                    k_radix_scan<<<n_seg, 256, 0, s>>>(Q); CKLC();
                    k_radix_scatter<WIDE><<<(unsigned)n_tiles, 256, 0, s>>>(Q); CKLC();
                    std::swap(in, out);
                }
            }
            CKC(cudaEventRecord(ctx->ev[7], s));
            RleParams R;
            R.keys = in; R.seg_base = d_tbl_base; R.tile_seg = d_tile_seg; R.seg_tile0 = d_seg_tile0; R.n_tiles = (unsigned)n_tiles;
            R.bin_lo = lo; R.tile_heads = d_tile_heads; R.bin_distinct = d_distinct; R.out_off = 0; R.n_keys = nk;
            R.out_keys = nullptr; R.out_cnt = nullptr; R.first_idx = nullptr;
            k_rle<WIDE, 0><<<(unsigned)n_tiles, 256, 0, s>>>(R); CKLC();
            k_scan_u32<<<1, 1024, 0, s>>>(d_tile_heads, (unsigned)n_tiles); CKLC();
            k_bin_offsets<<<1, 256, 0, s>>>(d_distinct, d_out_base, lo, hi, d_small); CKLC();
            unsigned long long batch_total = 0;
            CKC(cudaMemcpyAsync(&batch_total, d_small, 8, cudaMemcpyDeviceToHost, s));
            CKC(cudaStreamSynchronize(s));
            st->d2h_bytes += 8;
            Chunk ch; ch.n = batch_total;
            CKC(dmalloc(ctx, &ch.keys, (size_t)batch_total * sizeof(Key)));
            cudaError_t e2 = dmalloc(ctx, &ch.cnt, (size_t)batch_total * 4);
            if (e2 != cudaSuccess) { dfree(ctx, ch.keys); CKC(e2); }
            res->chunks.push_back(ch);
            R.out_keys = ch.keys; R.out_cnt = ch.cnt; R.first_idx = d_first;
            k_rle<WIDE, 1><<<(unsigned)n_tiles, 256, 0, s>>>(R); CKLC();
            k_rle_counts<<<(unsigned)((batch_total + 255) / 256), 256, 0, s>>>(d_first, batch_total, nk, ch.cnt, 0); CKLC();
            CKC(cudaEventRecord(ctx->ev[8], s));
            CKC(cudaStreamSynchronize(s));
            { float ms; cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]); ms_count += ms; cudaEventElapsedTime(&ms, ctx->ev[7], ctx->ev[8]); ms_compact += ms; }
            out_total += batch_total;
            st->n_batches++;
        }
    }
    CKC(cudaEventRecord(ctx->ev[3], s));

    // ---- stage 5: digest + bookkeeping
    if (!part_done) CKC(cudaMemcpyAsync(res->out_base.data(), d_out_base, bB + 8, cudaMemcpyDeviceToHost, s));
    if (!cfg->use_ht && !part_done) {           // the HT path folds its digest into k_compact_ht, the partitioned paths into their count kernels
        uint64_t origin = 0;
        for (const Chunk& ch : res->chunks) {
            if (!ch.n) continue;
            DigestParams D;
            D.keys = ch.keys; D.cnt = ch.cnt; D.out_base = d_out_base; D.B = B; D.n = ch.n; D.origin = origin; D.acc = d_acc; D.bin_shift = ctx->job_split;
            int grid = (int)std::min<uint64_t>((ch.n + 255) / 256, (uint64_t)ctx->n_sm * 8);
            k_digest<WIDE><<<grid, 256, 0, s>>>(D); CKLC();
            origin += ch.n;
        }
    }
    unsigned long long h_acc[192];
    CKC(cudaMemcpyAsync(h_acc, d_acc, 192 * 8, cudaMemcpyDeviceToHost, s));
    CKC(cudaEventRecord(ctx->ev[4], s));
    CKC(cudaStreamSynchronize(s));
    st->d2h_bytes += bB + 8 + 192 * 8;
    unsigned long long acc[3] = {0, 0, 0};
    for (int i = 0; i < 64; i++) { acc[0] += h_acc[i]; acc[1] ^= h_acc[64 + i]; acc[2] += h_acc[128 + i]; }
    tr.mark("digest synced");
    res->total = out_total;
    // empty trailing bins keep the running offset
    for (int b = 0; b < B; b++) if (res->out_base[(size_t)b + 1] < res->out_base[(size_t)b]) res->out_base[(size_t)b + 1] = res->out_base[(size_t)b];
    st->n_distinct = out_total; st->digest_sum = acc[0]; st->digest_xor = acc[1]; st->total_count = acc[2];
    float ms;
    if (!pre) { cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); st->ms_stage[2] = ms; }
    st->ms_stage[3] = ms_count; st->ms_stage[4] = ms_compact;
    cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]); st->ms_stage[5] = ms;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[4]); st->ms_stage[7] = ms;      // whole device pipeline
    cleanup();
    if (st->total_count != st->n_kmers) {
        // counts are 32-bit (the reference counts in Int, SBKC:562,676): every wrap takes exactly 2^32 off the sum of counts
        if (st->total_count < st->n_kmers && ((st->n_kmers - st->total_count) & 0xFFFFFFFFull) == 0)
            return fkm_set_error(FKM_EOVERFLOW, "a k-mer occurs more than 2^32-1 times: 32-bit counts overflow (the reference counts in Int, SBKC:562,676)");
        return fkm_set_error(FKM_ECUDA, "internal check failed: sum of counts %llu != valid k-windows %llu",
                             (unsigned long long)st->total_count, (unsigned long long)st->n_kmers);
    }
    return FKM_OK;
#undef CKC
#undef CKLC
}

// ------------------------------------------------------------------ the shared-memory count pipeline (fkm_smem.cuh)

// geometry of k_count_smem's shared memory for one key width
struct SmemGeom { uint32_t cap_slots, max_fill, stage_recs; size_t bytes; };
template <bool WIDE>
static SmemGeom smem_geometry(fkm_ctx* ctx) {
    typedef typename Traits<WIDE>::Key Key;
    const size_t rec_bytes = Traits<WIDE>::kRecWords * 8;
    SmemGeom g;
    g.stage_recs = (uint32_t)(16384 / rec_bytes);                        // kSmStages staging buffers of 16 KB
    const size_t avail = ctx->smem_optin - kSmStages * 16384 - 1024;     // 1 KB for the kernel's static shared memory
    size_t cap = avail / (sizeof(Key) + 4);                              // per slot: key + count
    if (ctx->smem_table_slots >= 64.0) cap = std::min<size_t>(cap, (size_t)ctx->smem_table_slots);
    cap = cap / 32 * 32;
    g.cap_slots = (uint32_t)cap;
    g.max_fill = (uint32_t)std::min<size_t>(cap * 3 / 4, cap > 2 * (size_t)kSmThreads + 64 ? cap - 2 * (size_t)kSmThreads : cap / 4);
    g.bytes = cap * (sizeof(Key) + 4) + (size_t)kSmStages * 16384;
    return g;
}

// Hash path with the tables in shared memory: dual scan -> (bin, cell) histogram -> mid bins -> records mid-bin-major ->
// k_count_smem.  Two phases: the first B/64 bins are a sample sized for all-distinct k-mers; the distinct / k-mer ratio
// they show sizes the mid bins and the output arrays of the rest.
template <bool WIDE>
static int run_pipeline_smem(fkm_ctx* ctx, const fkm_config* cfg, int32_t B, fkm_result* res, fkm_stats* st, ScanState& scan) {
    typedef typename Traits<WIDE>::Key Key;
    cudaStream_t s = ctx->stream;
    const int rec_bytes = Traits<WIDE>::kRecWords * 8;
    res->ctx = ctx; res->gen = ctx->gen;
    res->B = B; res->k = cfg->k; res->wide = WIDE; res->sorted = false; res->device = ctx->device;
    res->out_base.assign((size_t)B + 1, 0);
    Trace tr;
    const size_t bB = (size_t)B * 8;
    for (const ChunkScan& C : scan.chunks) if (C.ev_ovf) return kRetryGlobal;          // the event lists were too small
    const std::vector<unsigned long long>& h_rec = scan.h_rec; const std::vector<unsigned long long>& h_kmer = scan.h_kmer;
    uint64_t n_rec = 0, n_kmers = 0, nonempty = 0;
    for (int b = 0; b < B; b++) { n_rec += h_rec[(size_t)b]; n_kmers += h_kmer[(size_t)b]; nonempty += h_rec[(size_t)b] ? 1 : 0; }
    st->n_kmers = n_kmers; st->n_superkmers = n_rec; st->superkmer_bytes = n_rec * rec_bytes; st->n_nonempty_bins = nonempty;

    const SmemGeom G = smem_geometry<WIDE>(ctx);
    CK(cudaFuncSetAttribute(k_count_smem<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.bytes));
    const int cb = scan.cell_bits; const size_t n_cells = (size_t)B << cb;
    const uint64_t C_cells = 1ull << cb;
    const int s_hi = (int)std::min<uint32_t>(scan.sample_hi, (uint32_t)B);

    unsigned long long *d_acc = nullptr, *d_counters = nullptr, *d_mid_first = nullptr;
    uint32_t* d_cell2mid = nullptr; int* d_flags = nullptr; void* d_slow_keys = nullptr; uint32_t* d_slow_cnt = nullptr;
    CK(dmalloc(ctx, &d_acc, 192 * 8)); CK(dmalloc(ctx, &d_counters, 64)); CK(dmalloc(ctx, &d_flags, 16));
    CK(dmalloc(ctx, &d_mid_first, bB + 8)); CK(dmalloc(ctx, &d_cell2mid, n_cells * 4));
    const uint64_t slow_slots = std::max<uint64_t>(1024, (uint64_t)ctx->smem_slow_slots);
    CK(dmalloc(ctx, &d_slow_keys, (size_t)ctx->n_sm * slow_slots * sizeof(Key))); CK(dmalloc(ctx, &d_slow_cnt, (size_t)ctx->n_sm * slow_slots * 4));
    CK(cudaMemsetAsync(d_acc, 0, 192 * 8, s));
    CK(cudaMemsetAsync(d_counters, 0, 64, s)); CK(cudaMemsetAsync(d_flags, 0, 16, s));
    float ms_pack = 0, ms_scatter = 0, ms_count = 0;
    uint64_t out_total = 0, n_mid_total = 0;
    std::vector<unsigned long long> h_mid_first((size_t)B + 1);
    // one phase: bins [lo, hi), events of list `li`, mid bins of about T k-mers, output arrays of out_cap entries
    const unsigned grid = (unsigned)ctx->n_sm;
    std::vector<unsigned long long> h_cta_total((size_t)grid), h_bin_off((size_t)B);
    std::vector<uint32_t> h_bin_cta((size_t)B);
    unsigned long long* d_cta_total = nullptr; uint32_t* d_bin_cta = nullptr; unsigned long long* d_bin_off = nullptr;
    CK(dmalloc(ctx, &d_cta_total, (size_t)grid * 8)); CK(dmalloc(ctx, &d_bin_cta, (size_t)B * 4)); CK(dmalloc(ctx, &d_bin_off, bB));
    // one phase: bins [lo, hi), events of list `li`, mid bins of about T k-mers; rho = distinct / k-mers expected (sizes the output)
    auto run_phase = [&](int lo, int hi, int li, uint64_t T, double rho) -> int {
        if (lo >= hi) return FKM_OK;
        uint64_t n_mid = 0, rec_phase = 0, km_phase = 0;
        for (int b = 0; b <= B; b++) h_mid_first[(size_t)b] = 0;
        for (int b = lo; b < hi; b++) {
            h_mid_first[(size_t)b] = n_mid;
            const uint64_t km = h_kmer[(size_t)b];
            const uint64_t Tb = std::max<uint64_t>(T, (km + C_cells - 1) / C_cells);     // never more mid bins than cells
            n_mid += km ? (km - 1) / Tb + 1 : 0;
            rec_phase += h_rec[(size_t)b]; km_phase += km;
        }
        for (int b = hi; b <= B; b++) h_mid_first[(size_t)b] = n_mid;
        for (int b = lo; b < hi; b++) res->out_base[(size_t)b] = out_total;
        if (n_mid >= 0xFFFFFFFFull) return kRetryGlobal;
        if (n_mid == 0) return FKM_OK;
        // every CTA counts an equal share of the k-mers (plus at most one mid bin) into its own output region
        const uint64_t share = km_phase / grid + 8 * T + 1;
        const uint64_t region_cap = std::min<uint64_t>(share, (uint64_t)((double)share * std::min(1.0, rho * 1.25 + 0.01))) + 65536;
        unsigned long long *d_mid_rec = nullptr, *d_mid_kmer = nullptr, *d_mid_base = nullptr, *d_mid_kbase = nullptr;
        uint32_t* d_mid_bin = nullptr; void* d_records = nullptr; void* d_okeys = nullptr; uint32_t* d_ocnt = nullptr;
        CK(dmalloc(ctx, &d_mid_rec, n_mid * 8)); CK(dmalloc(ctx, &d_mid_kmer, n_mid * 8)); CK(dmalloc(ctx, &d_mid_base, (n_mid + 1) * 8)); CK(dmalloc(ctx, &d_mid_kbase, (n_mid + 1) * 8));
        CK(dmalloc(ctx, &d_mid_bin, n_mid * 4));
        CK(dmalloc(ctx, &d_records, std::max<size_t>(16, (size_t)rec_phase * rec_bytes)));
        CK(dmalloc(ctx, &d_okeys, (size_t)grid * region_cap * sizeof(Key))); CK(dmalloc(ctx, &d_ocnt, (size_t)grid * region_cap * 4));
        CK(cudaMemsetAsync(d_mid_rec, 0, n_mid * 8, s)); CK(cudaMemsetAsync(d_mid_kmer, 0, n_mid * 8, s));
        CK(cudaMemsetAsync(d_cta_total, 0, (size_t)grid * 8, s));
        CK(cudaMemcpyAsync(d_mid_first, h_mid_first.data(), bB + 8, cudaMemcpyHostToDevice, s));
        st->h2d_bytes += bB + 8;
        CK(cudaEventRecord(ctx->ev[5], s));
        CellsParams CP;
        CP.cell_rec = scan.d_cell_rec; CP.cell_kmer = scan.d_cell_kmer; CP.cell_bits = cb; CP.bin_lo = lo; CP.bin_hi = hi; CP.T = T;
        CP.bin_kmer = scan.d_hist_kmer; CP.mid_first = d_mid_first; CP.cell2mid = d_cell2mid; CP.mid_rec = d_mid_rec; CP.mid_kmer = d_mid_kmer; CP.mid_bin = d_mid_bin;
        k_cells_assign<<<(unsigned)(((uint64_t)(hi - lo) * 32 + 255) / 256), 256, 0, s>>>(CP); CKL();
        k_excl_scan_u64<<<1, 1024, 0, s>>>(d_mid_rec, d_mid_base, n_mid); CKL();
        k_excl_scan_u64<<<1, 1024, 0, s>>>(d_mid_kmer, d_mid_kbase, n_mid); CKL();
        const unsigned long long cell_lo = (unsigned long long)lo << cb, cell_hi = (unsigned long long)hi << cb;
        unsigned long long* d_mid_next = nullptr;                                     // write cursors of the scatter: start at the mid bins' offsets
        CK(dmalloc(ctx, &d_mid_next, n_mid * 8));
        CK(cudaMemcpyAsync(d_mid_next, d_mid_base, n_mid * 8, cudaMemcpyDeviceToDevice, s));
        CK(cudaEventRecord(ctx->ev[6], s));
        for (const ChunkScan& C : scan.chunks) {
            if (!C.n_events2[li]) continue;
            Scatter2Params Q;
            Q.events = C.d_events2[li]; Q.n_events = C.n_events2[li]; Q.bases = (const uint64_t*)C.d_bases; Q.n_words = (C.n_pos + 31) / 32;
            Q.cap = WIDE ? (125 - cfg->k) : (61 - cfg->k); Q.k = cfg->k; Q.cell_bits = cb;
            Q.cell_lo = cell_lo; Q.cell_hi = cell_hi; Q.cell2mid = d_cell2mid; Q.mid_next = d_mid_next; Q.records = d_records;
            k_scatter2<WIDE><<<(unsigned)((Q.n_events + 255) / 256), 256, 0, s>>>(Q); CKL();
        }
        CK(cudaEventRecord(ctx->ev[7], s));
        SmemCountParams P;
        P.records = d_records; P.mid_rec_base = d_mid_base; P.mid_kmer_base = d_mid_kbase; P.mid_bin = d_mid_bin; P.mid_first = d_mid_first; P.n_mid = (uint32_t)n_mid;
        P.out_keys = d_okeys; P.out_cnt = d_ocnt; P.region_cap = region_cap; P.cta_total = d_cta_total; P.bin_cta = d_bin_cta; P.bin_off = d_bin_off;
        P.acc = d_acc; P.k = cfg->k; P.cap_slots = G.cap_slots; P.max_fill = G.max_fill; P.stage_recs = G.stage_recs;
        P.slow_keys = d_slow_keys; P.slow_cnt = d_slow_cnt; P.slow_slots = slow_slots; P.slow_max_fill = slow_slots * 7 / 10; P.flags = d_flags; P.counters = d_counters;
        k_count_smem<WIDE><<<grid, kSmBlock, G.bytes, s>>>(P); CKL();
        CK(cudaEventRecord(ctx->ev[8], s));
        int flags[4] = {0, 0, 0, 0};
        CK(cudaMemcpyAsync(h_cta_total.data(), d_cta_total, (size_t)grid * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_bin_cta.data() + lo, d_bin_cta + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_bin_off.data() + lo, d_bin_off + lo, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(flags, d_flags, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        st->d2h_bytes += 16 + (size_t)grid * 8 + (size_t)(hi - lo) * 12;
        { float ms; cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]); ms_pack += ms; cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]); ms_scatter += ms;
          cudaEventElapsedTime(&ms, ctx->ev[7], ctx->ev[8]); ms_count += ms; }
        if (flags[2]) return fkm_set_error(FKM_EOVERFLOW, "a k-mer occurs more than 2^32-1 times: 32-bit counts overflow (the reference counts in Int, SBKC:562,676)");
        if (flags[0] || flags[1]) return kRetryGlobal;
        // one result chunk per CTA region; a bin's entries begin in the region of the CTA that counted its first mid bin
        std::vector<unsigned long long> cta_start((size_t)grid + 1);
        cta_start[0] = out_total;
        for (unsigned c = 0; c < grid; c++) {
            cta_start[(size_t)c + 1] = cta_start[(size_t)c] + h_cta_total[(size_t)c];
            if (h_cta_total[(size_t)c]) {
                Chunk ch; ch.keys = (char*)d_okeys + (size_t)c * region_cap * sizeof(Key); ch.cnt = d_ocnt + (size_t)c * region_cap; ch.n = h_cta_total[(size_t)c];
                res->chunks.push_back(ch);
            }
        }
        unsigned long long next = cta_start[(size_t)grid];
        for (int b = hi - 1; b >= lo; b--) {
            if (h_mid_first[(size_t)b + 1] > h_mid_first[(size_t)b]) next = cta_start[(size_t)h_bin_cta[(size_t)b]] + h_bin_off[(size_t)b];
            res->out_base[(size_t)b] = next;
        }
        out_total = cta_start[(size_t)grid];
        n_mid_total += n_mid;
        st->n_batches++;
        return FKM_OK;
    };

    CK(cudaEventRecord(ctx->ev[1], s));
    // phase A: the sample, sized for all-distinct k-mers
    uint64_t km_a = 0; for (int b = 0; b < s_hi; b++) km_a += h_kmer[(size_t)b];
    const uint64_t T_safe = std::max<uint64_t>(32, (uint64_t)((double)G.max_fill * ctx->smem_fill));
    int rc = run_phase(0, s_hi, 0, T_safe, 1.0); if (rc) return rc;
    tr.mark("sample phase done", (long long)out_total);
    if (s_hi < B) {
        double rho = 1.0;
        if (km_a > 100000) rho = std::min(1.0, (double)out_total / (double)km_a);
        rho *= ctx->debug_rho_scale;
        const double rho_plan = std::min(1.0, rho * 1.1 + 0.01);
        const uint64_t T_b = std::max<uint64_t>(T_safe, (uint64_t)((double)G.max_fill * ctx->smem_fill / rho_plan));
        rc = run_phase(s_hi, B, 1, T_b, rho); if (rc) return rc;
    }
    res->out_base[(size_t)B] = out_total;
    CK(cudaEventRecord(ctx->ev[3], s));
    unsigned long long h_acc[192], h_counters[8];
    CK(cudaMemcpyAsync(h_acc, d_acc, 192 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_counters, d_counters, 64, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[4], s));
    CK(cudaStreamSynchronize(s));
    st->d2h_bytes += 192 * 8 + 64;
    unsigned long long acc[3] = {0, 0, 0};
    for (int i = 0; i < 64; i++) { acc[0] += h_acc[i]; acc[1] ^= h_acc[64 + i]; acc[2] += h_acc[128 + i]; }
    res->total = out_total;
    st->n_distinct = out_total; st->digest_sum = acc[0]; st->digest_xor = acc[1]; st->total_count = acc[2];
    st->n_mid_bins = n_mid_total; st->n_slow_bins = h_counters[0];
    st->ms_stage[2] = ms_scatter; st->ms_stage[3] = ms_count; st->ms_stage[4] = ms_pack;
    float ms; cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]); st->ms_stage[5] = ms;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[4]); st->ms_stage[7] = ms;
    if (st->total_count != st->n_kmers)
        return fkm_set_error(FKM_ECUDA, "internal check failed (shared-memory path): sum of counts %llu != valid k-windows %llu",
                             (unsigned long long)st->total_count, (unsigned long long)st->n_kmers);
    return FKM_OK;
}

static bool want_smem_path(const fkm_ctx* ctx, const fkm_config* cfg) { return cfg->use_ht && ctx->count_mode >= 1.0 && ctx->count_mode < 2.0; }

static int count_device(fkm_ctx* ctx, const fkm_config* cfg, const void* d_bases, const void* d_inv, uint64_t n_pos,
                        fkm_result** out, fkm_stats* stats, const PreScattered* pre = nullptr, ScanState* scanned = nullptr) {
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    if (!ctx) return fkm_set_error(FKM_EINVAL, "ctx is NULL");
    CK(cudaSetDevice(ctx->device));
    // internal bins (fkm_common.h split_bin): chosen here unless the job's scan has already run with its choice
    const int32_t B_user = B;
    if (!pre && !scanned) ctx->job_split = choose_split(ctx, cfg, B, n_pos);
    B = B_user << ctx->job_split;
    fkm_stats local; fkm_stats* st = stats ? stats : &local;
    const uint64_t keep_h2d = st->h2d_bytes, keep_bases = st->n_bases; const double keep_ms0 = st->ms_stage[0];
    memset(st, 0, sizeof *st);
    st->h2d_bytes = keep_h2d; st->n_bases = keep_bases; st->ms_stage[0] = keep_ms0;
    st->n_positions = n_pos;
    auto t0 = std::chrono::steady_clock::now();
    fkm_result* res = new fkm_result();
    const bool wide = cfg->k > 32;
    rc = -1000;
    ScanState rescan;                                      // non-dual scan of the same chunks, for the fallback
    if (!pre && want_smem_path(ctx, cfg) && (!scanned || scanned->dual)) {
        // ---- tables in shared memory
        ScanState local_scan;
        ScanState* S = scanned;
        const Arena::Mark mk = ctx->arena.mark();
        if (!S) {
            S = &local_scan;
            rc = stage_scan(ctx, cfg, B, d_bases, d_inv, n_pos, S, st, true);
        } else { CK(cudaEventRecord(ctx->ev[0], ctx->stream)); rc = FKM_OK; }
        if (!rc) rc = wide ? run_pipeline_smem<true>(ctx, cfg, B, res, st, *S) : run_pipeline_smem<false>(ctx, cfg, B, res, st, *S);
        if (rc == kRetryGlobal) {
            // a table overflowed even on the slow path, or the output estimate was too small: the global-table pipeline redoes the job
            const uint64_t fb = st->n_fallbacks + 1, h2d = st->h2d_bytes, d2h = st->d2h_bytes;
            delete res; res = new fkm_result();
            if (!scanned) ctx->arena.release(mk);
            else {
                rc = scan_begin(ctx, cfg, B, &rescan);
                for (size_t c = 0; !rc && c < scanned->chunks.size(); c++) rc = scan_chunk(ctx, &rescan, scanned->chunks[c].d_bases, scanned->chunks[c].d_inv, scanned->chunks[c].n_pos);
                if (!rc) { fkm_stats tmp; memset(&tmp, 0, sizeof tmp); rc = scan_end(ctx, &rescan, &tmp); }
                if (rc) { fkm_result_free(res); return rc; }
                scanned = &rescan;
            }
            st->n_fallbacks = fb; st->h2d_bytes = h2d; st->d2h_bytes = d2h; st->n_batches = 0;
            rc = -1000;
        }
    }
    if (rc == -1000) {
        const Arena::Mark mk2 = ctx->arena.mark();
        rc = wide ? run_pipeline<true>(ctx, cfg, B, d_bases, d_inv, n_pos, res, st, pre, scanned)
                  : run_pipeline<false>(ctx, cfg, B, d_bases, d_inv, n_pos, res, st, pre, scanned);
        if (rc == kBinTooBig) {
            ctx->arena.release(mk2);
            fkm_config c2 = *cfg; c2.use_ht = 1;
            delete res; res = new fkm_result();
            rc = wide ? run_pipeline<true>(ctx, &c2, B, d_bases, d_inv, n_pos, res, st, pre, scanned)
                      : run_pipeline<false>(ctx, &c2, B, d_bases, d_inv, n_pos, res, st, pre, scanned);
            if (!rc) rc = fkm_set_error(FKM_EINVAL, "a bin holds 2^32 k-mers or more: the sort path (useHT=0) indexes a bin with 32 bits, use more bins or useHT=1");
        }
    }
    st->gpu_launches = ctx->job_launches;       // everything since job_begin (ingest included)
    st->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc) { fkm_result_free(res); return rc; }
    if (ctx->job_split) {                       // the caller sees the configuration's bins: internal bin b << split is where bin b begins
        std::vector<uint64_t> ob((size_t)B_user + 1);
        for (int32_t b = 0; b <= B_user; b++) ob[(size_t)b] = res->out_base[(size_t)b << ctx->job_split];
        res->out_base.swap(ob); res->B = B_user;
        uint64_t nonempty = 0;
        for (int32_t b = 0; b < B_user; b++) nonempty += res->out_base[(size_t)b + 1] > res->out_base[(size_t)b] ? 1 : 0;
        st->n_nonempty_bins = nonempty;
    }
    if (out) *out = res; else fkm_result_free(res);
    return FKM_OK;
}

extern "C" int fkm_count_packed_device(fkm_ctx* ctx, const fkm_config* cfg, const void* d_bases, const void* d_inv,
                                       uint64_t n_pos, fkm_result** out, fkm_stats* stats) {
    if (!ctx) return fkm_set_error(FKM_EINVAL, "ctx is NULL");
    if (stats) { stats->h2d_bytes = 0; stats->n_bases = 0; stats->ms_stage[0] = 0; }
    int rc = job_begin(ctx); if (rc) return rc;
    return count_device(ctx, cfg, d_bases, d_inv, n_pos, out, stats);
}

extern "C" int fkm_count_packed_host(fkm_ctx* ctx, const fkm_config* cfg, const uint64_t* bases, const uint32_t* inv,
                                     uint64_t n_pos, fkm_result** out, fkm_stats* stats) {
    if (!ctx) return fkm_set_error(FKM_EINVAL, "ctx is NULL");
    { int rc0 = job_begin(ctx); if (rc0) return rc0; }
    const uint64_t nw = (n_pos + 31) / 32;
    void *d_b = nullptr, *d_i = nullptr;
    CK(dmalloc(ctx, &d_b, std::max<size_t>(8, nw * 8)));
    cudaError_t e = dmalloc(ctx, &d_i, std::max<size_t>(4, nw * 4));
    if (e != cudaSuccess) { dfree(ctx, d_b); CK(e); }
    auto t0 = std::chrono::steady_clock::now();
    cudaMemcpyAsync(d_b, bases, nw * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_i, inv, nw * 4, cudaMemcpyHostToDevice, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { dfree(ctx, d_b); dfree(ctx, d_i); CK(e); }
    fkm_stats local; fkm_stats* st = stats ? stats : &local;
    st->h2d_bytes = nw * 12; st->n_bases = 0;
    st->ms_stage[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    int rc = count_device(ctx, cfg, d_b, d_i, n_pos, out, st);
    dfree(ctx, d_b); dfree(ctx, d_i);
    return rc;
}

// FASTA text already in device memory -> packed layout in device memory (fkm_ingest.cuh).
// d_bases / d_inv are preallocated with ingest_cap_words(n) words.
static inline uint64_t ingest_cap_words(uint64_t n_bytes) { return (n_bytes + 1 + 31) / 32 + 1; }   // every text byte keeps at most one position
static int ingest_device(fkm_ctx* ctx, const uint8_t* d_text, uint64_t n, void* d_bases, void* d_inv, uint64_t* n_pos, uint64_t* n_bases) {
    cudaStream_t s = ctx->stream;
    IngestParams P; memset(&P, 0, sizeof P);
    P.text = d_text; P.n = n; P.n_tiles = (n + kIngTile - 1) / kIngTile;
    const uint64_t cap_words = ingest_cap_words(n);
    unsigned long long* d_small = nullptr;
    CK(dmalloc(ctx, (void**)&P.tile_nl, (P.n_tiles + 1) * 8)); CK(dmalloc(ctx, (void**)&P.tile_pos, (P.n_tiles + 1) * 8));
    CK(dmalloc(ctx, (void**)&d_small, 64));
    P.bases = (unsigned long long*)d_bases; P.inv = (unsigned int*)d_inv;
    P.first_hdr = d_small; P.n_hdr = d_small + 1;
    unsigned long long init[4] = {n, 0, 0, 0};
    CK(cudaMemcpyAsync(d_small, init, 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(d_bases, 0, cap_words * 8, s)); CK(cudaMemsetAsync(d_inv, 0, cap_words * 4, s));
    CK(cudaMemsetAsync(P.tile_pos, 0, (P.n_tiles + 1) * 8, s));
    if (P.n_tiles) {
        k_ing_lines<<<(unsigned)P.n_tiles, kIngThreads, 0, s>>>(P); CKL();
        k_scan1<1><<<1, 1024, 0, s>>>(P.tile_nl, P.n_tiles); CKL();
        k_ing_emit<0><<<(unsigned)P.n_tiles, kIngThreads, 0, s>>>(P); CKL();
        k_scan1<0><<<1, 1024, 0, s>>>((long long*)P.tile_pos, P.n_tiles); CKL();
        k_ing_emit<1><<<(unsigned)P.n_tiles, kIngThreads, 0, s>>>(P); CKL();
    }
    k_ing_finish<<<1, 32, 0, s>>>(P, d_small + 2, d_small + 3); CKL();
    unsigned long long out[2] = {0, 0};
    CK(cudaMemcpyAsync(out, d_small + 2, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *n_pos = out[0]; *n_bases = out[1];
    return FKM_OK;
}

// H2D of the raw text + device ingest.  The packed arrays stay in the job's arena; the
// text and the ingest temporaries are released again before counting starts.
static int upload_and_ingest(fkm_ctx* ctx, const uint8_t* fasta, uint64_t n_bytes, void** d_bases, void** d_inv,
                             uint64_t* n_pos, uint64_t* n_bases) {
    const uint64_t cap_words = ingest_cap_words(n_bytes);
    CK(dmalloc(ctx, d_bases, cap_words * 8)); CK(dmalloc(ctx, d_inv, cap_words * 4));
    const Arena::Mark mk = ctx->arena.mark();
    uint8_t* d_text = nullptr;
    CK(dmalloc(ctx, (void**)&d_text, std::max<uint64_t>(n_bytes, 16)));
    CK(cudaMemcpyAsync(d_text, fasta, n_bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = ingest_device(ctx, d_text, n_bytes, *d_bases, *d_inv, n_pos, n_bases);
    ctx->arena.release(mk);
    return rc;
}

// Cut points of a FASTA text at record boundaries ('>' at the start of a line), about `target` bytes apart.
// A record longer than `target` (a long genome) simply makes a longer chunk.
static void split_fasta(const uint8_t* t, uint64_t n, uint64_t target, std::vector<uint64_t>& cuts) {
    cuts.clear(); cuts.push_back(0);
    // a header is searched only in a window after each target point (short reads always have one there);
    // inside a long record no cut is made and the host never scans the whole text
    const uint64_t window = std::max<uint64_t>(1u << 16, target / 8);
    uint64_t pos = target;
    while (pos < n) {
        const uint8_t* p = t + pos;
        const uint8_t* lim = t + std::min<uint64_t>(n, pos + window);
        uint64_t cut = n;
        while (p < lim) {
            p = (const uint8_t*)memchr(p, '>', (size_t)(lim - p));
            if (!p) break;
            if (p > t && p[-1] == '\n') { cut = (uint64_t)(p - t); break; }
            p++;
        }
        if (cut >= n) { pos += target; continue; }
        cuts.push_back(cut);
        pos = cut + target;
    }
    cuts.push_back(n);
}

// Front end for FASTA text in host memory: the text is streamed to the GPU chunk by chunk on a copy
// stream while earlier chunks are parsed (fkm_ingest.cuh) and scanned (k_scan) on the compute stream,
// so the PCIe copy — the longest single step of an end-to-end job — hides the parse and the scan.
static int front_end_fasta(fkm_ctx* ctx, const fkm_config* cfg, int32_t B, const uint8_t* fasta, uint64_t n_bytes,
                           ScanState* S, uint64_t* n_bases_total, bool dual, bool speculate) {
    std::vector<uint64_t> cuts;
    split_fasta(fasta, n_bytes, (uint64_t)std::max(4096.0, ctx->ingest_chunk_bytes), cuts);
    const size_t nc = cuts.size() - 1;
    uint64_t max_chunk = 16;
    for (size_t c = 0; c < nc; c++) max_chunk = std::max(max_chunk, cuts[c + 1] - cuts[c]);
    int rc = scan_begin(ctx, cfg, B, S, dual, n_bytes); if (rc) return rc;
    uint8_t* d_text[2] = {nullptr, nullptr};
    CK(dmalloc(ctx, (void**)&d_text[0], max_chunk));
    if (nc > 1) CK(dmalloc(ctx, (void**)&d_text[1], max_chunk));
    *n_bases_total = 0;
    auto enqueue_copy = [&](size_t c) -> int {
        CK(cudaMemcpyAsync(d_text[c & 1], fasta + cuts[c], cuts[c + 1] - cuts[c], cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->copied[c & 1], ctx->copy_stream));
        return FKM_OK;
    };
    rc = enqueue_copy(0); if (rc) return rc;
    for (size_t c = 0; c < nc; c++) {
        // chunk c+1 goes on the wire before the host blocks on chunk c; its buffer was released when
        // the ingest of chunk c-1 completed (ingest_device synchronises)
        if (c + 1 < nc) { rc = enqueue_copy(c + 1); if (rc) return rc; }
        CK(cudaStreamWaitEvent(ctx->stream, ctx->copied[c & 1], 0));
        const uint64_t len = cuts[c + 1] - cuts[c];
        void *d_b = nullptr, *d_i = nullptr;
        const uint64_t cap_words = ingest_cap_words(len);
        CK(dmalloc(ctx, &d_b, cap_words * 8)); CK(dmalloc(ctx, &d_i, cap_words * 4));
        const Arena::Mark mk = ctx->arena.mark();
        uint64_t n_pos = 0, n_bases = 0;
        rc = ingest_device(ctx, d_text[c & 1], len, d_b, d_i, &n_pos, &n_bases);
        ctx->arena.release(mk);
        if (rc) return rc;
        *n_bases_total += n_bases;
        rc = scan_chunk(ctx, S, d_b, d_i, n_pos); if (rc) return rc;
        // ---- speculative scatter: once a quarter of the chunks has been scanned the bins' final sizes are forecast from the
        // histogram so far (+15 %), and from then on every scanned chunk is scattered while the next ones are still on the bus.
        // A bin that outgrows its region, or a chunk whose event list overflowed, drops the forecast: the job then scatters
        // again after the last chunk, with the exact offsets, as it always did.
        if (speculate && !S->spec.failed && nc >= 4) {
            ChunkScan& C = S->chunks.back();
            const size_t bB = (size_t)S->B * 8;
            CK(cudaMemcpyAsync(C.n_events2, C.d_count, 16, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(&C.ev_ovf, C.d_ovf, 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (!S->spec.on) CK(cudaMemcpyAsync(S->h_rec.data(), S->d_hist_rec, bB, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            C.n_events = C.n_events2[0];
            if (C.ev_ovf) S->spec.failed = true;
            else {
                if (!S->spec.on && (c + 1) * 4 >= nc) {
                    const double scale = (double)n_bytes / (double)cuts[c + 1] * 1.15;
                    S->spec.base.assign((size_t)S->B + 1, 0);
                    for (int32_t b = 0; b < S->B; b++)
                        S->spec.base[(size_t)b + 1] = S->spec.base[(size_t)b] + (unsigned long long)((double)S->h_rec[(size_t)b] * scale) + 1024ull;
                    const int rbytes = cfg->k > 32 ? 32 : 16;
                    S->spec.cursor_shift = cursor_shift_for(S->B);
                    CK(dmalloc(ctx, &S->spec.d_records, (size_t)S->spec.base[(size_t)S->B] * rbytes));
                    CK(dmalloc(ctx, &S->spec.d_bin_base, bB + 8)); CK(dmalloc(ctx, &S->spec.d_ovf, 8));
                    CK(dmalloc(ctx, &S->spec.d_cursor, ((size_t)S->B << S->spec.cursor_shift) * 8));
                    CK(cudaMemcpyAsync(S->spec.d_bin_base, S->spec.base.data(), bB + 8, cudaMemcpyHostToDevice, ctx->stream));
                    CK(cudaMemsetAsync(S->spec.d_cursor, 0, ((size_t)S->B << S->spec.cursor_shift) * 8, ctx->stream));
                    CK(cudaMemsetAsync(S->spec.d_ovf, 0, 8, ctx->stream));
                    S->spec.on = true;
                }
                if (S->spec.on)
                    for (; S->spec.next_chunk <= c; S->spec.next_chunk++) {
                        rc = scatter_chunk(ctx, S, S->chunks[S->spec.next_chunk], S->spec.d_bin_base, S->spec.d_cursor, S->spec.cursor_shift,
                                           S->spec.d_records, S->spec.d_ovf);
                        if (rc) return rc;
                    }
            }
        }
    }
    return FKM_OK;
}

extern "C" int fkm_count_fasta(fkm_ctx* ctx, const fkm_config* cfg, const uint8_t* fasta, uint64_t n_bytes,
                               fkm_result** out, fkm_stats* stats) {
    if (!ctx) return fkm_set_error(FKM_EINVAL, "ctx is NULL");
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    rc = job_begin(ctx); if (rc) return rc;
    auto t0 = std::chrono::steady_clock::now();
    fkm_stats local; fkm_stats* st = stats ? stats : &local;
    memset(st, 0, sizeof *st);
    ScanState S; uint64_t n_bases = 0;
    ctx->job_split = choose_split(ctx, cfg, B, n_bytes);
    rc = front_end_fasta(ctx, cfg, B << ctx->job_split, fasta, n_bytes, &S, &n_bases, want_smem_path(ctx, cfg),
                         cfg->use_ht && ctx->count_mode >= 2.0 && ctx->speculative_scatter >= 1.0); if (rc) return rc;
    fkm_stats tmp; memset(&tmp, 0, sizeof tmp);
    rc = scan_end(ctx, &S, &tmp); if (rc) return rc;
    const double ms_in = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    st->h2d_bytes = n_bytes + 32 * S.chunks.size(); st->n_bases = n_bases; st->ms_stage[0] = ms_in;
    rc = count_device(ctx, cfg, nullptr, nullptr, S.n_pos_total, out, st, nullptr, &S);
    st->d2h_bytes += tmp.d2h_bytes + 16 * S.chunks.size(); st->ms_total += ms_in;
    return rc;
}

// test hook: the device ingest's packed arrays, copied back
extern "C" int fkm_debug_pack_fasta_device(fkm_ctx* ctx, const uint8_t* fasta, uint64_t n_bytes, uint64_t* bases, uint32_t* invalid,
                                           uint64_t cap_positions, uint64_t* n_positions, uint64_t* n_bases) {
    if (!ctx) return fkm_set_error(FKM_EINVAL, "ctx is NULL");
    int rc = job_begin(ctx); if (rc) return rc;
    void *d_b = nullptr, *d_i = nullptr; uint64_t n_pos = 0, nb = 0;
    rc = upload_and_ingest(ctx, fasta, n_bytes, &d_b, &d_i, &n_pos, &nb); if (rc) return rc;
    if (n_positions) *n_positions = n_pos;
    if (n_bases) *n_bases = nb;
    if (bases && invalid) {
        const uint64_t nw = (n_pos + 31) / 32;
        if (nw * 32 > cap_positions) return fkm_set_error(FKM_EINVAL, "packed buffer too small");
        CK(cudaMemcpy(bases, d_b, nw * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(invalid, d_i, nw * 4, cudaMemcpyDeviceToHost));
    }
    return FKM_OK;
}

extern "C" int fkm_execute_job(fkm_ctx* ctx, const fkm_config* cfg, fkm_stats* stats) {
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    if (!cfg->dataset) return fkm_set_error(FKM_EINVAL, "dataset path is NULL");
    if (!ctx) return fkm_set_error(FKM_EINVAL, "ctx is NULL");
    CK(cudaSetDevice(ctx->device));
    uint8_t* text = nullptr; uint64_t n_text = 0;
    rc = fkm_read_file_pinned(cfg->dataset, &text, &n_text); if (rc) return rc;
    fkm_result* res = nullptr;
    fkm_stats local; fkm_stats* st = stats ? stats : &local;
    rc = fkm_count_fasta(ctx, cfg, text, n_text, &res, st);
    cudaFreeHost(text);
    if (rc) return rc;
    if (cfg->write) {                                        // writers are lazy in the reference: nothing is created when !write (SBKC:552-554)
        char dir[4096];
        rc = fkm_derive(cfg, nullptr, dir, sizeof dir);
        auto t0 = std::chrono::steady_clock::now();
        if (!rc) rc = fkm_result_write(res, dir);
        st->ms_stage[6] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        st->ms_total += st->ms_stage[6];
    }
    fkm_result_free(res);
    return rc;
}

// ------------------------------------------------------------------ staged entry points (multi-GPU)
static int front_end_fasta(fkm_ctx* ctx, const fkm_config* cfg, int32_t B, const uint8_t* fasta, uint64_t n_bytes,
                           ScanState* S, uint64_t* n_bases_total, bool dual, bool speculate);
extern "C" int32_t fkm_record_bytes(const fkm_config* cfg) { return (cfg && cfg->k > 32) ? 32 : 16; }

extern "C" int fkm_mg_scan(fkm_ctx* ctx, const fkm_config* cfg, const void* d_bases, const void* d_inv, uint64_t n_pos,
                           uint64_t* hist_rec, uint64_t* hist_kmer) {
    if (!ctx || !hist_rec || !hist_kmer) return fkm_set_error(FKM_EINVAL, "null argument");
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    rc = job_begin(ctx); if (rc) return rc;
    if (!ctx->mg_scan) ctx->mg_scan = new ScanState();
    ctx->mg_scan->valid = false;
    fkm_stats st; memset(&st, 0, sizeof st);
    ctx->job_split = choose_split(ctx, cfg, B, n_pos);
    B <<= ctx->job_split;                        // the histograms are per INTERNAL bin (fkm_job_bins entries)
    rc = stage_scan(ctx, cfg, B, d_bases, d_inv, n_pos, ctx->mg_scan, &st); if (rc) return rc;
    for (int b = 0; b < B; b++) { hist_rec[b] = ctx->mg_scan->h_rec[(size_t)b]; hist_kmer[b] = ctx->mg_scan->h_kmer[(size_t)b]; }
    return FKM_OK;
}

// same, from FASTA text in host memory (H2D + device ingest first)
extern "C" int fkm_mg_scan_fasta(fkm_ctx* ctx, const fkm_config* cfg, const uint8_t* fasta, uint64_t n_bytes,
                                 uint64_t* hist_rec, uint64_t* hist_kmer, uint64_t* n_bases) {
    if (!ctx || !hist_rec || !hist_kmer) return fkm_set_error(FKM_EINVAL, "null argument");
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    rc = job_begin(ctx); if (rc) return rc;
    if (!ctx->mg_scan) ctx->mg_scan = new ScanState();
    uint64_t nb = 0;
    ctx->job_split = choose_split(ctx, cfg, B, n_bytes);
    B <<= ctx->job_split;                        // the histograms are per INTERNAL bin (fkm_job_bins entries)
    rc = front_end_fasta(ctx, cfg, B, fasta, n_bytes, ctx->mg_scan, &nb, false, false); if (rc) return rc;
    if (n_bases) *n_bases = nb;
    fkm_stats st; memset(&st, 0, sizeof st);
    rc = scan_end(ctx, ctx->mg_scan, &st); if (rc) return rc;
    for (int b = 0; b < B; b++) { hist_rec[b] = ctx->mg_scan->h_rec[(size_t)b]; hist_kmer[b] = ctx->mg_scan->h_kmer[(size_t)b]; }
    return FKM_OK;
}

extern "C" int fkm_mg_scatter(fkm_ctx* ctx, const uint64_t* bin_base, void* d_send) {
    if (!ctx || !bin_base || !ctx->mg_scan || !ctx->mg_scan->valid) return fkm_set_error(FKM_EINVAL, "fkm_mg_scatter needs a preceding fkm_mg_scan");
    if (ctx->mg_scan->gen != ctx->gen)       // another job ran on this context since fkm_mg_scan: its arena (events, histograms, packed chunks) is gone
        return fkm_set_error(FKM_EINVAL, "fkm_mg_scatter: the scan was invalidated by a later job on the same context (run fkm_mg_scan again)");
    CK(cudaSetDevice(ctx->device));
    ScanState* S = ctx->mg_scan;
    const size_t bB = (size_t)S->B * 8;
    unsigned long long* d_bin_base = nullptr;
    CK(dmalloc(ctx, &d_bin_base, bB + 8));
    CK(cudaMemcpyAsync(d_bin_base, bin_base, bB + 8, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t n_rec = 0; for (int b = 0; b < S->B; b++) n_rec += S->h_rec[(size_t)b];
    fkm_stats st; memset(&st, 0, sizeof st);
    int rc = stage_scatter(ctx, S, d_bin_base, d_send, n_rec, &st); if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return FKM_OK;
}

extern "C" int fkm_mg_regroup(fkm_ctx* ctx, const fkm_config* cfg, const void* d_recv, uint64_t n_records,
                              const uint64_t* seg_src, const uint64_t* seg_dst, uint64_t n_seg, void** d_out) {
    if (!ctx || !cfg || !d_out) return fkm_set_error(FKM_EINVAL, "null argument");
    if (!ctx->mg_scan || ctx->mg_scan->gen != ctx->gen)
        return fkm_set_error(FKM_EINVAL, "fkm_mg_regroup belongs to the job that fkm_mg_scan started; a later job on the context invalidated it");
    CK(cudaSetDevice(ctx->device));
    const int rb = fkm_record_bytes(cfg);
    unsigned long long *d_src = nullptr, *d_dst = nullptr; void* out = nullptr;
    CK(dmalloc(ctx, &out, std::max<uint64_t>(n_records, 1) * rb));
    CK(dmalloc(ctx, &d_src, (n_seg + 1) * 8)); CK(dmalloc(ctx, &d_dst, (n_seg + 1) * 8));
    if (n_records && n_seg) {
        CK(cudaMemcpyAsync(d_src, seg_src, (n_seg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_dst, seg_dst, n_seg * 8, cudaMemcpyHostToDevice, ctx->stream));
        RegroupParams P; P.in = d_recv; P.out = out; P.n = n_records; P.rec_words = rb / 8; P.seg_src = d_src; P.seg_dst = d_dst; P.n_seg = (int)n_seg;
        k_regroup<<<(unsigned)((n_records + 255) / 256), 256, 0, ctx->stream>>>(P); CKL();
        CK(cudaStreamSynchronize(ctx->stream));
    }
    *d_out = out;
    return FKM_OK;
}

extern "C" int fkm_mg_count(fkm_ctx* ctx, const fkm_config* cfg, const void* d_records, const uint64_t* bin_rec, const uint64_t* bin_kmer,
                            fkm_result** out, fkm_stats* stats) {
    if (!ctx || !bin_rec || !bin_kmer) return fkm_set_error(FKM_EINVAL, "null argument");
    if (!ctx->mg_scan || ctx->mg_scan->gen != ctx->gen)
        return fkm_set_error(FKM_EINVAL, "fkm_mg_count belongs to the job that fkm_mg_scan started; a later job on the context invalidated it");
    if (stats) { stats->h2d_bytes = 0; stats->n_bases = 0; stats->ms_stage[0] = 0; }
    PreScattered pre{d_records, bin_rec, bin_kmer};
    return count_device(ctx, cfg, nullptr, nullptr, 0, out, stats, &pre);
}

// ------------------------------------------------------------------ N GPUs of one node behind the drop-in call
// The multi-GPU job of fastkmer_b200/multigpu.py without Python or NCCL: one host thread per GPU, the bin exchange as
// peer-to-peer copies over NVLink (cudaMemcpyPeerAsync).  Replaces Spark's shuffle between the executors of one job
// (reduceByKey, SBKC:1035,1042, optionally behind MultiprocessorSchedulingPartitioner, MSP:35-69):
//   1. the FASTA text is cut at record boundaries into one byte range per GPU; every GPU parses and scans its range
//   2. the per-bin histograms meet on the host; bins go to GPUs by LPT over the exact k-mer counts (the role of MSP:35-69)
//   3. every GPU writes its records owner-major, the GPUs copy each other's blocks, regroup them bin-major
//   4. every GPU counts the bins it owns and writes their files
#include <thread>
namespace {
struct MultiPlan {                              // the part of multigpu.plan_exchange() rank r needs
    std::vector<uint64_t> send_base;            // [Bi+1] record offset of every bin in r's owner-major send buffer
    std::vector<uint64_t> send_off;             // [n+1]  where the block for GPU g begins in it
    std::vector<uint64_t> recv_off;             // [n+1]  where the block from GPU s begins in r's receive buffer
    std::vector<uint64_t> bin_rec, bin_kmer;    // [Bi]   records / k-mers of the bins r owns (0 elsewhere)
    std::vector<uint64_t> seg_src, seg_dst;     // received segments (source, bin) -> bin-major place
};
// Owners (LPT over the bins of the configuration: longest first, each to the least loaded GPU; ties by bin id, then rank — the same
// map as multigpu.assign_owners) and every rank's exchange plan.  H_rec / H_kmer: [n][B * sp] records / k-mers every rank puts
// into every internal bin; the sp internal bins of a bin share its owner.
void multi_plan(int n, int B, int sp, const std::vector<std::vector<uint64_t>>& H_rec, const std::vector<std::vector<uint64_t>>& H_kmer,
                std::vector<int>& owner_bin, std::vector<MultiPlan>& plan) {
    const int Bi = B * sp;
    std::vector<uint64_t> tot((size_t)B, 0);
    for (int b = 0; b < Bi; b++) for (int i = 0; i < n; i++) tot[(size_t)(b / sp)] += H_kmer[(size_t)i][(size_t)b];
    std::vector<int> order((size_t)B); for (int b = 0; b < B; b++) order[(size_t)b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b2) { return tot[(size_t)a] > tot[(size_t)b2]; });
    std::vector<uint64_t> load((size_t)n, 0);
    owner_bin.assign((size_t)B, 0);
    for (int b : order) { int g = 0; for (int j = 1; j < n; j++) if (load[(size_t)j] < load[(size_t)g]) g = j; owner_bin[(size_t)b] = g; load[(size_t)g] += tot[(size_t)b]; }
    auto owner = [&](int bi) { return owner_bin[(size_t)(bi / sp)]; };
    plan.assign((size_t)n, MultiPlan());
    for (int r = 0; r < n; r++) {
        MultiPlan& P = plan[(size_t)r];
        P.send_base.assign((size_t)Bi + 1, 0); P.send_off.assign((size_t)n + 1, 0); P.recv_off.assign((size_t)n + 1, 0);
        P.bin_rec.assign((size_t)Bi, 0); P.bin_kmer.assign((size_t)Bi, 0);
        uint64_t o = 0;
        for (int g = 0; g < n; g++) {                                             // send buffer: bins ordered by (owner, bin)
            P.send_off[(size_t)g] = o;
            for (int b = 0; b < Bi; b++) if (owner(b) == g) { P.send_base[(size_t)b] = o; o += H_rec[(size_t)r][(size_t)b]; }
        }
        P.send_off[(size_t)n] = o; P.send_base[(size_t)Bi] = o;
        uint64_t ro = 0;
        for (int s2 = 0; s2 < n; s2++) { P.recv_off[(size_t)s2] = ro; for (int b = 0; b < Bi; b++) if (owner(b) == r) ro += H_rec[(size_t)s2][(size_t)b]; }
        P.recv_off[(size_t)n] = ro;
        std::vector<uint64_t> dst((size_t)Bi + 1, 0);
        for (int b = 0; b < Bi; b++) {
            if (owner(b) == r) for (int s2 = 0; s2 < n; s2++) { P.bin_rec[(size_t)b] += H_rec[(size_t)s2][(size_t)b]; P.bin_kmer[(size_t)b] += H_kmer[(size_t)s2][(size_t)b]; }
            dst[(size_t)b + 1] = dst[(size_t)b] + P.bin_rec[(size_t)b];
        }
        P.seg_src.push_back(0);
        std::vector<uint64_t> before((size_t)Bi, 0);
        for (int s2 = 0; s2 < n; s2++)
            for (int b = 0; b < Bi; b++)
                if (owner(b) == r && H_rec[(size_t)s2][(size_t)b]) {
                    P.seg_dst.push_back(dst[(size_t)b] + before[(size_t)b]);
                    P.seg_src.push_back(P.seg_src.back() + H_rec[(size_t)s2][(size_t)b]);
                    before[(size_t)b] += H_rec[(size_t)s2][(size_t)b];
                }
    }
}
}
// test hook (no GPU needed): the plan fkm_execute_job_multi makes for rank `rank` of n, from host histograms [n][bins * split]
extern "C" int fkm_debug_multi_plan(int32_t n, int32_t bins, int32_t split, const uint64_t* h_rec, const uint64_t* h_kmer, int32_t rank,
                                    int32_t* owner, uint64_t* send_base, uint64_t* send_off, uint64_t* recv_off, uint64_t* bin_rec, uint64_t* bin_kmer,
                                    uint64_t* seg_src, uint64_t* seg_dst, uint64_t seg_cap, uint64_t* n_seg) {
    if (n < 1 || bins < 1 || split < 1 || rank < 0 || rank >= n || !h_rec || !h_kmer) return fkm_set_error(FKM_EINVAL, "bad argument");
    const int Bi = bins * split;
    std::vector<std::vector<uint64_t>> R((size_t)n), K((size_t)n);
    for (int i = 0; i < n; i++) { R[(size_t)i].assign(h_rec + (size_t)i * Bi, h_rec + (size_t)(i + 1) * Bi); K[(size_t)i].assign(h_kmer + (size_t)i * Bi, h_kmer + (size_t)(i + 1) * Bi); }
    std::vector<int> ob; std::vector<MultiPlan> plan;
    multi_plan(n, bins, split, R, K, ob, plan);
    const MultiPlan& P = plan[(size_t)rank];
    if (P.seg_dst.size() > seg_cap) return fkm_set_error(FKM_EINVAL, "segment arrays too small");
    if (owner) for (int b = 0; b < Bi; b++) owner[b] = ob[(size_t)(b / split)];
    if (send_base) memcpy(send_base, P.send_base.data(), ((size_t)Bi + 1) * 8);
    if (send_off) memcpy(send_off, P.send_off.data(), ((size_t)n + 1) * 8);
    if (recv_off) memcpy(recv_off, P.recv_off.data(), ((size_t)n + 1) * 8);
    if (bin_rec) memcpy(bin_rec, P.bin_rec.data(), (size_t)Bi * 8);
    if (bin_kmer) memcpy(bin_kmer, P.bin_kmer.data(), (size_t)Bi * 8);
    if (seg_src) memcpy(seg_src, P.seg_src.data(), P.seg_src.size() * 8);
    if (seg_dst) memcpy(seg_dst, P.seg_dst.data(), P.seg_dst.size() * 8);
    if (n_seg) *n_seg = P.seg_dst.size();
    return FKM_OK;
}
extern "C" int fkm_execute_job_multi(const int32_t* devices, int32_t n_devices, const fkm_config* cfg, fkm_stats* stats) {
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    if (!devices || n_devices < 1 || n_devices > 64) return fkm_set_error(FKM_EINVAL, "bad device list");
    if (!cfg->dataset) return fkm_set_error(FKM_EINVAL, "dataset path is NULL");
    const int n = n_devices;
    std::vector<fkm_ctx*> ctx((size_t)n, nullptr);
    auto destroy = [&]() { for (auto* c : ctx) fkm_ctx_destroy(c); };
    for (int i = 0; i < n; i++) { rc = fkm_ctx_create(devices[i], nullptr, &ctx[(size_t)i]); if (rc) { destroy(); return rc; } }
    if (n == 1) { rc = fkm_execute_job(ctx[0], cfg, stats); const std::string keep = g_err; destroy(); g_err = keep; return rc; }
    int split = 1;
    if (cfg->use_ht) while (split < n && split < 64) split *= 2;               // see multigpu.ShardedJob
    for (auto* c : ctx) { c->bin_split = (double)split; if (c->count_mode >= 1.0 && c->count_mode < 2.0) c->count_mode = 2.0; }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (i != j) {
                cudaSetDevice(devices[i]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) { cudaGetLastError(); }              // no peer access: the copies below are staged by the driver
            }
    uint8_t* text = nullptr; uint64_t n_text = 0;
    rc = fkm_read_file_pinned(cfg->dataset, &text, &n_text); if (rc) { destroy(); return rc; }
    // ---- byte ranges at record boundaries
    std::vector<uint64_t> cut((size_t)n + 1, n_text);
    cut[0] = 0;
    for (int i = 1; i < n; i++) {
        uint64_t pos = std::max(cut[(size_t)i - 1], n_text / (uint64_t)n * (uint64_t)i), c = n_text;
        const uint8_t* p0 = text + pos;
        while (p0 < text + n_text) {
            p0 = (const uint8_t*)memchr(p0, '>', (size_t)(text + n_text - p0));
            if (!p0) break;
            if (p0 == text || p0[-1] == '\n') { c = (uint64_t)(p0 - text); break; }
            p0++;
        }
        cut[(size_t)i] = c;
    }
    int32_t Bi = 0;
    rc = fkm_job_bins(ctx[0], cfg, 0, &Bi); if (rc) { cudaFreeHost(text); destroy(); return rc; }
    const int sp = Bi / B;                                                        // internal bins per bin
    std::vector<std::vector<uint64_t>> H_rec((size_t)n, std::vector<uint64_t>((size_t)Bi)), H_kmer((size_t)n, std::vector<uint64_t>((size_t)Bi));
    std::vector<int> trc((size_t)n, FKM_OK); std::vector<std::string> terr((size_t)n);
    std::vector<uint64_t> nbases((size_t)n, 0);
    auto run_all = [&](auto&& fn) -> int {                                        // fn(rank) on one host thread per GPU
        std::vector<std::thread> th;
        for (int i = 0; i < n; i++) th.emplace_back([&, i]() { trc[(size_t)i] = fn(i); if (trc[(size_t)i]) terr[(size_t)i] = g_err; });
        for (auto& t : th) t.join();
        for (int i = 0; i < n; i++) if (trc[(size_t)i]) { g_err = "GPU " + std::to_string(devices[i]) + ": " + terr[(size_t)i]; return trc[(size_t)i]; }
        return FKM_OK;
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<void*> d_send((size_t)n, nullptr), d_recv((size_t)n, nullptr);
    auto cleanup = [&]() {
        for (int i = 0; i < n; i++) { cudaSetDevice(devices[i]); cudaDeviceSynchronize(); cudaFree(d_send[(size_t)i]); cudaFree(d_recv[(size_t)i]); }
        cudaFreeHost(text); destroy();
    };
    rc = run_all([&](int i) { return fkm_mg_scan_fasta(ctx[(size_t)i], cfg, text + cut[(size_t)i], cut[(size_t)i + 1] - cut[(size_t)i],
                                                       H_rec[(size_t)i].data(), H_kmer[(size_t)i].data(), &nbases[(size_t)i]); });
    if (rc) { const std::string keep = g_err; cleanup(); g_err = keep; return rc; }
    // ---- owners and the exchange plan of every rank
    const int rb = fkm_record_bytes(cfg);
    std::vector<int> owner_bin; std::vector<MultiPlan> plan;
    multi_plan(n, B, sp, H_rec, H_kmer, owner_bin, plan);
    // ---- records owner-major, GPU-to-GPU copies, bin-major again, count, write
    rc = run_all([&](int i) -> int {
        CK(cudaSetDevice(devices[i]));
        CK(cudaMalloc(&d_send[(size_t)i], std::max<uint64_t>(plan[(size_t)i].send_off[(size_t)n], 1) * rb));
        CK(cudaMalloc(&d_recv[(size_t)i], std::max<uint64_t>(plan[(size_t)i].recv_off[(size_t)n], 1) * rb));
        return fkm_mg_scatter(ctx[(size_t)i], plan[(size_t)i].send_base.data(), d_send[(size_t)i]);
    });
    if (rc) { const std::string keep = g_err; cleanup(); g_err = keep; return rc; }
    auto tx0 = std::chrono::steady_clock::now();
    rc = run_all([&](int s2) -> int {                                             // rank s2 pushes its blocks to their owners
        CK(cudaSetDevice(devices[s2]));
        for (int g = 0; g < n; g++) {
            const uint64_t cnt = plan[(size_t)s2].send_off[(size_t)g + 1] - plan[(size_t)s2].send_off[(size_t)g];
            if (!cnt) continue;
            CK(cudaMemcpyPeerAsync((char*)d_recv[(size_t)g] + plan[(size_t)g].recv_off[(size_t)s2] * rb, devices[g],
                                   (const char*)d_send[(size_t)s2] + plan[(size_t)s2].send_off[(size_t)g] * rb, devices[s2], cnt * rb, ctx[(size_t)s2]->stream));
        }
        CK(cudaStreamSynchronize(ctx[(size_t)s2]->stream));
        return FKM_OK;
    });
    if (rc) { const std::string keep = g_err; cleanup(); g_err = keep; return rc; }
    const double ms_x = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tx0).count();
    std::vector<fkm_stats> sts((size_t)n);
    char dir[4096] = {0};
    if (cfg->write) { rc = fkm_derive(cfg, nullptr, dir, sizeof dir); if (rc) { cleanup(); return rc; } }
    rc = run_all([&](int g) -> int {
        const MultiPlan& P = plan[(size_t)g];
        void* d_records = nullptr;
        int r2 = fkm_mg_regroup(ctx[(size_t)g], cfg, d_recv[(size_t)g], P.recv_off[(size_t)n], P.seg_src.data(), P.seg_dst.data(), P.seg_dst.size(), &d_records);
        if (r2) return r2;
        fkm_result* res = nullptr;
        memset(&sts[(size_t)g], 0, sizeof(fkm_stats));
        r2 = fkm_mg_count(ctx[(size_t)g], cfg, d_records, P.bin_rec.data(), P.bin_kmer.data(), &res, &sts[(size_t)g]);
        if (r2) return r2;
        if (cfg->write) r2 = fkm_result_write(res, dir);                          // only the bins this GPU owns hold entries: their files
        fkm_result_free(res);
        return r2;
    });
    if (rc) { const std::string keep = g_err; cleanup(); g_err = keep; return rc; }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int g = 0; g < n; g++) {
            const fkm_stats& a = sts[(size_t)g];
            stats->n_kmers += a.n_kmers; stats->n_superkmers += a.n_superkmers; stats->superkmer_bytes += a.superkmer_bytes;
            stats->n_distinct += a.n_distinct; stats->total_count += a.total_count; stats->digest_sum += a.digest_sum; stats->digest_xor ^= a.digest_xor;
            stats->n_nonempty_bins += a.n_nonempty_bins; stats->gpu_launches += a.gpu_launches; stats->n_batches += a.n_batches;
            stats->n_fallbacks += a.n_fallbacks; stats->n_mid_bins += a.n_mid_bins; stats->n_slow_bins += a.n_slow_bins;
            stats->h2d_bytes += a.h2d_bytes + (cut[(size_t)g + 1] - cut[(size_t)g]); stats->d2h_bytes += a.d2h_bytes;
            stats->n_bases += nbases[(size_t)g];
            for (int j = 0; j < 8; j++) stats->ms_stage[j] = std::max(stats->ms_stage[j], a.ms_stage[j]);
            stats->ms_partition = std::max(stats->ms_partition, a.ms_partition); stats->ms_fold = std::max(stats->ms_fold, a.ms_fold);
        }
        stats->ms_stage[6] = ms_x;                                                // (the GPU-to-GPU exchange, host clock)
        stats->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    cleanup();
    return FKM_OK;
}

// ------------------------------------------------------------------ multi-sample distances
// The reference's prototype (skc.multisequence, MSKC:29-165,300-547) tags super-k-mers with the sample a read came
// from, keeps per-sample counts of every distinct k-mer of a bin and adds (c_a - c_b)^2 to the distance of every
// sample pair (multiseq/SquaredEuclidean.java:19-32); the shipped job never executes (MSKC:585-588), so the
// semantics are the intended ones of SURVEY App. A.7: the sample of a read is the leading \w+ of its header.
// Here: reads are split by sample on the host, every sample is counted with the sorted pipeline, and
// k_sparse_dot gives sum_k c_a(k) c_b(k) for every pair, all in 64-bit integers.
extern "C" int fkm_multiseq_fasta(fkm_ctx* ctx, const fkm_config* cfg, const uint8_t* fasta, uint64_t n_bytes, int32_t max_samples,
                                  int32_t* n_samples, char* names, double* dist, fkm_result** merged, fkm_stats* stats) {
    if (!ctx || !n_samples || !dist || max_samples < 1) return fkm_set_error(FKM_EINVAL, "bad argument");
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    // ---- split the records by sample tag (first run of [A-Za-z0-9_] of the header).  Two passes: the record boundaries
    // (memchr for '>' at a line start) and their samples first, then every sample's text is gathered by its own host thread.
    std::vector<std::string> tags; std::vector<std::vector<uint8_t>> texts;
    {
        const uint8_t* t = fasta;
        struct Rec { uint64_t lo, hi; uint32_t sample; };
        std::vector<Rec> recs;
        auto next_header = [&](uint64_t from) -> uint64_t {         // first '>' at the start of a line at or after `from`
            while (from < n_bytes) {
                const uint8_t* p0 = (const uint8_t*)memchr(t + from, '>', (size_t)(n_bytes - from));
                if (!p0) return n_bytes;
                const uint64_t at = (uint64_t)(p0 - t);
                if (at == 0 || t[at - 1] == '\n') return at;
                from = at + 1;
            }
            return n_bytes;
        };
        auto isw = [](uint8_t c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_'; };
        size_t last = 0;
        for (uint64_t i = next_header(0); i < n_bytes;) {
            uint64_t j = i + 1;
            while (j < n_bytes && t[j] != '\n' && !isw(t[j])) j++;
            uint64_t e = j; while (e < n_bytes && isw(t[e])) e++;
            const size_t len = (size_t)(e - j);
            const uint64_t nxt = next_header(i + 1);
            size_t sidx = last;                                       // the records of a sample usually follow each other
            if (!(sidx < tags.size() && tags[sidx].size() == len && !memcmp(tags[sidx].data(), t + j, len))) {
                for (sidx = 0; sidx < tags.size(); sidx++) if (tags[sidx].size() == len && !memcmp(tags[sidx].data(), t + j, len)) break;
                if (sidx == tags.size()) {
                    if ((int32_t)tags.size() == max_samples) return fkm_set_error(FKM_EINVAL, "more than %d samples in the input", max_samples);
                    tags.emplace_back((const char*)t + j, len);
                }
            }
            last = sidx;
            recs.push_back(Rec{i, nxt, (uint32_t)sidx});
            i = nxt;
        }
        texts.resize(tags.size());
        std::vector<uint64_t> bytes(tags.size(), 0);
        for (const Rec& r : recs) bytes[r.sample] += r.hi - r.lo + 1;
        std::vector<std::thread> th;
        for (size_t a = 0; a < tags.size(); a++)
            th.emplace_back([&, a]() {
                std::vector<uint8_t>& out = texts[a];
                out.reserve((size_t)bytes[a]);
                for (const Rec& r : recs)
                    if (r.sample == a) {
                        out.insert(out.end(), t + r.lo, t + r.hi);
                        if (out.back() != '\n') out.push_back('\n');
                    }
            });
        for (auto& x : th) x.join();
    }
    const int S = (int)tags.size();
    *n_samples = S;
    if (names) for (int a = 0; a < S; a++) { memset(names + 64 * a, 0, 64); strncpy(names + 64 * a, tags[a].c_str(), 63); }
    for (int a = 0; a < max_samples * max_samples; a++) dist[a] = 0.0;
    // ---- count every sample (sorted per bin) and keep a clone of each result
    fkm_config c2 = *cfg; c2.use_ht = 0; if (c2.x < 1) c2.x = 1;
    std::vector<fkm_result*> sc((size_t)S, nullptr);
    auto release = [&]() { for (auto* q : sc) fkm_result_free(q); };
    fkm_stats total; memset(&total, 0, sizeof total);
    for (int a = 0; a < S && !rc; a++) {
        fkm_result* r = nullptr; fkm_stats st;
        rc = fkm_count_fasta(ctx, &c2, texts[(size_t)a].data(), texts[(size_t)a].size(), &r, &st);
        if (rc) break;
        rc = fkm_result_clone(r, &sc[(size_t)a]);
        fkm_result_free(r);
        total.n_bases += st.n_bases; total.n_kmers += st.n_kmers; total.n_distinct += st.n_distinct; total.total_count += st.total_count;
        total.gpu_launches += st.gpu_launches; total.h2d_bytes += st.h2d_bytes; total.d2h_bytes += st.d2h_bytes; total.ms_total += st.ms_total;
    }
    // ---- pairwise sparse dot products
    std::vector<unsigned long long> dot((size_t)S * S, 0);
    for (int a = 0; a < S && !rc; a++)
        for (int b = a; b < S && !rc; b++) {
            uint64_t v = 0;
            rc = fkm_result_dot(ctx, sc[(size_t)a], sc[(size_t)b], &v);
            dot[(size_t)a * S + b] = v; total.gpu_launches++;
        }
    release();
    if (rc) return rc;
    for (int a = 0; a < S; a++)
        for (int b = a + 1; b < S; b++) {
            // sum_k (c_a - c_b)^2 = dot(a,a) + dot(b,b) - 2 dot(a,b), exact in 64-bit integers, then to double
            const unsigned long long v = dot[(size_t)a * S + a] + dot[(size_t)b * S + b] - 2ull * dot[(size_t)a * S + b];
            dist[(size_t)a * max_samples + b] = dist[(size_t)b * max_samples + a] = (double)v;
        }
    // ---- the merged counts (kmer, sum over samples): the ordinary job on the whole input (MSKC:487,524)
    if (merged || cfg->write) {
        fkm_result* r = nullptr; fkm_stats st;
        rc = fkm_count_fasta(ctx, &c2, fasta, n_bytes, &r, &st); if (rc) return rc;
        r->eof_trailer = false;
        total.n_nonempty_bins = st.n_nonempty_bins; total.digest_sum = st.digest_sum; total.digest_xor = st.digest_xor;
        total.gpu_launches += st.gpu_launches; total.ms_total += st.ms_total;
        if (cfg->write) {                                   // the reference writes the merged counts per bin (MSKC:487,524), without a trailer
            char dir[4096];
            rc = fkm_derive(cfg, nullptr, dir, sizeof dir);
            if (!rc) rc = fkm_result_write(r, dir);
            if (rc) { fkm_result_free(r); return rc; }
        }
        if (merged) *merged = r; else fkm_result_free(r);
    }
    if (stats) *stats = total;
    return FKM_OK;
}

// ------------------------------------------------------------------ results
extern "C" uint64_t fkm_result_size(const fkm_result* r) { return r ? r->total : 0; }
extern "C" int32_t fkm_result_num_bins(const fkm_result* r) { return r ? r->B : 0; }
extern "C" int32_t fkm_result_sorted(const fkm_result* r) { return r && r->sorted ? 1 : 0; }
extern "C" int fkm_result_bin_offsets(const fkm_result* r, uint64_t* offsets) {
    if (!r || !offsets) return fkm_set_error(FKM_EINVAL, "null argument");
    memcpy(offsets, r->out_base.data(), r->out_base.size() * 8);
    return FKM_OK;
}
extern "C" int fkm_result_copy(const fkm_result* r, int32_t* bin, uint64_t* key_hi, uint64_t* key_lo, uint32_t* count) {
    if (!r) return fkm_set_error(FKM_EINVAL, "null result");
    if (!r->owned && r->total && r->gen != r->ctx->gen)
        return fkm_set_error(FKM_EINVAL, "result was invalidated by a later job on the same context (copy it out first)");
    CK(cudaSetDevice(r->device));
    uint64_t o = 0;
    std::vector<uint64_t> tmp;
    for (const Chunk& ch : r->chunks) {
        if (!ch.n) continue;
        if (count) CK(cudaMemcpy(count + o, ch.cnt, ch.n * 4, cudaMemcpyDeviceToHost));
        if (key_hi || key_lo) {
            if (!r->wide) {
                if (key_lo) CK(cudaMemcpy(key_lo + o, ch.keys, ch.n * 8, cudaMemcpyDeviceToHost));
                if (key_hi) memset(key_hi + o, 0, ch.n * 8);
            } else {
                tmp.resize(ch.n * 2);
                CK(cudaMemcpy(tmp.data(), ch.keys, ch.n * 16, cudaMemcpyDeviceToHost));
                for (uint64_t i = 0; i < ch.n; i++) { if (key_lo) key_lo[o + i] = tmp[2 * i]; if (key_hi) key_hi[o + i] = tmp[2 * i + 1]; }
            }
        }
        o += ch.n;
    }
    if (bin)
        for (int b = 0; b < r->B; b++)
            for (uint64_t i = r->out_base[(size_t)b]; i < r->out_base[(size_t)b + 1]; i++) bin[i] = b;
    return FKM_OK;
}
// Per-bin text files (SBKC:550-606 sort path with the "EOF" trailer, SBKC:715-734 HT path).  The lines are
// formatted on the device (k_fmt), copied to pinned host memory in pieces and appended to <out_dir>/bin<id>.
template <bool WIDE>
static int write_result_device(fkm_ctx* ctx, const fkm_result* r, const char* out_dir) {
    typedef typename Traits<WIDE>::Key Key;
    cudaStream_t s = ctx->stream;
    int rc = fkm_make_dirs(out_dir); if (rc) return rc;
    const uint64_t piece = (uint64_t)32 << 20;                                   // entries per device pass
    const uint64_t max_line = (uint64_t)r->k + 12;
    const uint64_t n_tiles_max = (piece + 255) / 256;
    unsigned long long *d_tile = nullptr, *d_bounds = nullptr, *d_boff = nullptr; uint8_t* d_text = nullptr;
    const Arena::Mark mk = ctx->arena.mark();
    CK(dmalloc(ctx, &d_tile, (n_tiles_max + 1) * 8));
    CK(dmalloc(ctx, &d_bounds, ((size_t)r->B + 2) * 8)); CK(dmalloc(ctx, &d_boff, ((size_t)r->B + 2) * 8));
    CK(dmalloc(ctx, (void**)&d_text, std::min<uint64_t>(piece, std::max<uint64_t>(r->total, 1)) * max_line));
    uint8_t* h_text = nullptr;
    CK(cudaMallocHost((void**)&h_text, std::min<uint64_t>(piece, std::max<uint64_t>(r->total, 1)) * max_line));
    std::vector<unsigned long long> bounds, boff;
    std::vector<int> bbin;
    FILE* f = nullptr; int open_bin = -1;
    auto close_bin = [&]() { if (f) { if (r->sorted && r->eof_trailer) fputs("EOF", f); fclose(f); f = nullptr; } };
    uint64_t origin = 0;
    int bin = 0;
    rc = FKM_OK;
    for (const Chunk& ch : r->chunks) {
        for (uint64_t first = 0; first < ch.n && !rc; first += piece) {
            const uint64_t n = std::min(piece, ch.n - first);
            const uint64_t g0 = origin + first, g1 = g0 + n;                     // global entry range of this piece
            // bin boundaries inside the piece
            bounds.clear(); bbin.clear();
            while (bin < r->B && r->out_base[(size_t)bin + 1] <= g0) bin++;
            int b = bin;
            bounds.push_back(0); bbin.push_back(b);
            while (b < r->B && r->out_base[(size_t)b + 1] < g1) { b++; bounds.push_back(r->out_base[(size_t)b] - g0); bbin.push_back(b); }
            bounds.push_back(n);
            FmtParams P;
            P.keys = ch.keys; P.cnt = ch.cnt; P.first = first; P.n = n; P.k = r->k; P.tile_off = d_tile; P.text = d_text;
            P.bounds = d_bounds; P.bound_off = d_boff; P.n_bounds = (int)bounds.size();
            const unsigned n_tiles = (unsigned)((n + 255) / 256);
            auto fail = [&](cudaError_t e) { return fkm_set_error(FKM_ECUDA, "text output: %s", cudaGetErrorString(e)); };
            cudaError_t e = cudaMemcpyAsync(d_bounds, bounds.data(), bounds.size() * 8, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) { rc = fail(e); break; }
            k_fmt<WIDE, 0><<<n_tiles, 256, 0, s>>>(P); g_launches++;
            k_scan1<0><<<1, 1024, 0, s>>>((long long*)d_tile, n_tiles); g_launches++;
            k_fmt<WIDE, 1><<<n_tiles, 256, 0, s>>>(P); g_launches++;
            k_fmt_bounds<WIDE><<<(P.n_bounds + 127) / 128, 128, 0, s>>>(P); g_launches++;
            boff.resize(bounds.size());
            e = cudaMemcpyAsync(boff.data(), d_boff, bounds.size() * 8, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { rc = fail(e); break; }
            e = cudaMemcpyAsync(h_text, d_text, boff.back(), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { rc = fail(e); break; }
            for (size_t j = 0; j + 1 < bounds.size(); j++) {
                if (boff[j + 1] == boff[j]) continue;
                if (bbin[j] != open_bin) {
                    close_bin();
                    const std::string path = std::string(out_dir) + "/bin" + std::to_string(bbin[j]);
                    f = fopen(path.c_str(), "wb"); open_bin = bbin[j];
                    if (!f) { rc = fkm_set_error(FKM_EIO, "cannot create %s", path.c_str()); break; }
                }
                if (fwrite(h_text + boff[j], 1, boff[j + 1] - boff[j], f) != boff[j + 1] - boff[j]) { rc = fkm_set_error(FKM_EIO, "short write under %s", out_dir); break; }
            }
        }
        origin += ch.n;
        if (rc) break;
    }
    close_bin();
    cudaFreeHost(h_text);
    ctx->arena.release(mk);
    return rc;
}

extern "C" int fkm_result_write(const fkm_result* r, const char* out_dir) {
    if (!r || !out_dir) return fkm_set_error(FKM_EINVAL, "null argument");
    if (!r->owned && r->total && r->gen != r->ctx->gen)
        return fkm_set_error(FKM_EINVAL, "result was invalidated by a later job on the same context (write it first)");
    CK(cudaSetDevice(r->device));
    return r->wide ? write_result_device<true>(r->ctx, r, out_dir) : write_result_device<false>(r->ctx, r, out_dir);
}
extern "C" void fkm_result_free(fkm_result* r) {
    if (!r) return;
    if (r->owned) {                                // a clone owns its arrays; everything else lives in the context arena
        cudaSetDevice(r->device);
        for (Chunk& ch : r->chunks) { cudaFree(ch.keys); cudaFree(ch.cnt); }
        cudaFree(r->d_base);
    }
    delete r;
}

// A copy of a result in plain device memory (one chunk): it stays valid when later jobs reuse the context's arena.
extern "C" int fkm_result_clone(const fkm_result* r, fkm_result** out) {
    if (!r || !out) return fkm_set_error(FKM_EINVAL, "null argument");
    if (!r->owned && r->total && r->gen != r->ctx->gen) return fkm_set_error(FKM_EINVAL, "result was invalidated by a later job on the same context");
    CK(cudaSetDevice(r->device));
    const size_t ksz = r->wide ? 16 : 8;
    fkm_result* c = new fkm_result();
    c->owned = true; c->eof_trailer = r->eof_trailer; c->ctx = r->ctx; c->gen = 0; c->device = r->device;
    c->B = r->B; c->k = r->k; c->wide = r->wide; c->sorted = r->sorted; c->out_base = r->out_base; c->total = r->total;
    Chunk ch; ch.n = r->total;
    cudaError_t e = cudaMalloc(&ch.keys, std::max<uint64_t>(r->total, 1) * ksz);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ch.cnt, std::max<uint64_t>(r->total, 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_base, ((size_t)r->B + 1) * 8);
    c->chunks.push_back(ch);
    uint64_t o = 0;
    cudaStream_t s = r->ctx->stream;
    for (const Chunk& src : r->chunks) {
        if (e != cudaSuccess || !src.n) continue;
        e = cudaMemcpyAsync((char*)ch.keys + o * ksz, src.keys, src.n * ksz, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ch.cnt + o, src.cnt, src.n * 4, cudaMemcpyDeviceToDevice, s);
        o += src.n;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_base, c->out_base.data(), ((size_t)r->B + 1) * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { fkm_result_free(c); return fkm_set_error(e == cudaErrorMemoryAllocation ? FKM_ENOMEM : FKM_ECUDA, "clone: %s", cudaGetErrorString(e)); }
    *out = c;
    return FKM_OK;
}

// sum over the (bin, k-mer) pairs present in both results of count_a * count_b.  Both must be clones of sorted
// (use_ht=0) results of the same configuration.  (MSKC:474-482: the cross term of the squared euclidean distance.)
extern "C" int fkm_result_dot(fkm_ctx* ctx, const fkm_result* a, const fkm_result* b, uint64_t* dot) {
    if (!ctx || !a || !b || !dot) return fkm_set_error(FKM_EINVAL, "null argument");
    if (!a->owned || !b->owned || !a->sorted || !b->sorted || a->B != b->B || a->k != b->k)
        return fkm_set_error(FKM_EINVAL, "fkm_result_dot needs two clones of sorted results of one configuration");
    CK(cudaSetDevice(ctx->device));
    *dot = 0;
    if (!a->total || !b->total) return FKM_OK;
    unsigned long long* d_acc = nullptr;
    CK(cudaMalloc((void**)&d_acc, 8));
    CK(cudaMemsetAsync(d_acc, 0, 8, ctx->stream));
    DotParams P;
    P.keysA = a->chunks[0].keys; P.cntA = a->chunks[0].cnt; P.baseA = a->d_base; P.nA = a->total;
    P.keysB = b->chunks[0].keys; P.cntB = b->chunks[0].cnt; P.baseB = b->d_base; P.B = a->B; P.acc = d_acc;
    const int grid = (int)std::min<uint64_t>((P.nA + 255) / 256, (uint64_t)ctx->n_sm * 8);
    if (a->wide) k_sparse_dot<true><<<grid, 256, 0, ctx->stream>>>(P); else k_sparse_dot<false><<<grid, 256, 0, ctx->stream>>>(P);
    g_launches++;
    unsigned long long v = 0;
    cudaError_t e = cudaMemcpyAsync(&v, d_acc, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_acc);
    if (e != cudaSuccess) return fkm_set_error(FKM_ECUDA, "dot: %s", cudaGetErrorString(e));
    *dot = v;
    return FKM_OK;
}

// ------------------------------------------------------------------ synthetic data, test hooks
extern "C" int fkm_synth_packed_device(fkm_ctx* ctx, const fkm_synth* sy, void** d_bases, void** d_inv, uint64_t* n_positions) {
    if (!ctx || !sy || !d_bases || !d_inv) return fkm_set_error(FKM_EINVAL, "null argument");
    if (sy->genome_len < sy->read_len || sy->read_len == 0) return fkm_set_error(FKM_EINVAL, "genome shorter than a read");
    CK(cudaSetDevice(ctx->device));
    SynthParams P;
    P.S = SynthSpec{sy->seed_genome, sy->seed_reads, sy->seed_errors, sy->genome_len, sy->n_reads, sy->read_len, sy->first_read};
    P.n_pos = sy->n_reads * (sy->read_len + 1); P.n_words = (P.n_pos + 31) / 32;
    // caller-owned (outlives jobs): plain cudaMalloc, not the job arena
    CK(cudaMalloc((void**)&P.bases, std::max<size_t>(8, P.n_words * 8)));
    cudaError_t e = cudaMalloc((void**)&P.inv, std::max<size_t>(4, P.n_words * 4));
    if (e != cudaSuccess) { cudaFree(P.bases); CK(e); }
    if (P.n_words) { k_synth<<<(unsigned)((P.n_words + 255) / 256), 256, 0, ctx->stream>>>(P); CKL(); }
    CK(cudaStreamSynchronize(ctx->stream));
    *d_bases = P.bases; *d_inv = P.inv; if (n_positions) *n_positions = P.n_pos;
    return FKM_OK;
}
extern "C" int fkm_synth_long_packed_device(fkm_ctx* ctx, const fkm_synth_long* sy, void** d_bases, void** d_inv, uint64_t* n_positions) {
    if (!ctx || !sy || !d_bases || !d_inv) return fkm_set_error(FKM_EINVAL, "null argument");
    CK(cudaSetDevice(ctx->device));
    SynthLongParams P;
    P.S = LongSpec{sy->seed_genome, sy->seed_repeats, sy->seed_n, sy->first_pos};
    P.n_pos = sy->n_bases + 1; P.n_words = (P.n_pos + 31) / 32;
    CK(cudaMalloc((void**)&P.bases, std::max<size_t>(8, P.n_words * 8)));
    cudaError_t e = cudaMalloc((void**)&P.inv, std::max<size_t>(4, P.n_words * 4));
    if (e != cudaSuccess) { cudaFree(P.bases); CK(e); }
    k_synth_long<<<(unsigned)((P.n_words + 255) / 256), 256, 0, ctx->stream>>>(P); CKL();
    CK(cudaStreamSynchronize(ctx->stream));
    *d_bases = P.bases; *d_inv = P.inv; if (n_positions) *n_positions = P.n_pos;
    return FKM_OK;
}

extern "C" int fkm_device_free(fkm_ctx* ctx, void* p) {
    if (!ctx) return fkm_set_error(FKM_EINVAL, "ctx is NULL");
    CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->stream)); CK(cudaFree(p));
    return FKM_OK;
}

extern "C" int fkm_debug_window_bins(fkm_ctx* ctx, const fkm_config* cfg, const uint64_t* bases, const uint32_t* inv,
                                     uint64_t n_pos, int32_t* bins_out) {
    int32_t B = 0; int rc = validate(cfg, &B); if (rc) return rc;
    if (!ctx) return fkm_set_error(FKM_EINVAL, "ctx is NULL");
    rc = job_begin(ctx); if (rc) return rc;
    const uint64_t nw = (n_pos + 31) / 32;
    void *d_b = nullptr, *d_i = nullptr; int32_t* d_o = nullptr;
    CK(dmalloc(ctx, &d_b, std::max<size_t>(8, nw * 8))); CK(dmalloc(ctx, &d_i, std::max<size_t>(4, nw * 4))); CK(dmalloc(ctx, (void**)&d_o, std::max<size_t>(4, n_pos * 4)));
    CK(cudaMemcpyAsync(d_b, bases, nw * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_i, inv, nw * 4, cudaMemcpyHostToDevice, ctx->stream));
    ScanSetup S; S.grid = 1; S.smem = 0; S.fn = nullptr;
    rc = scan_setup<2, false>(ctx, cfg, B, d_b, d_i, n_pos, &S);
    if (!rc) { S.P.dbg_bins = d_o; S.fn<<<S.grid, kScanThreads, S.smem, ctx->stream>>>(S.P); }
    if (!rc) { CKL(); CK(cudaMemcpyAsync(bins_out, d_o, n_pos * 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); }
    dfree(ctx, d_b); dfree(ctx, d_i); dfree(ctx, d_o);
    return rc;
}
