// fkm_smem.cuh — the shared-memory count path (useHT = 1): hash tables that live in one SM's shared memory.
//
// Replaces extractKXmersHT (SBKC:664-739: one Object2IntOpenHashMap per bin, addTo per k-mer, dump of the entries) for
// inputs of any size.  A bin of the reference holds millions of k-mers, far more than fit on chip, so every bin is cut
// into "mid bins" small enough for one table of ~14 K slots in shared memory:
//
//   k_scan<0,NL,true>   runs of windows that share the signature AND a second, hash-ordered minimizer (length m2);
//                       the second minimizer's hash picks one of 2^cell_bits cells of the bin; per-cell histograms
//   k_cells_assign      consecutive cells of a bin -> mid bins of about T k-mers (prefix sum / T), exact sizes
//   k_excl_scan_u64     record offsets of the cells and of the mid bins (a mid bin is a run of consecutive cells)
//   k_scatter2          run events -> super-k-mer records, mid-bin-major (the "shuffle", SBKC:1035)
//   k_count_smem        persistent CTAs, one mid bin at a time: the records arrive in shared memory through the bulk-copy
//                       engine (cp.async.bulk + mbarrier, double buffered), every k-mer is inserted into the shared-memory
//                       table (ATOMS.CAS on the key, ATOMS.ADD on the count: 3-10 cycles per warp instruction per SM,
//                       scripts/microbench/ub_ops.cu), and the distinct (k-mer, count) pairs leave ONCE, densely, into the
//                       CTA's own region of the output (every CTA owns a contiguous range of mid bins): the output is
//                       bin-major without a compaction pass, no CTA waits for another, and DRAM sees the records once
//                       and the result once.
//
// A canonical k-mer has ONE signature and ONE second minimizer whatever the strand and the read it is seen in, so all of
// its occurrences meet in the same mid bin.  A mid bin whose distinct k-mers exceed the table is redone by the same CTA
// in a private global-memory table (slow path); if that overflows too, a flag sends the whole job to the global-table
// pipeline of fkm_lib.cu.
#pragma once
#include "fkm_kernels.cuh"

namespace fkm {

static constexpr int kSmWarps = 32;                          // warps of k_count_smem (thread 0 also issues the bulk copies)
static constexpr int kSmThreads = kSmWarps * 32;
static constexpr int kSmBlock = kSmThreads;
static constexpr int kSmMaxProbe = 4096;                     // slow path: probes before a table counts as full
static constexpr uint32_t kNoMid = 0xFFFFFFFFu;

// per-bin totals of a dual scan's (bin, cell) histograms: one warp per bin (the dual scan keeps no per-bin histogram)
__global__ void __launch_bounds__(256) k_cells_bin_totals(const unsigned long long* cell_rec, const unsigned long long* cell_kmer, int cell_bits, int B,
                                                          unsigned long long* hist_rec, unsigned long long* hist_kmer) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= B) return;
    unsigned long long r = 0, km = 0;
    const uint32_t C = 1u << cell_bits;
    for (uint32_t c = lane; c < C; c += 32) { r += cell_rec[((size_t)b << cell_bits) + c]; km += cell_kmer[((size_t)b << cell_bits) + c]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) { r += __shfl_xor_sync(0xFFFFFFFFu, r, o); km += __shfl_xor_sync(0xFFFFFFFFu, km, o); }
    if (lane == 0) { hist_rec[b] = r; hist_kmer[b] = km; }
}

// ------------------------------------------------------------------ cells -> mid bins
struct CellsParams {
    const unsigned long long* cell_rec; const unsigned long long* cell_kmer; int cell_bits;
    int bin_lo, bin_hi;
    unsigned long long T;                       // k-mers per mid bin (target)
    const unsigned long long* bin_kmer;         // [B] k-mers per bin (the scan's histogram)
    const unsigned long long* mid_first;        // [B+1] first mid bin of every bin (relative to this launch: mid_first[bin_lo] = 0)
    uint32_t* cell2mid;                         // [(B << cell_bits) - (bin_lo << cell_bits)] mid bin of every cell of the phase
    unsigned long long* mid_rec;                // [n_mid] records per mid bin (zeroed)
    unsigned long long* mid_kmer;               // [n_mid]
    uint32_t* mid_bin;                          // [n_mid]
};
// one warp per bin: mid bin of a cell = (k-mers of the bin's cells before it) / T
__global__ void __launch_bounds__(256) k_cells_assign(const CellsParams P) {
    const int lane = threadIdx.x & 31;
    const int b = P.bin_lo + (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= P.bin_hi) return;
    const unsigned long long m0 = P.mid_first[b], m1 = P.mid_first[b + 1];
    for (unsigned long long mm = m0 + lane; mm < m1; mm += 32) P.mid_bin[mm] = (uint32_t)b;
    const uint32_t C = 1u << P.cell_bits;
    const unsigned long long Tb = max(P.T, (P.bin_kmer[b] + C - 1) / C);      // never more mid bins than cells (the host plans with the same rule)
    unsigned long long carry = 0;
    for (uint32_t c0 = 0; c0 < C; c0 += 32) {
        const uint32_t c = c0 + lane;
        const size_t cell = ((size_t)b << P.cell_bits) + c;
        unsigned long long km = 0, rc = 0;
        if (c < C) { km = P.cell_kmer[cell]; rc = P.cell_rec[cell]; }
        unsigned long long incl = km;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        const unsigned long long excl = carry + incl - km;
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (c < C) {
            uint32_t mid = kNoMid;
            if (rc) {
                const unsigned long long mm = m0 + excl / Tb;             // < m1 because excl < k-mers of the bin
                mid = (uint32_t)mm;
                atomicAdd(&P.mid_rec[mm], rc); atomicAdd(&P.mid_kmer[mm], km);
            }
            P.cell2mid[cell - ((size_t)P.bin_lo << P.cell_bits)] = mid;
        }
    }
}

// exclusive scan of n u64 values by ONE CTA of 1024 threads (8 values per thread per round); out[n] = total
__global__ void __launch_bounds__(1024) k_excl_scan_u64(const unsigned long long* in, unsigned long long* out, unsigned long long n) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (unsigned long long base = 0; base < n; base += 8192) {
        const unsigned long long i0 = base + (unsigned long long)threadIdx.x * 8ull;
        unsigned long long v[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { v[j] = (i0 + j < n) ? in[i0 + j] : 0ull; sum += v[j]; }
        unsigned long long incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        unsigned long long wb = 0, tot = 0;
        for (int j = 0; j < 32; j++) { const unsigned long long t = s_w[j]; if (j < warp) wb += t; tot += t; }
        unsigned long long run = s_carry + wb + incl - sum;
#pragma unroll
        for (int j = 0; j < 8; j++) { if (i0 + j < n) out[i0 + j] = run; run += v[j]; }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}

// ------------------------------------------------------------------ run events -> records, mid-bin-major
struct Scatter2Params {
    const ulonglong2* events; unsigned long long n_events;     // {first window end, n | h16 << 16 | bin << 32}
    const uint64_t* bases; uint64_t n_words;
    int cap; int k; int cell_bits;
    unsigned long long cell_lo, cell_hi;                       // cells of this phase; events of other cells are skipped
    const uint32_t* cell2mid;                                  // [cell - cell_lo] mid bin of the cell
    unsigned long long* mid_next;                              // [n_mid] next record slot of the mid bin (starts at its record offset): ONE atomic per
                                                               // event, and few enough write tails (one per mid bin) for L2 to merge the 16-byte records
    void* records;
};
template <bool WIDE>
__global__ void __launch_bounds__(256) k_scatter2(const Scatter2Params P) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_events) return;
    const ulonglong2 ev = __ldcs(P.events + i);
    const unsigned long long rs = ev.x;
    const uint32_t n = (uint32_t)(ev.y & 0xFFFFull), h16 = (uint32_t)(ev.y >> 16) & 0xFFFFu, bin = (uint32_t)(ev.y >> 32);
    constexpr int NW = WIDE ? 5 : 3;
    const unsigned long long a0 = rs - (unsigned long long)(P.k - 1);
    const unsigned long long j0 = a0 >> 5; const uint32_t sh = 2u * (uint32_t)(a0 & 31ull);
    uint64_t w[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) w[q] = (j0 + q < P.n_words) ? P.bases[j0 + q] : 0ull;
    const unsigned long long cell = ((unsigned long long)bin << P.cell_bits) | (unsigned long long)(h16 >> (16 - P.cell_bits));
    if (cell < P.cell_lo || cell >= P.cell_hi) return;
    const uint32_t pieces = n <= (uint32_t)P.cap ? 1u : (n + (uint32_t)P.cap - 1) / (uint32_t)P.cap;
    const unsigned long long slot0 = atomicAdd(&P.mid_next[P.cell2mid[cell - P.cell_lo]], (unsigned long long)pieces);
    {
        const uint32_t nn = min((uint32_t)P.cap, n);
        uint64_t r[NW - 1];
#pragma unroll
        for (int q = 0; q < NW - 1; q++) r[q] = sh ? ((w[q] << sh) | (w[q + 1] >> (64 - sh))) : w[q];
        r[NW - 2] = (r[NW - 2] & ~0xFFull) | (uint64_t)nn;
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(P.records) + (WIDE ? 2 : 1) * slot0;
        dst[0] = make_ulonglong2(r[0], r[1]);
        if constexpr (WIDE) dst[1] = make_ulonglong2(r[2], r[3]);
    }
    for (uint32_t pc = 1, off = (uint32_t)P.cap; off < n; off += (uint32_t)P.cap, pc++) {
        const uint32_t nn = min((uint32_t)P.cap, n - off);
        write_record<WIDE>(P.records, slot0 + pc, P.bases, P.n_words, rs + off - (unsigned long long)(P.k - 1), nn);
    }
}

// ------------------------------------------------------------------ bulk copy + mbarrier (PTX)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// global -> shared bulk copy by the copy engine (TMA, 1-D form); bytes is a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ key128 atoms_cas128(key128* addr, key128 cmp, key128 val) {
    key128 old;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\t"
                 "mov.b128 c, {%2, %3};\n\t"
                 "mov.b128 v, {%4, %5};\n\t"
                 "atom.shared.cas.b128 o, [%6], c, v;\n\t"
                 "mov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old.lo), "=l"(old.hi)
                 : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "r"(smem_addr(addr)) : "memory");
    return old;
}

// Canonical k-mer number j of a record in memory, cut out with 32-bit funnel shifts.  A record is 2 (4) little-endian u64
// holding the bases MSB-first, so the i-th 32-bit word of the base string sits at word index i ^ 1; the record's last byte
// is its k-mer count and never belongs to a k-mer (j + k <= 60 resp. 124 bases).
struct RecShared {                                           // record words in shared memory, by 32-bit shared address
    uint32_t base;
    __device__ __forceinline__ uint32_t word(uint32_t byte_off) const {
        uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + byte_off)); return v;
    }
};
struct RecGlobal {                                           // ... in global memory
    const unsigned char* base;
    __device__ __forceinline__ uint32_t word(uint32_t byte_off) const { return *reinterpret_cast<const uint32_t*>(base + byte_off); }
};
template <typename R>
__device__ __forceinline__ uint64_t kmer_of_record(const R& recs, uint32_t rec_byte, uint32_t j, int k) {
    const uint32_t bit = 2u * j, q = bit >> 5, sh = bit & 31u;             // q <= 2
    const uint32_t o = rec_byte + 4u * q;                                   // word i of the base string is at word index i ^ 1
    const uint32_t A = recs.word(o ^ 4u), B = recs.word((o + 4u) ^ 4u), C = (q + 2u <= 3u) ? recs.word((o + 8u) ^ 4u) : 0u;
    const uint32_t hi = __funnelshift_l(B, A, sh), lo = __funnelshift_l(C, B, sh);
    const uint64_t fwd = (((uint64_t)hi << 32) | lo) >> (64 - 2 * k);
    const uint64_t rc = revcomp64(fwd, k);
    return fwd < rc ? fwd : rc;
}
template <typename R>
__device__ __forceinline__ key128 kmer_of_record_wide(const R& recs, uint32_t rec_byte, uint32_t j, int k) {
    const uint32_t bit = 2u * j, q = bit >> 5, sh = bit & 31u;             // q <= 5: j <= 124 - k <= 91
    const uint32_t o = rec_byte + 4u * q;
    uint32_t W[5];
#pragma unroll
    for (uint32_t d = 0; d < 5; d++) W[d] = (q + d <= 7u) ? recs.word((o + 4u * d) ^ 4u) : 0u;
    const uint32_t x0 = __funnelshift_l(W[1], W[0], sh), x1 = __funnelshift_l(W[2], W[1], sh), x2 = __funnelshift_l(W[3], W[2], sh), x3 = __funnelshift_l(W[4], W[3], sh);
    const uint64_t h0 = ((uint64_t)x0 << 32) | x1, h1 = ((uint64_t)x2 << 32) | x3;
    const int s = 128 - 2 * k;                                              // 0..62
    key128 fwd;
    fwd.hi = h0 >> s; fwd.lo = s ? ((h0 << (64 - s)) | (h1 >> s)) : h1;
    const key128 rc = revcomp128(fwd, k);
    return key_less(rc, fwd) ? rc : fwd;
}

// ------------------------------------------------------------------ the count kernel
struct SmemCountParams {
    const void* records;                        // mid-bin-major
    const unsigned long long* mid_rec_base;     // [n_mid+1] record offsets
    const unsigned long long* mid_kmer_base;    // [n_mid+1] k-mer offsets (the CTAs split the mid bins by k-mers)
    const uint32_t* mid_bin;                    // [n_mid] bin of every mid bin
    const unsigned long long* mid_first;        // [B+1] first mid bin of every bin
    uint32_t n_mid;
    void* out_keys; uint32_t* out_cnt;          // CTA c writes its entries densely from index c * region_cap
    unsigned long long region_cap;
    unsigned long long* cta_total;              // [gridDim.x] entries every CTA wrote
    uint32_t* bin_cta; unsigned long long* bin_off;   // [B] where a bin's entries begin: CTA and offset inside its region
    unsigned long long* acc;                    // [3][64] digest accumulators (sum, xor, count)
    int k;
    uint32_t cap_slots;                         // slots of the shared-memory table
    uint32_t max_fill;                          // distinct k-mers it may take (also the length of the claim list)
    uint32_t stage_recs;                        // records per staging buffer
    void* slow_keys; uint32_t* slow_cnt;        // private global table of every CTA: keys / counts of CTA c at c * slow_slots (filled on first use)
    unsigned long long slow_slots; unsigned long long slow_max_fill;
    int* flags;                                 // [0] slow table overflow (job must be redone), [1] output region too small, [2] a 32-bit count wrapped
    unsigned long long* counters;               // [0] mid bins that took the slow path
};

template <bool WIDE> struct SmTraits;
template <> struct SmTraits<false> { typedef uint64_t Key; static constexpr int kRecBytes = 16; };
template <> struct SmTraits<true> { typedef key128 Key; static constexpr int kRecBytes = 32; };

__device__ __forceinline__ bool sm_is_empty(uint64_t k) { return k == ~0ull; }
__device__ __forceinline__ bool sm_is_empty(key128 k) { return k.lo == ~0ull && k.hi == ~0ull; }
__device__ __forceinline__ uint64_t sm_empty(uint64_t) { return ~0ull; }
__device__ __forceinline__ key128 sm_empty(key128) { key128 e; e.lo = ~0ull; e.hi = ~0ull; return e; }

// Insert one k-mer per lane into a table in SHARED memory, warp-synchronously.  The table is a sequence of two-slot buckets
// probed linearly: every round a lane reads its bucket (both keys at once), looks at slot 0 then slot 1 — present: count
// it (ATOMS.ADD); empty: try to claim it (ATOMS.CAS; a lost race to the same key counts as present, to another key moves
// on) — and goes to the next bucket when both slots belong to other keys.  The rounds are straight-line code for the whole
// warp.  Returns the number of slots this lane claimed (0 or 1), or 0x10000 when the table was full (the caller keeps it
// below max_fill < cap, so that only happens with tiny test tables).
__device__ __forceinline__ unsigned sm_insert_block(uint64_t* keys, uint32_t* cnt, uint32_t cap, uint64_t key, bool act, int* wrapped) {
    const uint32_t nb = cap >> 1;                            // cap is even
    uint32_t b = umulhi32(key_hash(key), nb);
    unsigned claims = 0;
    for (uint32_t round = 0; round < nb && __any_sync(0xFFFFFFFFu, act); round++) {
        uint64_t k0 = 0, k1 = 0;
        if (act) {
            asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(k0), "=l"(k1) : "r"(smem_addr(&keys[2 * b])) : "memory");
        }
        int at = -1;                                         // slot the key was found or claimed at
        if (act) {
            if (k0 == key) at = 0;
            else {
                if (k0 == ~0ull) {
                    const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(&keys[2 * b]), ~0ull, (unsigned long long)key);
                    if (old == ~0ull) { at = 0; claims++; } else if (old == key) at = 0;
                }
                if (at < 0) {
                    if (k1 == key) at = 1;
                    else if (k1 == ~0ull) {
                        const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(&keys[2 * b + 1]), ~0ull, (unsigned long long)key);
                        if (old == ~0ull) { at = 1; claims++; } else if (old == key) at = 1;
                    }
                }
            }
            if (at >= 0) {
                if (atomicAdd(&cnt[2 * b + at], 1u) == 0xFFFFFFFFu) *wrapped = 1;
                act = false;
            }
        }
        b = (b + 1 == nb) ? 0 : b + 1;
    }
    return act ? 0x10000u : claims;
}
__device__ __forceinline__ unsigned sm_insert_block(key128* keys, uint32_t* cnt, uint32_t cap, key128 key, bool act, int* wrapped) {
    uint32_t slot = umulhi32(key_hash(key), cap);
    const key128 empty = sm_empty(key);
    unsigned claims = 0;
    for (uint32_t round = 0; round < cap && __any_sync(0xFFFFFFFFu, act); round++) {
        key128 cur = empty;
        if (act) {
            asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(cur.lo), "=l"(cur.hi) : "r"(smem_addr(&keys[slot])) : "memory");
        }
        // (a 16-byte shared-memory load of an aligned slot is not torn, but a half equal to all ones is handed to the CAS anyway)
        const bool emp = act && (cur.lo == ~0ull || cur.hi == ~0ull);
        if (emp) cur = atoms_cas128(&keys[slot], empty, key);
        const bool claimed = emp && sm_is_empty(cur);
        const bool hit = act && (claimed || key_eq(cur, key));
        if (hit && atomicAdd(&cnt[slot], 1u) == 0xFFFFFFFFu) *wrapped = 1;
        claims += claimed ? 1u : 0u;
        act = act && !hit;
        slot = (slot + 1 == cap) ? 0 : slot + 1;
    }
    return act ? 0x10000u : claims;
}
// The same in GLOBAL memory (slow path).  Returns 1 when the key was new, 0 when it was present, -1 when no slot was found.
__device__ __forceinline__ int gm_insert(uint64_t* keys, uint32_t* cnt, unsigned long long cap, uint64_t key, int* wrapped) {
    unsigned long long slot = slot_of(key_hash(key), cap);
    for (int probe = 0; probe < kSmMaxProbe; probe++) {
        uint64_t cur = __ldcg(&keys[slot]);
        int claimed = 0;
        if (cur == ~0ull) {
            cur = atomicCAS(reinterpret_cast<unsigned long long*>(&keys[slot]), ~0ull, (unsigned long long)key);
            if (cur == ~0ull) { claimed = 1; cur = key; }
        }
        if (cur == key) { if (atomicAdd(&cnt[slot], 1u) == 0xFFFFFFFFu) *wrapped = 1; return claimed; }
        if (++slot == cap) slot = 0;
    }
    return -1;
}
__device__ __forceinline__ int gm_insert(key128* keys, uint32_t* cnt, unsigned long long cap, key128 key, int* wrapped) {
    unsigned long long slot = slot_of(key_hash(key), cap);
    const key128 empty = sm_empty(key);
    for (int probe = 0; probe < kSmMaxProbe; probe++) {
        const ulonglong2 q = __ldcg(reinterpret_cast<const ulonglong2*>(&keys[slot]));
        key128 cur; cur.lo = q.x; cur.hi = q.y;
        int claimed = 0;
        if (cur.lo == ~0ull || cur.hi == ~0ull) {
            cur = cas128(&keys[slot], empty, key);
            if (sm_is_empty(cur)) { claimed = 1; cur = key; }
        }
        if (key_eq(cur, key)) { if (atomicAdd(&cnt[slot], 1u) == 0xFFFFFFFFu) *wrapped = 1; return claimed; }
        if (++slot == cap) slot = 0;
    }
    return -1;
}

// Walks the k-mers of `n_rec` records (shared or global memory) with the whole CTA: warp w takes records 32 w .. 32 w + 31,
// then 32 (w + kSmWarps) ..., the warp's k-mers form a pool and lane l takes k-mers l, l + 32, ... of it, so every lane is busy
// whatever the lengths of the records.  F(key, active) is called by the converged warp once per 32 k-mers of the pool.
template <bool WIDE, typename R, typename F>
__device__ __forceinline__ void for_each_kmer_cta(const R& recs, uint32_t n_rec, int k, F&& f) {
    typedef typename SmTraits<WIDE>::Key Key;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t RBY = WIDE ? 32u : 16u;              // bytes per record
    const uint32_t le_mask = (lane == 31) ? 0xFFFFFFFFu : ((2u << lane) - 1u);
    for (uint32_t r0 = (uint32_t)warp * 32u; r0 < n_rec; r0 += kSmWarps * 32u) {
        const uint32_t r = r0 + (uint32_t)lane;
        uint32_t nk = 0;
        if (r < n_rec) nk = recs.word(r * RBY + (RBY - 8u)) & 0xFFu;                             // low byte of the last u64
        uint32_t incl = nk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
        const uint32_t excl = incl - nk;
        const uint32_t T = __shfl_sync(FULL, incl, 31);
        uint32_t n_before = 0;                              // records that start before the current block of 32 k-mers
        for (uint32_t t0 = 0; t0 < T; t0 += 32) {
            const uint32_t rel = excl - t0;
            const uint32_t M = __reduce_or_sync(FULL, (nk && rel < 32u) ? (1u << rel) : 0u);    // record starts inside the block
            const uint32_t t = t0 + (uint32_t)lane;
            const int ri = min(31, max(0, (int)n_before + __popc(M & le_mask) - 1));            // lane that holds this k-mer's record
            const uint32_t j = t - __shfl_sync(FULL, excl, ri);
            n_before += __popc(M);
            const bool active = t < T;
            Key key = Key();
            if (active) {
                if constexpr (!WIDE) key = kmer_of_record(recs, (r0 + (uint32_t)ri) * RBY, j, k); else key = kmer_of_record_wide(recs, (r0 + (uint32_t)ri) * RBY, j, k);
            }
            f(key, active);
        }
    }
}

static constexpr int kSmStages = 3;                          // staging buffers of k_count_smem

// dynamic shared memory: keys[cap_slots] | cnt[cap_slots] | stage[kSmStages][stage_recs]
//
// Every CTA owns a contiguous range of mid bins (equal shares of the k-mers) and therefore a contiguous range of records
// and a contiguous region of the output: no CTA ever waits for another.  The records stream through kSmStages staging
// buffers filled by the bulk-copy engine (thread 0 issues the copies kSmStages - 1 chunks ahead); the 32
// warps insert a mid bin's k-mers into the table, and when the mid bin ends its distinct (k-mer, count) pairs are
// appended to the CTA's output region and the table is emptied again.
template <bool WIDE>
__global__ void __launch_bounds__(kSmBlock, 1) k_count_smem(const SmemCountParams P) {
    typedef typename SmTraits<WIDE>::Key Key;
    constexpr int RB = SmTraits<WIDE>::kRecBytes;
    extern __shared__ __align__(128) unsigned char sm_raw[];
    Key* keys = reinterpret_cast<Key*>(sm_raw);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(keys + P.cap_slots);
    unsigned char* stage = reinterpret_cast<unsigned char*>(cnt + P.cap_slots);          // cap_slots is a multiple of 32: 128-byte aligned
    __shared__ __align__(8) uint64_t s_full[kSmStages];
    __shared__ unsigned int s_nclaim[2], s_pos;
    __shared__ unsigned int s_range[2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool consumer = true;
    const bool producer = threadIdx.x == 0;                  // issues the bulk copies (a few instructions per chunk)
    const Key empty = sm_empty(Key());
    for (uint32_t i = threadIdx.x; i < P.cap_slots; i += kSmBlock) { keys[i] = empty; cnt[i] = 0; }
    if (producer) {
        for (int b = 0; b < kSmStages; b++) mbar_init(&s_full[b], 1);
        mbar_fence_init();
        s_nclaim[0] = 0; s_nclaim[1] = 0; s_pos = 0;
        // this CTA's mid bins: [first mid bin whose k-mer offset >= c * total / G, the same for c + 1)
        const unsigned long long total = P.mid_kmer_base[P.n_mid];
        for (int e = 0; e < 2; e++) {
            const unsigned long long want = (total / gridDim.x) * (blockIdx.x + e) + min((unsigned long long)(blockIdx.x + e), total % gridDim.x);
            uint32_t lo = 0, hi = P.n_mid;                    // first m with mid_kmer_base[m] >= want
            if (blockIdx.x + e >= gridDim.x) lo = P.n_mid;
            else while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (P.mid_kmer_base[mid] >= want) hi = mid; else lo = mid + 1; }
            s_range[e] = lo;
        }
    }
    __syncthreads();
    const uint32_t m_lo = s_range[0], m_hi = s_range[1];
    const unsigned long long R_lo = P.mid_rec_base[m_lo], R_hi = P.mid_rec_base[m_hi];
    const unsigned long long n_chunks = (R_hi - R_lo + P.stage_recs - 1) / P.stage_recs;
    auto issue = [&](unsigned long long i) {                 // chunk i -> buffer i % kSmStages
        const unsigned long long lo = R_lo + i * P.stage_recs;
        const uint32_t n = (uint32_t)min((unsigned long long)P.stage_recs, R_hi - lo);
        const int b = (int)(i % kSmStages);
        mbar_expect_tx(&s_full[b], n * RB);
        bulk_g2s(stage + (size_t)b * P.stage_recs * RB, reinterpret_cast<const unsigned char*>(P.records) + lo * RB, n * RB, &s_full[b]);
    };
    if (producer) for (unsigned long long i = 0; i < n_chunks && i < kSmStages - 1; i++) issue(i);

    Key* const okeys = reinterpret_cast<Key*>(P.out_keys) + (size_t)blockIdx.x * P.region_cap;
    uint32_t* const ocnt = P.out_cnt + (size_t)blockIdx.x * P.region_cap;
    unsigned long long off = 0;                              // entries this CTA has written (uniform)
    unsigned long long dsum = 0, dxor = 0, dcnt = 0;
    bool slow = false;                                       // the current mid bin overflowed the shared-memory table (uniform)
    bool slow_ready = false;                                 // this CTA's global table has been initialised
    bool out_full = false;
    int wrapped = 0;
    uint32_t m = m_lo;                                       // current mid bin
    unsigned long long mid_end = m < m_hi ? P.mid_rec_base[m + 1] : 0;
    int par = 0;                                             // claim counter of the current mid bin: s_nclaim[par]

    // the mid bin m is complete: its distinct k-mers leave, the table is emptied
    auto finish_mid = [&]() {
        const uint32_t bin = P.mid_bin[m];
        const bool had_records = P.mid_rec_base[m + 1] > P.mid_rec_base[m];
        if (producer && P.mid_first[bin] == (unsigned long long)m) { P.bin_cta[bin] = blockIdx.x; P.bin_off[bin] = off; }
        if (!had_records) return;
        __syncthreads();                                     // all inserts of the mid bin are done
        if (!slow && s_nclaim[par] > P.max_fill) slow = true;
        Key* gkeys = nullptr; uint32_t* gcnt = nullptr;
        if (slow) {
            // too many distinct k-mers for the shared-memory table: redo the mid bin in this CTA's private global table,
            // reading the records straight from global memory
            for (uint32_t i = threadIdx.x; i < P.cap_slots; i += kSmBlock) { keys[i] = empty; cnt[i] = 0; }
            gkeys = reinterpret_cast<Key*>(P.slow_keys) + (size_t)blockIdx.x * P.slow_slots;
            gcnt = P.slow_cnt + (size_t)blockIdx.x * P.slow_slots;
            if (!slow_ready) {                               // first use by this CTA: empty table (the dump below leaves it empty again)
                for (unsigned long long i = threadIdx.x; i < P.slow_slots; i += kSmBlock) { gkeys[i] = empty; gcnt[i] = 0; }
                slow_ready = true;
            }
            __syncthreads();
            if (producer) { s_nclaim[par] = 0; atomicAdd(&P.counters[0], 1ull); }
            __syncthreads();
            const unsigned long long r_lo = P.mid_rec_base[m], r_hi = P.mid_rec_base[m + 1];
            for (unsigned long long c0 = r_lo; c0 < r_hi; c0 += 4096ull) {
                const uint32_t nn = (uint32_t)min(4096ull, r_hi - c0);
                const bool full = s_nclaim[par] > P.slow_max_fill;                   // uniform: read between two barriers
                if (consumer) {
                    unsigned int claims = 0, failed = 0;
                    if (!full)
                        for_each_kmer_cta<WIDE>(RecGlobal{reinterpret_cast<const unsigned char*>(P.records) + c0 * RB}, nn, P.k,
                                                [&](Key key, bool active) {
                                                    if (!active) return;
                                                    const int c = gm_insert(gkeys, gcnt, P.slow_slots, key, &wrapped);
                                                    if (c > 0) claims++; else if (c < 0) failed = 1;
                                                });
                    claims = __reduce_add_sync(0xFFFFFFFFu, claims);
                    failed = __reduce_or_sync(0xFFFFFFFFu, failed);
                    if (lane == 0 && claims) atomicAdd(&s_nclaim[par], claims);
                    if (lane == 0 && failed) P.flags[0] = 1;
                }
                __syncthreads();
            }
            if (producer && s_nclaim[par] > P.slow_max_fill) P.flags[0] = 1;         // even the global table is too small: the job is redone elsewhere
        }
        const unsigned int D = s_nclaim[par];
        if (producer) s_nclaim[par ^ 1] = 0;                 // the next mid bin's counter (nobody touches it before the barrier below)
        const bool skip = off + D > P.region_cap;
        if (skip) out_full = true;
        // pass 1: the occupied slots leave in warp-compacted order and are emptied
        auto emit = [&](unsigned long long o, Key kk, uint32_t n) { if (!skip) { okeys[o] = kk; ocnt[o] = n; } };
        if (!slow) {
            for (uint32_t s0 = 0; s0 < P.cap_slots; s0 += kSmThreads) {
                const uint32_t s = s0 + threadIdx.x;
                Key kk = empty; uint32_t n = 0;
                if (s < P.cap_slots) kk = keys[s];
                const bool occ = !sm_is_empty(kk);
                if (occ) { n = cnt[s]; keys[s] = empty; cnt[s] = 0; }
                const uint32_t mk = __ballot_sync(0xFFFFFFFFu, occ);
                if (mk == 0u) continue;
                unsigned int wb = 0;
                if (lane == 0) wb = atomicAdd(&s_pos, (unsigned)__popc(mk));
                wb = __shfl_sync(0xFFFFFFFFu, wb, 0);
                if (occ) emit(off + wb + __popc(mk & ((1u << lane) - 1u)), kk, n);
            }
        } else {
            for (unsigned long long s0 = 0; s0 < P.slow_slots; s0 += kSmThreads) {
                const unsigned long long s = s0 + threadIdx.x;
                Key kk = empty; uint32_t n = 0;
                if (s < P.slow_slots) {
                    if constexpr (!WIDE) kk = __ldcg(&gkeys[s]);
                    else { const ulonglong2 q = __ldcg(reinterpret_cast<const ulonglong2*>(&gkeys[s])); kk.lo = q.x; kk.hi = q.y; }
                }
                const bool occ = !sm_is_empty(kk);
                if (occ) { n = __ldcg(&gcnt[s]); gkeys[s] = empty; gcnt[s] = 0; }
                const uint32_t mk = __ballot_sync(0xFFFFFFFFu, occ);
                unsigned int wb = 0;
                if (lane == 0 && mk) wb = atomicAdd(&s_pos, (unsigned)__popc(mk));
                wb = __shfl_sync(0xFFFFFFFFu, wb, 0);
                if (occ) {
                    const unsigned int o = wb + __popc(mk & ((1u << lane) - 1u));
                    if (o < D) emit(off + o, kk, n);         // (o >= D only after a slow-table overflow, when the job is redone anyway)
                }
            }
        }
        __syncthreads();                                     // the mid bin's entries are in the output region (visible to the whole CTA)
        // pass 2: digest over the dense entries (all lanes busy; the lines were just written and sit in L2)
        if (!skip) {
            const uint64_t hbin = mix64((uint64_t)bin);
            const uint64_t hpre = mix64(hbin);               // entry_hash's inner term when hi == 0 (64-bit keys)
            for (unsigned int i = threadIdx.x; i < D; i += kSmThreads) {
                Key kk; const uint32_t n = __ldcg(&ocnt[off + i]);
                uint64_t h;                                  // == entry_hash(bin, hi, lo)
                if constexpr (!WIDE) { kk = __ldcg(&okeys[off + i]); h = mix64(kk ^ hpre); }
                else { const ulonglong2 q = __ldcg(reinterpret_cast<const ulonglong2*>(&okeys[off + i])); kk.lo = q.x; kk.hi = q.y; h = mix64(kk.lo ^ mix64(kk.hi ^ hbin)); }
                dsum += h * (uint64_t)n; dxor ^= mix64(h + n); dcnt += n;
            }
        }
        if (!skip) off += D;
        // (the barrier above also separates this mid bin's table accesses from the next one's inserts)
        if (producer) s_pos = 0;
        slow = false;
        par ^= 1;
    };

    uint32_t phase[kSmStages];
#pragma unroll
    for (int b = 0; b < kSmStages; b++) phase[b] = 0u;
    for (unsigned long long ci = 0; ci < n_chunks; ci++) {
        const int buf = (int)(ci % kSmStages);
        if (producer && ci + kSmStages - 1 < n_chunks) issue(ci + kSmStages - 1);    // its buffer was released by the barrier that ended chunk ci - 1
        const unsigned long long c_lo = R_lo + ci * P.stage_recs, c_hi = min(c_lo + (unsigned long long)P.stage_recs, R_hi);
        if (consumer) {
            uint32_t ph = 0;
#pragma unroll
            for (int b = 0; b < kSmStages; b++) if (b == buf) { ph = phase[b]; phase[b] ^= 1u; }
            while (!mbar_try_wait(&s_full[buf], ph)) {}
        }
        const unsigned char* cbase = stage + (size_t)buf * P.stage_recs * RB;
        unsigned long long pos = c_lo;
        while (pos < c_hi) {
            const unsigned long long seg_hi = min(c_hi, mid_end);
            if (seg_hi > pos && consumer && !slow) {
                const RecShared recs{smem_addr(cbase) + (uint32_t)(pos - c_lo) * RB};
                unsigned int* const nclaim = &s_nclaim[par];
                for_each_kmer_cta<WIDE>(recs, (uint32_t)(seg_hi - pos), P.k, [&](Key key, bool active) {
                    // stop claiming when the table is as full as it may get (at most one insert per thread is in flight
                    // beyond max_fill, and cap_slots - max_fill > kSmThreads): the mid bin then takes the slow path
                    if (*reinterpret_cast<volatile unsigned int*>(nclaim) > P.max_fill) return;          // (warp-uniform: one shared word)
                    unsigned claims = sm_insert_block(keys, cnt, P.cap_slots, key, active, &wrapped);
                    claims = __reduce_add_sync(0xFFFFFFFFu, claims);
                    if (claims >= 0x10000u) claims = P.max_fill + 1u;                                     // table full: the mid bin takes the slow path
                    if (lane == 0 && claims) atomicAdd(nclaim, claims);
                });
            }
            if (seg_hi > pos) pos = seg_hi;
            while (m < m_hi && pos == mid_end) {             // the mid bin ends here (and so do the empty ones that follow it)
                finish_mid();
                m++;
                mid_end = m < m_hi ? P.mid_rec_base[m + 1] : ~0ull;
            }
        }
        __syncthreads();                                     // chunk consumed: its buffer may be refilled
    }
    while (m < m_hi) { finish_mid(); m++; }                  // mid bins without records at the end of the range (bookkeeping only)
    if (producer) P.cta_total[blockIdx.x] = off;
    if (out_full && threadIdx.x == 0) P.flags[1] = 1;
    if (wrapped) P.flags[2] = 1;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        dsum += __shfl_xor_sync(0xFFFFFFFFu, dsum, o); dxor ^= __shfl_xor_sync(0xFFFFFFFFu, dxor, o); dcnt += __shfl_xor_sync(0xFFFFFFFFu, dcnt, o);
    }
    if (lane == 0 && dcnt) {
        const int a = (blockIdx.x * kSmWarps + warp) & 63;
        atomicAdd(&P.acc[a], dsum); atomicXor(&P.acc[64 + a], dxor); atomicAdd(&P.acc[128 + a], dcnt);
    }
}

}  // namespace fkm
