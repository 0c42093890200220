// fkm_kernels.cuh — hand-written sm_100a kernels of the k-mer counting path.
//
// Stage map (SURVEY.md §2 "kernel inventory"; reference sites in brackets):
//   k_scan<MODE>      K1+K2+K4  sliding window, signature, bin, super-k-mer records
//                               [SBKC:75-157, UTIL:46-100,310-357,686-695]
//   k_count_ht        K5+K6a    canonical k-mers -> open-addressing tables   [SBKC:694-705]
//   k_compact_ht      K6a       tables -> dense per-bin (k-mer,count)        [SBKC:723-730]
//   k_expand          K5        canonical k-mers -> key array (sort path)    [SBKC:484-524]
//   k_radix_*         K6b       segmented LSD radix sort                     [SBKC:540-542]
//   k_rle_*           K6b       run-length count of sorted keys              [SBKC:566-597]
//   k_digest          K7        order-independent digest of the result
//   k_synth           —         synthetic reads of SURVEY §8(d), packed, on device
//
// Layouts.  Input: `bases` 32 positions per u64, first position in the two MSBs;
// `inv` 32 positions per u32, first position in the MSB.  All records of the
// input are concatenated with ONE invalid separator position after each record,
// so short reads and long sequences are the same problem: count every k-window
// that holds no invalid position.
//
// Super-k-mer record (the "shuffle" payload).  NARROW (k <= 32): 16 bytes, two
// u64 {w0,w1}: 60 bases MSB-first in w0[63:0],w1[63:8]; w1[7:0] = number of
// k-mers n (1..61-k).  WIDE (32 < k <= 64): 32 bytes, four u64: 124 bases, n in
// w3[7:0] (1..125-k).  A record holds n consecutive k-windows that share one
// signature, i.e. n+k-1 bases.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "fkm_common.h"
#include "fkm_math.h"

namespace fkm {

static constexpr uint32_t kInvalidMin = 0xFFFFFFFFu;
static constexpr int kScanThreads = 256;
static constexpr int kSmemHistMaxB = 4096;

struct __align__(16) SlotN { uint64_t key; uint32_t cnt; uint32_t pad; };
struct __align__(32) SlotW { key128 key; uint32_t cnt; uint32_t pad[3]; };

template <bool WIDE> struct Traits;
template <> struct Traits<false> {
    typedef uint64_t Key; typedef SlotN Slot;
    static constexpr int kRecWords = 2, kRecBases = 60;
};
template <> struct Traits<true> {
    typedef key128 Key; typedef SlotW Slot;
    static constexpr int kRecWords = 4, kRecBases = 124;
};

// ------------------------------------------------------------------ small helpers (the pure arithmetic ones live in fkm_math.h)
// 64 bits starting at bit position `pos` of an MSB-first u32 bit array
__device__ __forceinline__ uint64_t bits64_at(const uint32_t* a, uint32_t pos) {
    uint32_t wi = pos >> 5, sh = pos & 31;
    uint32_t w0 = a[wi], w1 = a[wi + 1], w2 = a[wi + 2];
    uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
    return ((uint64_t)hi << 32) | lo;
}
// 64 bits (32 bases) starting at base position `pos` of an MSB-first u64 base array
__device__ __forceinline__ uint64_t bases64_at(const uint64_t* a, uint32_t pos) {
    uint32_t wi = pos >> 5, sh = 2 * (pos & 31);
    uint64_t w0 = a[wi], w1 = a[wi + 1];
    return sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;
}

// ------------------------------------------------------------------ K1: scan
struct ScanParams {
    const uint64_t* bases; const uint32_t* inv;
    uint64_t n_pos;                 // positions in the input
    uint64_t n_words;               // u64 words in bases (= u32 words in inv)
    uint64_t e_total;               // window-end positions to visit
    uint64_t n_seg; uint32_t seg_len;   // segments of seg_len window ends (multiple of 32), one warp each
    int k, m, w;                    // w = k-m+1 m-mers per window
    int sh1[6];                     // shifts of the doubling levels of the signature window: w = 1 + sum(sh1), 0 = unused level
    uint32_t B;                     // the configuration's bins (hash_to_bucket)
    uint32_t Bi; int split;         // internal bins = B << split (fkm_common.h split_bin): what the histograms and the scatter index
    int cap;                        // max k-mers per record
    int wide;                       // record format (MODE 1 only)
    int smem_hist;                  // 1: per-CTA histogram in shared memory (B <= 4096)
    unsigned long long* hist_rec;   // [B] records per bin            (MODE 0)
    unsigned long long* hist_kmer;  // [B] k-mers per bin             (MODE 0)
    ulonglong2* events;             // run events {first window end, n | value<<32}   (MODE 0, !DUAL)
    unsigned long long ev_cap; unsigned long long* ev_count; int* ev_overflow;
    const unsigned long long* bin_base;   // [B+1] record offsets     (MODE 1)
    unsigned long long* cursor;     // [B << cursor_shift], one per bin at stride 1 << cursor_shift   (MODE 1)
    int cursor_shift;
    void* records;                  //                                (MODE 1)
    int32_t* dbg_bins;              // [n_pos]                        (MODE 2)
    // ---- DUAL (shared-memory count path): a second, hash-ordered minimizer of length m2 cuts the runs further and
    // spreads every bin over (1 << cell_bits) cells
    int m2; uint32_t mask2;
    int sh2[6];                     // doubling shifts of the second window: k-m2+1 = 1 + sum(sh2)
    int cell_bits;                  // cells per bin = 1 << cell_bits (<= 12)
    uint32_t sample_hi;             // list 1 holds every event, list 0 (also) those of the bins < sample_hi
    ulonglong2* events2[2];         // {first window end, n | h16 << 16 | bin << 32}
    unsigned long long ev_cap2[2];  // ev_count[0..1] are the two list lengths
    unsigned long long* cell_rec;   // [B << cell_bits] records per cell
    unsigned long long* cell_kmer;  // [B << cell_bits] k-mers per cell
};

// bases [a, a+n+k-1) -> one super-k-mer record at `slot`
template <bool WIDE>
__device__ __forceinline__ void write_record(void* records, unsigned long long slot, const uint64_t* bases, uint64_t n_words,
                                             unsigned long long a, uint32_t nn) {
    auto ld = [&](unsigned long long gw) -> uint64_t { return gw < n_words ? bases[gw] : 0ull; };
    const unsigned long long j = a >> 5; const uint32_t sh = 2u * (uint32_t)(a & 31ull);
    if constexpr (!WIDE) {
        uint64_t w0 = ld(j), w1 = ld(j + 1), w2 = ld(j + 2);
        uint64_t r0 = sh ? ((w0 << sh) | (w1 >> (64 - sh))) : w0;
        uint64_t r1 = sh ? ((w1 << sh) | (w2 >> (64 - sh))) : w1;
        r1 = (r1 & ~0xFFull) | (uint64_t)nn;
        reinterpret_cast<ulonglong2*>(records)[slot] = make_ulonglong2(r0, r1);
    } else {
        uint64_t x0 = ld(j), x1 = ld(j + 1), x2 = ld(j + 2), x3 = ld(j + 3), x4 = ld(j + 4);
        uint64_t r0 = sh ? ((x0 << sh) | (x1 >> (64 - sh))) : x0;
        uint64_t r1 = sh ? ((x1 << sh) | (x2 >> (64 - sh))) : x1;
        uint64_t r2 = sh ? ((x2 << sh) | (x3 >> (64 - sh))) : x2;
        uint64_t r3 = sh ? ((x3 << sh) | (x4 >> (64 - sh))) : x3;
        r3 = (r3 & ~0xFFull) | (uint64_t)nn;
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(records) + 2 * slot;
        dst[0] = make_ulonglong2(r0, r1);
        dst[1] = make_ulonglong2(r2, r3);
    }
}

// hash order of the second minimizer (any fixed bijection-like mix: only the PARTITION depends on it, never a count)
__device__ __forceinline__ uint32_t mmer2_hash(uint32_t v2, int m2) {
    const uint32_t r2 = revcomp32(v2, m2);
    return fmix32((v2 < r2 ? v2 : r2) + 0x9E3779B9u);
}
// 16 partition bits of a run from the minimum hash of its windows (the minimum itself is biased towards small values)
__device__ __forceinline__ uint32_t cell_hash16(uint32_t sig2) { return fmix32(sig2 ^ 0x5bd1e995u) >> 16; }

// MODE 0: bin histogram (records and k-mers per bin) + the list of run events.
// MODE 1: scatter records directly (recomputes the scan; fallback when the event list overflowed).
// MODE 2: per-window bin ids (test hook).
// DUAL (MODE 0 only): runs are cut where the signature OR the second minimizer changes; events carry the bin and 16 hash
// bits of the second minimizer, and the histogram is kept per (bin, cell) in global memory as well.
//
// Warp-striped sliding window.  A warp owns a segment of consecutive window-END
// positions and walks it 32 at a time, lane l <-> position e = 32 g + l:
//   * the m-mer ending at e is cut out of two broadcast 64-bit words (funnel shift),
//     its reverse complement comes from brev, norm() is the closed form of UTIL:46-100;
//   * the minimum over the window's w m-mers is built by doubling with warp shuffles:
//     x_{j+1}[e] = min(x_j[e], x_j[e - s_j]), one shuffle per level (the source lane sends its value
//     of the previous group when its reader wrapped below lane 0); shifts 1, 2, 4, ..., then the
//     remainder, w = 1 + sum of the shifts; NL (levels) is a template parameter, the shifts are
//     kernel parameters (a level with shift 0 is a no-op);
//   * a window is valid iff the last invalid position at or before e is >= k behind;
//   * run boundaries (signature value changes / validity changes) are found with one
//     shuffle and two ballots; every lane that sees a run END pushes (start, length,
//     value) into a per-warp queue in shared memory, and the queue is drained 32 events
//     at a time so that hashing, atomics and stores run with full lanes.
// Positions inside a segment are 32-bit offsets from its warm-up start.  Two warm-up
// groups before each segment rebuild the register state, so segments are independent;
// runs are cut at segment boundaries (the reference's own cutting rule, SBKC:102-136,
// is not observable: SURVEY §7).
template <int MODE, int NL, bool DUAL>
__global__ void __launch_bounds__(kScanThreads) k_scan(const ScanParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long q_rs[kScanThreads / 32][64];
    __shared__ uint32_t q_n[kScanThreads / 32][64];
    __shared__ uint32_t q_v[kScanThreads / 32][64];
    __shared__ uint32_t q_v2[DUAL ? kScanThreads / 32 : 1][64];
    uint32_t* s_hist_rec = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* s_hist_kmer = s_hist_rec + P.Bi;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int k = P.k, m = P.m;
    const uint32_t mmask = (1u << (2 * m)) - 1u;
    const bool upper = lane >= 16;                          // m-mer lies inside W (else it reaches into Wprev)
    const uint32_t fsh = (2u * (31u - (uint32_t)lane)) & 31u;

    if (MODE == 0 && !DUAL && P.smem_hist) {
        for (uint32_t b = threadIdx.x; b < P.Bi; b += kScanThreads) { s_hist_rec[b] = 0; s_hist_kmer[b] = 0; }
        __syncthreads();
    }

    auto load_bases = [&](unsigned long long gw) -> uint64_t { return gw < P.n_words ? P.bases[gw] : 0ull; };
    auto load_inv = [&](unsigned long long gw) -> uint32_t {
        if (gw >= P.n_words) return 0xFFFFFFFFu;
        uint32_t iv = P.inv[gw];
        const unsigned long long p0 = gw << 5;
        if (p0 + 32 > P.n_pos) iv |= (p0 >= P.n_pos) ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (uint32_t)(P.n_pos - p0));
        return iv;
    };
    // drain `cnt` (<= 32) queued events: lane i takes event i
    auto drain = [&](int cnt) {
        const bool have = lane < cnt;
        unsigned long long rs = 0; uint32_t n = 0, v = 0, v2 = 0;
        if (have) { rs = q_rs[warp][lane]; n = q_n[warp][lane]; v = q_v[warp][lane]; if (DUAL) v2 = q_v2[warp][lane]; }
        if (MODE == 0 && !DUAL) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.ev_count, (unsigned long long)cnt);
            base = __shfl_sync(FULL, base, 0);
            if (base + (unsigned long long)cnt <= P.ev_cap) {
                if (have) P.events[base + lane] = make_ulonglong2(rs, (unsigned long long)n | ((unsigned long long)v << 32));
            } else if (lane == 0) *P.ev_overflow = 1;
        }
        const uint32_t bin = have ? split_bin(v, P.B, P.split) : 0u;
        if (MODE == 0 && DUAL) {
            // every event goes to list 1 in lane order; the events of the sample bins (few) are appended to list 0 as well
            const uint32_t h16 = cell_hash16(v2);
            const ulonglong2 ev = make_ulonglong2(rs, (unsigned long long)n | ((unsigned long long)h16 << 16) | ((unsigned long long)bin << 32));
            unsigned long long b1 = 0;
            if (lane == 0) b1 = atomicAdd(&P.ev_count[1], (unsigned long long)cnt);
            b1 = __shfl_sync(FULL, b1, 0);
            if (b1 + (unsigned long long)cnt <= P.ev_cap2[1]) { if (have) P.events2[1][b1 + lane] = ev; }
            else if (lane == 0) *P.ev_overflow = 1;
            const bool to0 = have && bin < P.sample_hi;
            const uint32_t m0 = __ballot_sync(FULL, to0);
            if (m0) {
                unsigned long long b0 = 0;
                if (lane == 0) b0 = atomicAdd(&P.ev_count[0], (unsigned long long)__popc(m0));
                b0 = __shfl_sync(FULL, b0, 0);
                if (to0) {
                    const unsigned long long idx = b0 + __popc(m0 & lt_mask);
                    if (idx < P.ev_cap2[0]) P.events2[0][idx] = ev; else *P.ev_overflow = 1;
                }
            }
            if (have) {
                const uint32_t pieces = n <= (uint32_t)P.cap ? 1u : (n + (uint32_t)P.cap - 1) / (uint32_t)P.cap;
                const size_t cell = ((size_t)bin << P.cell_bits) | (size_t)(h16 >> (16 - P.cell_bits));
                atomicAdd(&P.cell_rec[cell], (unsigned long long)pieces);
                atomicAdd(&P.cell_kmer[cell], (unsigned long long)n);
            }
        }
        if (have && !(MODE == 0 && DUAL)) {
            if (MODE == 0) {
                const uint32_t pieces = (n + (uint32_t)P.cap - 1) / (uint32_t)P.cap;
                if (P.smem_hist) { atomicAdd(&s_hist_rec[bin], pieces); atomicAdd(&s_hist_kmer[bin], n); }
                else { atomicAdd(&P.hist_rec[bin], (unsigned long long)pieces); atomicAdd(&P.hist_kmer[bin], (unsigned long long)n); }
            } else if (MODE == 1) {
                for (uint32_t off = 0; off < n; off += (uint32_t)P.cap) {
                    const uint32_t nn = min((uint32_t)P.cap, n - off);
                    const unsigned long long slot = P.bin_base[bin] + atomicAdd(&P.cursor[(size_t)bin << P.cursor_shift], 1ull);
                    const unsigned long long a = rs + off - (unsigned long long)(k - 1);
                    if (P.wide) write_record<true>(P.records, slot, P.bases, P.n_words, a, nn);
                    else write_record<false>(P.records, slot, P.bases, P.n_words, a, nn);
                }
            }
        }
    };

    int qn = 0;                                              // events waiting in this warp's queue (warp-uniform)
    const unsigned long long n_warps = (unsigned long long)gridDim.x * (kScanThreads / 32);
    for (unsigned long long seg = (unsigned long long)blockIdx.x * (kScanThreads / 32) + warp; seg < P.n_seg; seg += n_warps) {
        const unsigned long long e0 = seg * P.seg_len;
        const unsigned long long e1 = min(e0 + (unsigned long long)P.seg_len, (unsigned long long)P.e_total);
        const unsigned long long g0 = e0 >> 5, g1 = (e1 + 31) >> 5;
        const unsigned long long gs = g0 >= 2 ? g0 - 2 : 0;   // two warm-up groups (>= k-1 positions)
        const unsigned long long ws = gs << 5;                // absolute position of relative position 0
        const int n_groups = (int)(g1 - gs), n_warm = (int)(g0 - gs);
        int carry_bad = -1;                                   // last invalid relative position seen so far
        uint32_t prev[NL], prev2[DUAL ? NL : 1];
#pragma unroll
        for (int j = 0; j < NL; j++) { prev[j] = kInvalidMin; if (DUAL) prev2[j] = kInvalidMin; }
        uint32_t prev_last = kInvalidMin, prev_last2 = 0;     // values of the window ending just before this group
        uint32_t run_start = 0;                               // relative END position of the first window of the open run
        uint64_t Wprev = gs > 0 ? load_bases(gs - 1) : 0ull;

        for (int gb = 0; gb < n_groups; gb += 32) {
            const uint64_t Wb = load_bases(gs + gb + lane);
            const uint32_t Ib = load_inv(gs + gb + lane);
            const int ng = min(32, n_groups - gb);
            for (int gi = 0; gi < ng; gi++) {
                const uint64_t W = __shfl_sync(FULL, Wb, gi);
                const uint32_t IW = __shfl_sync(FULL, Ib, gi);
                const int gr = gb + gi;                       // group index inside the segment
                const uint32_t e = ((uint32_t)gr << 5) + (uint32_t)lane;
                const bool real = gr >= n_warm;
                // ---- the 16 bases ending at e: low 32 bits of (Wprev:W) >> 2*(31-lane)
                const uint32_t lo = upper ? (uint32_t)W : (uint32_t)(W >> 32);
                const uint32_t hi = upper ? (uint32_t)(W >> 32) : (uint32_t)Wprev;
                const uint32_t u16b = __funnelshift_r(lo, hi, fsh);
                const uint32_t v = u16b & mmask;
                uint32_t x = mmer_norm(v, revcomp32(v, m), m, mmask);
                uint32_t y = 0;
                if (DUAL) y = mmer2_hash(u16b & P.mask2, P.m2);
                Wprev = W;
                // ---- minimum over the last w m-mers (doubling).  One shuffle per level: lane l reads from lane
                // (l - s) & 31, and that source lane knows what its reader needs — its value of the current group
                // if the reader is s lanes above it, of the previous group if the reader wrapped.
#pragma unroll
                for (int j = 0; j < NL; j++) {
                    const int s = P.sh1[j];
                    const uint32_t cur = x;
                    const uint32_t send = (lane < 32 - s) ? cur : prev[j];
                    x = min(cur, __shfl_sync(FULL, send, (lane - s) & 31));
                    prev[j] = cur;
                }
                if (DUAL) {
#pragma unroll
                    for (int j = 0; j < NL; j++) {
                        const int s = P.sh2[j];
                        const uint32_t cur = y;
                        const uint32_t send = (lane < 32 - s) ? cur : prev2[j];
                        y = min(cur, __shfl_sync(FULL, send, (lane - s) & 31));
                        prev2[j] = cur;
                    }
                }
                // ---- validity: last invalid position <= e must be at least k behind
                const uint32_t tb = IW >> (31 - lane);
                const int last_bad = tb ? (int)e - (__ffs((int)tb) - 1) : carry_bad;
                const bool valid = ((int)e - last_bad) >= k;
                carry_bad = IW ? (gr << 5) + 31 - (__ffs((int)IW) - 1) : carry_bad;
                const uint32_t Ev = valid ? x : kInvalidMin;
                const uint32_t Ev2 = (DUAL && valid) ? y : 0u;
                if (MODE == 2) {
                    const unsigned long long ea = ws + e;
                    if (real && ea + 1 >= (unsigned long long)k) {
                        const unsigned long long i = ea + 1 - (unsigned long long)k;
                        if (i < P.n_pos) P.dbg_bins[i] = valid ? (int32_t)hash_to_bucket(x, P.B) : -1;
                    }
                    continue;
                }
                // ---- run boundaries (shuffles and votes stay outside any branch)
                uint32_t Ep = __shfl_up_sync(FULL, Ev, 1), Ep2 = 0;
                if (DUAL) Ep2 = __shfl_up_sync(FULL, Ev2, 1);
                if (lane == 0) { Ep = prev_last; Ep2 = prev_last2; }
                const bool differs = (Ev != Ep) || (DUAL && Ev2 != Ep2);
                const bool is_start = real && (Ev != kInvalidMin) && differs;
                const bool is_end = real && (Ep != kInvalidMin) && differs;          // the run ending at e-1
                const uint32_t Sm = __ballot_sync(FULL, is_start), Em = __ballot_sync(FULL, is_end);
                const uint32_t last = __shfl_sync(FULL, Ev, 31);
                uint32_t last2 = 0;
                if (DUAL) last2 = __shfl_sync(FULL, Ev2, 31);
                if (Em) {
                    if (is_end) {
                        const uint32_t below = Sm & lt_mask;
                        const uint32_t rs = below ? ((uint32_t)gr << 5) + (uint32_t)(31 - __clz((int)below)) : run_start;
                        const int idx = qn + __popc(Em & lt_mask);
                        q_rs[warp][idx] = ws + rs; q_n[warp][idx] = e - rs; q_v[warp][idx] = Ep;
                        if (DUAL) q_v2[warp][idx] = Ep2;
                    }
                    qn += __popc(Em);
                    __syncwarp();
                    if (qn >= 32) {
                        drain(32);
                        __syncwarp();
                        const bool mv = lane < qn - 32;
                        unsigned long long t0 = 0; uint32_t t1 = 0, t2 = 0, t3 = 0;
                        if (mv) { t0 = q_rs[warp][32 + lane]; t1 = q_n[warp][32 + lane]; t2 = q_v[warp][32 + lane]; if (DUAL) t3 = q_v2[warp][32 + lane]; }
                        __syncwarp();
                        if (mv) { q_rs[warp][lane] = t0; q_n[warp][lane] = t1; q_v[warp][lane] = t2; if (DUAL) q_v2[warp][lane] = t3; }
                        qn -= 32;
                        __syncwarp();
                    }
                }
                if (Sm) run_start = ((uint32_t)gr << 5) + (uint32_t)(31 - __clz((int)Sm));
                if (real) { prev_last = last; prev_last2 = last2; }
            }
        }
        // the run still open at the end of the segment is cut here
        if (MODE != 2 && prev_last != kInvalidMin) {
            if (lane == 0) {
                q_rs[warp][qn] = ws + run_start; q_n[warp][qn] = ((uint32_t)n_groups << 5) - run_start; q_v[warp][qn] = prev_last;
                if (DUAL) q_v2[warp][qn] = prev_last2;
            }
            qn += 1;
            __syncwarp();
            if (qn >= 32) { drain(32); qn = 0; __syncwarp(); }     // qn == 32 exactly
        }
    }
    if (MODE != 2) {
        __syncwarp();
        if (qn > 0) drain(qn);
    }

    if (MODE == 0 && !DUAL && P.smem_hist) {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < P.Bi; b += kScanThreads) {
            uint32_t r = s_hist_rec[b];
            if (r) { atomicAdd(&P.hist_rec[b], (unsigned long long)r); atomicAdd(&P.hist_kmer[b], (unsigned long long)s_hist_kmer[b]); }
        }
    }
}

// K4: run events -> super-k-mer records, bin-major (the "shuffle").  One thread per event.
struct ScatterParams {
    const ulonglong2* events; unsigned long long n_events;
    const uint64_t* bases; uint64_t n_words;
    uint32_t B; int split; int cap; int k;          // bin of an event = split_bin(signature, B, split)
    const unsigned long long* bin_base; void* records;
    unsigned long long* cursor; int cursor_shift;   // write cursor of bin b at cursor[b << cursor_shift]
    int* overflow;                                  // speculative scatter (bin regions sized from a forecast): set when a bin outgrows its region; else NULL
};
template <bool WIDE>
__global__ void __launch_bounds__(256) k_scatter_events(const ScatterParams P) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_events) return;
    const ulonglong2 ev = __ldcs(P.events + i);
    const unsigned long long rs = ev.x;
    const uint32_t n = (uint32_t)ev.y, v = (uint32_t)(ev.y >> 32);
    // the bases of the first (usually only) piece are requested before the cursor atomic, so the two round trips overlap
    constexpr int NW = WIDE ? 5 : 3;
    const unsigned long long a0 = rs - (unsigned long long)(P.k - 1);
    const unsigned long long j0 = a0 >> 5; const uint32_t sh = 2u * (uint32_t)(a0 & 31ull);
    uint64_t w[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) w[q] = (j0 + q < P.n_words) ? P.bases[j0 + q] : 0ull;
    const uint32_t bin = split_bin(v, P.B, P.split);
    const uint32_t pieces = (n + (uint32_t)P.cap - 1) / (uint32_t)P.cap;
    const unsigned long long slot0 = P.bin_base[bin] + atomicAdd(&P.cursor[(size_t)bin << P.cursor_shift], (unsigned long long)pieces);
    if (P.overflow && slot0 + pieces > P.bin_base[bin + 1]) { *P.overflow = 1; return; }
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

// multi-GPU: records received source-major -> bin-major.  Segment i of the receive buffer
// ([seg_src[i], seg_src[i+1])) holds one (source rank, bin) pair and goes to seg_dst[i].
struct RegroupParams {
    const void* in; void* out; unsigned long long n; int rec_words;
    const unsigned long long* seg_src; const unsigned long long* seg_dst; int n_seg;
};
__global__ void __launch_bounds__(256) k_regroup(const RegroupParams P) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    int lo = 0, hi = P.n_seg;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (P.seg_src[mid] <= i) lo = mid; else hi = mid; }
    const unsigned long long o = P.seg_dst[lo] + (i - P.seg_src[lo]);
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(P.in); ulonglong2* dst = reinterpret_cast<ulonglong2*>(P.out);
    if (P.rec_words == 2) dst[o] = __ldcs(src + i);
    else { dst[2 * o] = __ldcs(src + 2 * i); dst[2 * o + 1] = __ldcs(src + 2 * i + 1); }
}

// ------------------------------------------------------------------ record -> k-mers
// Walks the n canonical k-mers of one record; F(key) is called once per k-mer.
template <typename F>
__device__ __forceinline__ void for_each_kmer_narrow(uint64_t w0, uint64_t w1, int k, F&& f) {
    const int n = (int)(w1 & 0xFFull);
    w1 &= ~0xFFull;
    const int s = 64 - 2 * k;
    uint64_t fwd = w0 >> s;
    uint64_t rc = revcomp64(fwd, k);
    for (int j = 0;;) {
        f(min(fwd, rc));
        if (++j == n) break;
        w0 = (w0 << 2) | (w1 >> 62); w1 <<= 2;
        fwd = w0 >> s;
        rc = (rc >> 2) | ((3ull - (fwd & 3ull)) << (2 * k - 2));
    }
}
template <typename F>
__device__ __forceinline__ void for_each_kmer_wide(uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3, int k, F&& f) {
    const int n = (int)(w3 & 0xFFull);
    w3 &= ~0xFFull;
    const int s = 128 - 2 * k;                 // 0..62
    key128 fwd;
    fwd.hi = w0 >> s; fwd.lo = s ? ((w0 << (64 - s)) | (w1 >> s)) : w1;
    key128 rc = revcomp128(fwd, k);
    for (int j = 0;;) {
        f(key_less(rc, fwd) ? rc : fwd);
        if (++j == n) break;
        w0 = (w0 << 2) | (w1 >> 62); w1 = (w1 << 2) | (w2 >> 62); w2 = (w2 << 2) | (w3 >> 62); w3 <<= 2;
        fwd.hi = w0 >> s; fwd.lo = s ? ((w0 << (64 - s)) | (w1 >> s)) : w1;
        uint64_t nb = 3ull - (fwd.lo & 3ull);
        rc.lo = (rc.lo >> 2) | (rc.hi << 62);
        rc.hi = (rc.hi >> 2) | (nb << (2 * k - 66));
    }
}

// largest b in [lo, hi) with base[b] <= r   (base ascending, base[lo] <= r < base[hi])
__device__ __forceinline__ int find_bin(const unsigned long long* base, int lo, int hi, unsigned long long r) {
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (base[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------ K6a: hash-table count
__device__ __forceinline__ key128 cas128(key128* addr, key128 cmp, key128 val) {
    key128 old;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\t"
                 "mov.b128 c, {%2, %3};\n\t"
                 "mov.b128 v, {%4, %5};\n\t"
                 "atom.global.cas.b128 o, [%6], c, v;\n\t"
                 "mov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old.lo), "=l"(old.hi)
                 : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr) : "memory");
    return old;
}

struct CountParams {
    const void* records;
    unsigned long long rec_lo, rec_hi;          // record range of this batch
    const unsigned long long* bin_base;         // [B+1] global record offsets
    int bin_lo, bin_hi;                         // bins of this batch
    void* table;                                // slots of this batch
    const unsigned long long* tbl_base;         // [bin_hi-bin_lo+1] slot offsets inside `table`
    unsigned long long* bin_distinct;           // [B] distinct k-mers per bin (claims)
    int* overflow;                              // set when a probe sequence exceeds max_probe
    int k; int max_probe;
    const uint32_t* weights;                    // [records] multiplicity of a folded record, or NULL (every record counts once)
    int first_state;                            // 0: read the slot, CAS only when it looks empty; 1: CAS straight away (every probe is ONE atomic at the slot's home L2 slice)
};

// returns 1 if this call claimed a new slot, 0 if the key was present, -1 on overflow
__device__ __forceinline__ int ht_insert(SlotN* tbl, unsigned long long size, uint64_t key, int max_probe, uint32_t weight = 1u) {
    unsigned long long slot = slot_of(key_hash(key), size);
    for (int probe = 0; probe < max_probe; probe++) {
        SlotN* s = tbl + slot;
        uint64_t cur = __ldcg(&s->key);
        int claimed = 0;
        if (cur == ~0ull) {
            cur = atomicCAS((unsigned long long*)&s->key, ~0ull, (unsigned long long)key);
            if (cur == ~0ull) { if (weight > 1u) atomicAdd(&s->cnt, weight - 1u); return 1; }   // claimed: the slot's count of 0 already means "seen once"
        }
        if (cur == key) { atomicAdd(&s->cnt, weight); return claimed; }
        if (++slot == size) slot = 0;
    }
    return -1;
}
__device__ __forceinline__ int ht_insert(SlotW* tbl, unsigned long long size, key128 key, int max_probe, uint32_t weight = 1u) {
    unsigned long long slot = slot_of(key_hash(key), size);
    const key128 empty = {~0ull, ~0ull};
    for (int probe = 0; probe < max_probe; probe++) {
        SlotW* s = tbl + slot;
        key128 cur;
        cur.lo = __ldcg(&s->key.lo); cur.hi = __ldcg(&s->key.hi);
        int claimed = 0;
        // a half equal to all-ones may be a torn read of a slot being claimed: let the CAS decide
        if (cur.lo == ~0ull || cur.hi == ~0ull) {
            cur = cas128(&s->key, empty, key);
            if (key_eq(cur, empty)) { if (weight > 1u) atomicAdd(&s->cnt, weight - 1u); return 1; }   // claimed: count 0 == seen once
        }
        if (key_eq(cur, key)) { atomicAdd(&s->cnt, weight); return claimed; }
        if (++slot == size) slot = 0;
    }
    return -1;
}

// One warp per 32 super-k-mer records.  The records go to shared memory, an exclusive
// prefix sum of their k-mer counts maps k-mer slot t to (record, offset): with M the
// bitmask of record starts inside the current block of 32 slots (one redux.or),
// record(t) = #starts before the block + popc(M & lanes<=t) - 1.  Every lane then cuts
// its own k-mer out of the record, canonicalises it and inserts it — no lane idles
// because its record is shorter than its neighbour's.  An empty slot is {key = all ones,
// count = 0}; claiming it needs no increment, so the stored count is occurrences - 1.
template <bool WIDE>
__global__ void __launch_bounds__(256) k_count_ht(const CountParams P) {
    typedef typename Traits<WIDE>::Slot Slot;
    typedef typename Traits<WIDE>::Key Key;
    constexpr int RW = Traits<WIDE>::kRecWords;
    __shared__ uint64_t s_rec[8][32 * RW];
    __shared__ uint32_t s_off[8][32];
    __shared__ uint32_t s_wt[WIDE ? 1 : 8][32];
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long r = P.rec_lo + ((unsigned long long)blockIdx.x * 8 + warp) * 32 + lane;
    const bool in = r < P.rec_hi;
    uint64_t w[RW];
    uint32_t n = 0; int bin = -1; uint32_t wt = 1u;
    if (in) {
        if constexpr (!WIDE) { if (P.weights) wt = P.weights[r]; }          // only 16-byte records are ever folded
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(P.records) + (RW / 2) * r;
#pragma unroll
        for (int i = 0; i < RW / 2; i++) { ulonglong2 v = __ldcs(src + i); w[2 * i] = v.x; w[2 * i + 1] = v.y; }   // read once: keep L2 for the tables
        n = (uint32_t)(w[RW - 1] & 0xFFull);
        w[RW - 1] &= ~0xFFull;
        bin = find_bin(P.bin_base, P.bin_lo, P.bin_hi, r);
#pragma unroll
        for (int i = 0; i < RW; i++) s_rec[warp][lane * RW + i] = w[i];
    }
    uint32_t incl = n;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { uint32_t t = __shfl_up_sync(FULL, incl, dd); if (lane >= dd) incl += t; }
    const uint32_t excl = incl - n;
    const uint32_t T = __shfl_sync(FULL, incl, 31);
    if (T == 0) return;
    s_off[warp][lane] = excl;
    if constexpr (!WIDE) s_wt[warp][lane] = wt;
    __syncwarp();
    const int bin0 = __shfl_sync(FULL, bin, 0);
    const bool uniform = __all_sync(FULL, bin == bin0 || bin < 0);
    unsigned int claims = 0; bool ovf = false;
    if (uniform) {
        const unsigned long long tb0 = P.tbl_base[bin0 - P.bin_lo];
        const unsigned long long size = P.tbl_base[bin0 - P.bin_lo + 1] - tb0;
        Slot* tbl = reinterpret_cast<Slot*>(P.table) + tb0;
        // Lane refill: the warp's T k-mers form a pool; a lane whose k-mer is done takes the next unassigned
        // one at once (ballot + popc, no atomics), so every round trip to L2 / HBM carries one probe per
        // lane — the warp never waits for its longest probe chain.  Per-lane state machine:
        // 0 = read the slot, 1 = CAS the empty slot, 2 = done (wants a refill), 3 = idle (pool exhausted).
        const uint32_t lt_mask = (1u << lane) - 1u;
        uint32_t next = 0;                                   // first unassigned k-mer of the pool (warp-uniform)
        const int first_state = P.first_state;
        Key key = Key(); unsigned long long slot = 0; int state = 2; uint32_t kw = 1u;   // kw: weight of the k-mer's record
        for (int round = 0;; round++) {
            const uint32_t want = __ballot_sync(FULL, state == 2);
            if (want) {
                if (state == 2) {
                    const uint32_t t = next + __popc(want & lt_mask);
                    if (t < T) {
                        int lo_ = 0, hi_ = 32;               // record ri with s_off[ri] <= t < s_off[ri+1]
#pragma unroll
                        for (int it = 0; it < 5; it++) { const int mid = (lo_ + hi_) >> 1; if (s_off[warp][mid] <= t) lo_ = mid; else hi_ = mid; }
                        const int j = (int)(t - s_off[warp][lo_]);
                        if constexpr (!WIDE) key = kmer_at_narrow(&s_rec[warp][lo_ * RW], j, P.k);
                        else key = kmer_at_wide(&s_rec[warp][lo_ * RW], j, P.k);
                        slot = slot_of(key_hash(key), size);
                        if constexpr (!WIDE) kw = s_wt[warp][lo_];
                        state = first_state;
                    } else state = 3;
                }
                next += __popc(want);
            }
            if (__all_sync(FULL, state == 3)) break;
            Key got = Key();
            if constexpr (!WIDE) {
                SlotN* sp = reinterpret_cast<SlotN*>(tbl) + slot;
                if (state == 0) got = __ldcg(&sp->key);
                else if (state == 1) got = atomicCAS((unsigned long long*)&sp->key, ~0ull, (unsigned long long)key);
            } else {
                SlotW* sp = reinterpret_cast<SlotW*>(tbl) + slot;
                if (state == 0) { got.lo = __ldcg(&sp->key.lo); got.hi = __ldcg(&sp->key.hi); }
                else if (state == 1) { const key128 empty = {~0ull, ~0ull}; got = cas128(&sp->key, empty, key); }
            }
            if (state < 2) {
                bool is_empty, maybe_empty;
                if constexpr (!WIDE) { is_empty = got == ~0ull; maybe_empty = is_empty; }
                else {      // a half equal to all-ones may be a torn read of a slot being claimed: the CAS decides
                    is_empty = got.lo == ~0ull && got.hi == ~0ull;
                    maybe_empty = got.lo == ~0ull || got.hi == ~0ull;
                }
                uint32_t* cp;
                if constexpr (!WIDE) cp = &(reinterpret_cast<SlotN*>(tbl) + slot)->cnt;
                else cp = &(reinterpret_cast<SlotW*>(tbl) + slot)->cnt;
                if (state == 1 && is_empty) { claims++; if (kw > 1u) atomicAdd(cp, kw - 1u); state = 2; }   // claimed: count 0 == seen once, no RED
                else if (state == 0 && maybe_empty) state = 1;               // (before any comparison: a torn view of a slot being claimed must not match)
                else if (key_eq(got, key)) { atomicAdd(cp, kw); state = 2; }
                else { if (++slot == size) slot = 0; state = first_state; }
            }
            if (round > (int)T + 2 * P.max_probe) { ovf = true; break; }            // warp-uniform bound on the rounds
        }
        const unsigned int tot = __reduce_add_sync(FULL, claims);
        if (lane == 0 && tot) atomicAdd(&P.bin_distinct[bin0], (unsigned long long)tot);
    } else if (in) {
        // the 32 records straddle a bin boundary (rare): one lane per record
        const unsigned long long tb0 = P.tbl_base[bin - P.bin_lo];
        const unsigned long long size = P.tbl_base[bin - P.bin_lo + 1] - tb0;
        Slot* tbl = reinterpret_cast<Slot*>(P.table) + tb0;
        for (int j = 0; j < (int)n; j++) {
            int c;
            if constexpr (!WIDE) c = ht_insert(reinterpret_cast<SlotN*>(tbl), size, kmer_at_narrow(&s_rec[warp][lane * RW], j, P.k), P.max_probe, wt);
            else c = ht_insert(reinterpret_cast<SlotW*>(tbl), size, kmer_at_wide(&s_rec[warp][lane * RW], j, P.k), P.max_probe, wt);
            if (c < 0) ovf = true; else claims += (unsigned)c;
        }
        if (claims) atomicAdd(&P.bin_distinct[bin], (unsigned long long)claims);
    }
    if (ovf) *P.overflow = 1;
}

// Record folding: every NARROW record (16 bytes, brought into canonical form first, see canon_record_narrow) is inserted
// as a 128-bit key into its bin's table of SlotW; the slot count becomes its multiplicity - 1.  The tables are
// then compacted by k_compact_ht<true> into dense (record, multiplicity) arrays, bin-major like the input.  A
// record is never all ones (its low byte is n <= 60), so the table's EMPTY key cannot collide.
struct FoldParams {
    const void* records; unsigned long long rec_lo, rec_hi;
    const unsigned long long* bin_base; int bin_lo, bin_hi;
    void* table; const unsigned long long* tbl_base;
    unsigned long long* bin_distinct;           // [B] distinct records per bin
    int* overflow; int max_probe; int k;
};
// One warp per kFoldPerWarp consecutive records, staged in shared memory; the lanes then run k_count_ht's refill state
// machine over that pool (0 = read the slot, 1 = CAS the empty slot, 2 = done, take the next record, 3 = pool
// exhausted), so every round trip carries one probe per lane whatever the lengths of the probe chains.
template <int kFoldPerWarp>
__global__ void __launch_bounds__(256) k_fold_insert(const FoldParams P) {
    __shared__ ulonglong2 s_rec[8][kFoldPerWarp];
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long r0 = P.rec_lo + ((unsigned long long)blockIdx.x * 8 + warp) * kFoldPerWarp;
    if (r0 >= P.rec_hi) return;                              // warp-uniform
    const uint32_t T = (uint32_t)min((unsigned long long)kFoldPerWarp, P.rec_hi - r0);
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(P.records) + r0;
    for (uint32_t i = lane; i < T; i += 32) {                // stage the records in canonical form (see canon_record_narrow)
        const ulonglong2 v = __ldcs(src + i);
        uint64_t w0 = v.x, w1 = v.y & ~0xFFull; const int n = (int)(v.y & 0xFFull);
        canon_record_narrow(w0, w1, n + P.k - 1);
        s_rec[warp][i] = make_ulonglong2(w0, w1 | (uint64_t)n);
    }
    __syncwarp();
    const int bin_first = find_bin(P.bin_base, P.bin_lo, P.bin_hi, r0);
    const int bin_last = find_bin(P.bin_base, P.bin_lo, P.bin_hi, r0 + T - 1);
    bool ovf = false;
    if (bin_first == bin_last) {
        const unsigned long long tb0 = P.tbl_base[bin_first - P.bin_lo];
        const unsigned long long size = P.tbl_base[bin_first - P.bin_lo + 1] - tb0;
        SlotW* tbl = reinterpret_cast<SlotW*>(P.table) + tb0;
        const uint32_t lt_mask = (1u << lane) - 1u;
        const key128 empty = {~0ull, ~0ull};
        uint32_t next = 0; unsigned int claims = 0;
        key128 key = empty; unsigned long long slot = 0; int state = 2;
        for (int round = 0;; round++) {
            const uint32_t want = __ballot_sync(FULL, state == 2);
            if (want) {
                if (state == 2) {
                    const uint32_t t = next + __popc(want & lt_mask);
                    if (t < T) {
                        const ulonglong2 v = s_rec[warp][t];
                        key.lo = v.x; key.hi = v.y;          // the memory image of the record
                        slot = slot_of(key_hash(key), size);
                        state = 0;
                    } else state = 3;
                }
                next += __popc(want);
            }
            if (__all_sync(FULL, state == 3)) break;
            SlotW* sp = tbl + slot;
            key128 got = empty;
            if (state == 0) { const ulonglong2 kv = __ldcg(reinterpret_cast<const ulonglong2*>(&sp->key)); got.lo = kv.x; got.hi = kv.y; }
            else if (state == 1) got = cas128(&sp->key, empty, key);
            if (state < 2) {
                // a half equal to all-ones may be a torn read of a slot being claimed: the CAS decides
                const bool is_empty = got.lo == ~0ull && got.hi == ~0ull;
                const bool maybe_empty = got.lo == ~0ull || got.hi == ~0ull;
                if (state == 1 && is_empty) { claims++; state = 2; }              // count 0 == seen once
                else if (key_eq(got, key)) { atomicAdd(&sp->cnt, 1u); state = 2; }
                else if (state == 0 && maybe_empty) state = 1;
                else { if (++slot == size) slot = 0; state = 0; }
            }
            if (round > (int)T + 2 * P.max_probe) { ovf = true; break; }           // warp-uniform bound on the rounds
        }
        const unsigned int tot = __reduce_add_sync(FULL, claims);
        if (lane == 0 && tot) atomicAdd(&P.bin_distinct[bin_first], (unsigned long long)tot);
    } else {
        // the warp's records straddle a bin boundary (rare): one lane per record
        for (uint32_t i = lane; i < T; i += 32) {
            const int bin = find_bin(P.bin_base, bin_first, bin_last + 1, r0 + i);
            const unsigned long long tb0 = P.tbl_base[bin - P.bin_lo];
            const unsigned long long size = P.tbl_base[bin - P.bin_lo + 1] - tb0;
            const ulonglong2 v = s_rec[warp][i];
            key128 key; key.lo = v.x; key.hi = v.y;
            const int c = ht_insert(reinterpret_cast<SlotW*>(P.table) + tb0, size, key, P.max_probe);
            if (c < 0) ovf = true; else if (c) atomicAdd(&P.bin_distinct[bin], 1ull);
        }
    }
    if (ovf) *P.overflow = 1;
}

// empty table: key = all ones, count = 0
template <bool WIDE>
__global__ void __launch_bounds__(256) k_fill_table(void* table, unsigned long long n_slots) {
    ulonglong2* t = reinterpret_cast<ulonglong2*>(table);
    const unsigned long long n16 = n_slots * (WIDE ? 2ull : 1ull);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * blockDim.x) {
        if constexpr (!WIDE) t[i] = make_ulonglong2(~0ull, 0ull);
        else t[i] = (i & 1ull) ? make_ulonglong2(0ull, 0ull) : make_ulonglong2(~0ull, ~0ull);
    }
}

struct CompactParams {
    void* table; unsigned long long n_slots;               // batch table space; every 1024-slot tile lies in one bin
    const unsigned long long* tbl_base; int n_bins; int bin_lo;
    const unsigned long long* out_base;                    // [B+1] global output offsets (entries)
    unsigned long long out_origin;                         // global offset of the first entry of the output arrays
    unsigned long long out_cap;                            // entries the output arrays can hold
    unsigned long long* out_cursor;                        // [B]
    void* out_keys; uint32_t* out_cnt;
    int clear;                                             // 1: write EMPTY back into every occupied slot (table reusable without a memset)
    int* cap_overflow;                                     // set when the output arrays are too small
    unsigned long long* acc;                               // [3][64] digest accumulators (sum, xor, count), or NULL
    int bin_shift;                                         // the digest's bin id = internal bin >> bin_shift
};
// table -> dense output, persistent grid-stride over 1024-slot tiles.  Order inside a bin
// is slot order up to tile permutation (the reference's HT order is fastutil's iteration
// order: unspecified, SBKC:723).  Also folds the entries into the result digest.
template <bool WIDE>
__global__ void __launch_bounds__(256) k_compact_ht(const CompactParams P) {
    typedef typename Traits<WIDE>::Slot Slot;
    typedef typename Traits<WIDE>::Key Key;
    __shared__ Key s_keys[1024];                           // the tile's entries, dense (thread-major slot order)
    __shared__ uint32_t s_cnts[1024];
    __shared__ unsigned int s_warp[8];
    __shared__ unsigned long long s_base;
    __shared__ int s_bin;
    __shared__ unsigned long long s_lo, s_hi;              // slot range of the bin the previous tile belonged to
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Slot* tbl = reinterpret_cast<Slot*>(P.table);
    const unsigned long long n_tiles = (P.n_slots + 1023) / 1024;
    // every CTA owns a contiguous range of tiles: consecutive tiles are nearly always in the same bin, so the
    // serial part of a tile (which bin? reserve output space) is one compare and one atomic
    const unsigned long long per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const unsigned long long tile_begin = (unsigned long long)blockIdx.x * per_cta;
    const unsigned long long tile_end = min(n_tiles, tile_begin + per_cta);
    if (threadIdx.x == 0) { s_lo = 1; s_hi = 0; s_bin = 0; }
    unsigned long long dsum = 0, dxor = 0, dcnt = 0;
    for (unsigned long long tile = tile_begin; tile < tile_end; tile++) {
        const unsigned long long tile0 = tile * 1024ull;
        Key keys[4]; uint32_t cnts[4]; unsigned int have = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned long long sl = tile0 + (unsigned long long)j * 256ull + threadIdx.x;
            cnts[j] = 0u;
            if (sl < P.n_slots) {
                if constexpr (!WIDE) {
                    ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&tbl[sl]);
                    keys[j] = v.x; cnts[j] = (uint32_t)v.y;
                    if (v.x != ~0ull) { have |= 1u << j; if (P.clear) *reinterpret_cast<ulonglong2*>(&tbl[sl]) = make_ulonglong2(~0ull, 0ull); }
                } else {
                    ulonglong2* p = reinterpret_cast<ulonglong2*>(&tbl[sl]);
                    ulonglong2 kv = p[0]; ulonglong2 cv = p[1];
                    key128 kk; kk.lo = kv.x; kk.hi = kv.y;
                    keys[j] = kk; cnts[j] = (uint32_t)cv.x;
                    if (!(kv.x == ~0ull && kv.y == ~0ull)) {
                        have |= 1u << j;
                        if (P.clear) { p[0] = make_ulonglong2(~0ull, ~0ull); p[1] = make_ulonglong2(0ull, 0ull); }
                    }
                }
            }
        }
        const unsigned int c = __popc(have);
        unsigned int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();                                   // also: the previous tile's reads of s_keys / s_cnts / s_base are done
        unsigned int wbase = 0, total = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { unsigned int t = s_warp[i]; if (i < warp) wbase += t; total += t; }
        if (total) {                                       // block-uniform
            if (threadIdx.x == 0) {
                if (!(tile0 >= s_lo && tile0 < s_hi)) {
                    const int bl = find_bin(P.tbl_base, 0, P.n_bins, tile0);
                    s_bin = P.bin_lo + bl; s_lo = P.tbl_base[bl]; s_hi = P.tbl_base[bl + 1];
                }
                s_base = P.out_base[s_bin] - P.out_origin + atomicAdd(&P.out_cursor[s_bin], (unsigned long long)total);
            }
            unsigned int o = wbase + (incl - c);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (have & (1u << j)) { s_keys[o] = keys[j]; s_cnts[o] = cnts[j] + 1u; o++; }
        }
        __syncthreads();
        if (total == 0) continue;
        const unsigned long long base = s_base; const uint32_t bin = (uint32_t)s_bin >> P.bin_shift;
        if (base + total > P.out_cap) { if (threadIdx.x == 0) *P.cap_overflow = 1; continue; }
        // dense part: coalesced stores, every lane busy in the digest
        const uint64_t hbin = mix64((uint64_t)bin);
        const uint64_t hpre = mix64(hbin);                 // entry_hash's inner term when hi == 0 (narrow keys)
        for (unsigned int i = threadIdx.x; i < total; i += 256u) {
            const Key kk = s_keys[i]; const uint32_t n = s_cnts[i];
            reinterpret_cast<Key*>(P.out_keys)[base + i] = kk;
            P.out_cnt[base + i] = n;
            if (P.acc) {
                uint64_t h;                                // == entry_hash(bin, hi, lo)
                if constexpr (!WIDE) h = mix64(kk ^ hpre); else h = mix64(kk.lo ^ mix64(kk.hi ^ hbin));
                dsum += h * (uint64_t)n; dxor ^= mix64(h + n); dcnt += n;
            }
        }
    }
    if (P.acc) {
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            dsum += __shfl_xor_sync(0xFFFFFFFFu, dsum, d); dxor ^= __shfl_xor_sync(0xFFFFFFFFu, dxor, d); dcnt += __shfl_xor_sync(0xFFFFFFFFu, dcnt, d);
        }
        if (lane == 0 && dcnt) {
            const int a = (blockIdx.x * 8 + warp) & 63;
            atomicAdd(&P.acc[a], dsum); atomicXor(&P.acc[64 + a], dxor); atomicAdd(&P.acc[128 + a], dcnt);
        }
    }
}

// ------------------------------------------------------------------ K5 (sort path): expand
struct ExpandParams {
    const void* records; unsigned long long rec_lo, rec_hi;
    const unsigned long long* bin_base; int bin_lo, bin_hi;
    const unsigned long long* key_base;       // [bin_hi-bin_lo+1] key offsets inside the batch key array
    unsigned long long* key_cursor;           // [B]
    void* keys; int k;
};
template <bool WIDE>
__global__ void __launch_bounds__(256) k_expand(const ExpandParams P) {
    const unsigned long long r = P.rec_lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.rec_hi) return;
    const int bin = find_bin(P.bin_base, P.bin_lo, P.bin_hi, r);
    if constexpr (!WIDE) {
        ulonglong2 rec = reinterpret_cast<const ulonglong2*>(P.records)[r];
        const int n = (int)(rec.y & 0xFFull);
        unsigned long long o = P.key_base[bin - P.bin_lo] + atomicAdd(&P.key_cursor[bin], (unsigned long long)n);
        uint64_t* dst = reinterpret_cast<uint64_t*>(P.keys);
        for_each_kmer_narrow(rec.x, rec.y, P.k, [&](uint64_t key) { dst[o++] = key; });
    } else {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(P.records) + 2 * r;
        ulonglong2 a = src[0], b = src[1];
        const int n = (int)(b.y & 0xFFull);
        unsigned long long o = P.key_base[bin - P.bin_lo] + atomicAdd(&P.key_cursor[bin], (unsigned long long)n);
        key128* dst = reinterpret_cast<key128*>(P.keys);
        for_each_kmer_wide(a.x, a.y, b.x, b.y, P.k, [&](key128 key) { dst[o++] = key; });
    }
}

// ------------------------------------------------------------------ K6b: segmented radix sort
// Segments = bins of the batch.  Tiles of 2048 keys never straddle a segment.
//
// Fast path (MSD first): ONE global partition pass on the top `bits` (<= 10) bits of the key
// (k_radix_hist / k_radix_scan / k_radix_scatter with nd = 2^bits digits), k_form_chunks groups the
// resulting sub-buckets into chunks of <= kLocalCap keys, and k_radix_local sorts every chunk
// completely inside shared memory (LSD, 8-bit digits) and writes it back in place: the whole
// segment is then ascending.  HBM traffic: 2 reads + 1 write (partition) + 1 read + 1 write.
// Fallback (a sub-bucket larger than a chunk): the same three kernels run one LSD pass per 8 bits.
static constexpr int kSortTile = 2048;
static constexpr int kMaxDigits = 1024;
struct SortParams {
    const void* in; void* out;
    const unsigned long long* seg_base;      // [n_seg+1] key offsets
    const unsigned int* tile_seg;            // [n_tiles] segment of each tile
    const unsigned int* seg_tile0;           // [n_seg+1] first tile of each segment
    unsigned int* tile_hist;                 // [n_tiles*nd]
    unsigned int* seg_digit_tot;             // [n_seg*nd] keys per (segment, digit), or NULL
    unsigned int n_tiles; int n_seg; int shift; int nd;   // digit = (key >> shift) & (nd-1)
};
template <bool WIDE> __device__ __forceinline__ unsigned int digit_of(typename Traits<WIDE>::Key key, int shift, unsigned int mask);
template <> __device__ __forceinline__ unsigned int digit_of<false>(uint64_t key, int shift, unsigned int mask) { return (unsigned int)(key >> shift) & mask; }
template <> __device__ __forceinline__ unsigned int digit_of<true>(key128 key, int shift, unsigned int mask) {
    uint64_t v;
    if (shift >= 64) v = key.hi >> (shift - 64);
    else v = shift ? ((key.lo >> shift) | (key.hi << (64 - shift))) : key.lo;
    return (unsigned int)v & mask;
}

template <bool WIDE>
__global__ void __launch_bounds__(256) k_radix_hist(const SortParams P) {
    typedef typename Traits<WIDE>::Key Key;
    __shared__ unsigned int s_h[kMaxDigits];
    const unsigned int t = blockIdx.x;
    const unsigned int seg = P.tile_seg[t];
    const unsigned long long lo = P.seg_base[seg] + (unsigned long long)(t - P.seg_tile0[seg]) * kSortTile;
    const unsigned long long hi = min(lo + (unsigned long long)kSortTile, P.seg_base[seg + 1]);
    for (int d = threadIdx.x; d < P.nd; d += 256) s_h[d] = 0;
    __syncthreads();
    const Key* in = reinterpret_cast<const Key*>(P.in);
    const unsigned int mask = (unsigned int)P.nd - 1u;
    for (unsigned long long i = lo + threadIdx.x; i < hi; i += 256) atomicAdd(&s_h[digit_of<WIDE>(in[i], P.shift, mask)], 1u);
    __syncthreads();
    for (int d = threadIdx.x; d < P.nd; d += 256) P.tile_hist[(size_t)t * P.nd + d] = s_h[d];
}

// one CTA per segment: tile_hist[t][d] <- exclusive offset of (digit d, tile t) inside the
// segment (digit-major, tile-minor); seg_digit_tot[seg][d] <- keys of the segment with digit d.
__global__ void __launch_bounds__(256) k_radix_scan(const SortParams P) {
    __shared__ unsigned int s_tot[kMaxDigits];
    __shared__ unsigned int s_base[kMaxDigits];
    const int seg = blockIdx.x;
    const unsigned int t0 = P.seg_tile0[seg], t1 = P.seg_tile0[seg + 1];
    constexpr int UB = 8;                          // tiles whose counters are loaded together (independent loads in flight)
    for (int d = threadIdx.x; d < P.nd; d += 256) {
        unsigned int tot = 0;
        unsigned int t = t0;
        for (; t + UB <= t1; t += UB) {
            unsigned int c[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) c[u] = P.tile_hist[(size_t)(t + u) * P.nd + d];
#pragma unroll
            for (int u = 0; u < UB; u++) tot += c[u];
        }
        for (; t < t1; t++) tot += P.tile_hist[(size_t)t * P.nd + d];
        s_tot[d] = tot;
        if (P.seg_digit_tot) P.seg_digit_tot[(size_t)seg * P.nd + d] = tot;
    }
    __syncthreads();
    if (threadIdx.x < 32) {                       // exclusive scan over the digits by one warp
        unsigned int carry = 0;
        for (int d0 = 0; d0 < P.nd; d0 += 32) {
            const unsigned int v = s_tot[d0 + threadIdx.x];
            unsigned int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)threadIdx.x >= o) incl += t; }
            s_base[d0 + threadIdx.x] = carry + incl - v;
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < P.nd; d += 256) {
        unsigned int run = s_base[d];
        unsigned int t = t0;
        for (; t + UB <= t1; t += UB) {
            unsigned int c[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) c[u] = P.tile_hist[(size_t)(t + u) * P.nd + d];
#pragma unroll
            for (int u = 0; u < UB; u++) { P.tile_hist[(size_t)(t + u) * P.nd + d] = run; run += c[u]; }
        }
        for (; t < t1; t++) {
            unsigned int c = P.tile_hist[(size_t)t * P.nd + d];
            P.tile_hist[(size_t)t * P.nd + d] = run;
            run += c;
        }
    }
}

// stable scatter of one tile: warp `wi` owns keys [wi*256, wi*256+256) of the tile
// in 8 rounds of 32; ranks come from match.any peer masks and per-warp digit counters.
template <bool WIDE>
__global__ void __launch_bounds__(256, 5) k_radix_scatter(const SortParams P) {
    typedef typename Traits<WIDE>::Key Key;
    __shared__ unsigned int s_wc[8][kMaxDigits];
    __shared__ unsigned int s_off[kMaxDigits];
    const unsigned int t = blockIdx.x;
    const unsigned int seg = P.tile_seg[t];
    const unsigned long long seg0 = P.seg_base[seg];
    const unsigned long long lo = seg0 + (unsigned long long)(t - P.seg_tile0[seg]) * kSortTile;
    const unsigned long long hi = min(lo + (unsigned long long)kSortTile, P.seg_base[seg + 1]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int mask = (unsigned int)P.nd - 1u;
    for (int w = 0; w < 8; w++) for (int d = threadIdx.x; d < P.nd; d += 256) s_wc[w][d] = 0;
    for (int d = threadIdx.x; d < P.nd; d += 256) s_off[d] = P.tile_hist[(size_t)t * P.nd + d];
    __syncthreads();
    const Key* in = reinterpret_cast<const Key*>(P.in);
    Key* out = reinterpret_cast<Key*>(P.out);
    Key keys[8]; unsigned int rank[8]; unsigned int dig[8];
#pragma unroll
    for (int rd = 0; rd < 8; rd++) {
        const unsigned long long i = lo + (unsigned long long)(warp * 256 + rd * 32 + lane);
        const bool ok = i < hi;
        if (ok) keys[rd] = in[i];
        const unsigned int d = ok ? digit_of<WIDE>(keys[rd], P.shift, mask) : ((unsigned)kMaxDigits + (unsigned)lane);
        const unsigned int peers = __match_any_sync(0xFFFFFFFFu, d);
        unsigned int old = 0;
        if (ok) old = s_wc[warp][d];
        __syncwarp();
        if (ok && (peers & ((1u << lane) - 1u)) == 0u) s_wc[warp][d] = old + __popc(peers);
        __syncwarp();
        rank[rd] = old + __popc(peers & ((1u << lane) - 1u));
        dig[rd] = d;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < P.nd; d += 256) {      // exclusive prefix over the 8 warps
        unsigned int run = 0;
#pragma unroll
        for (int wi = 0; wi < 8; wi++) { unsigned int c = s_wc[wi][d]; s_wc[wi][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int rd = 0; rd < 8; rd++) {
        if (dig[rd] < (unsigned)kMaxDigits) out[seg0 + s_off[dig[rd]] + s_wc[warp][dig[rd]] + rank[rd]] = keys[rd];
    }
}

// groups the sub-buckets of every segment into chunks of <= cap keys (one thread per segment);
// sets *too_big when a single sub-bucket exceeds cap.
struct ChunkDesc { unsigned long long start; unsigned int n; unsigned int pad; };
__global__ void k_form_chunks(const unsigned long long* seg_base, const unsigned int* seg_digit_tot, int n_seg, int nd,
                              unsigned int cap, ChunkDesc* chunks, unsigned int max_chunks, unsigned int* n_chunks, int* too_big) {
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_seg) return;
    unsigned long long start = seg_base[seg];
    unsigned int acc = 0;
    auto emit = [&](unsigned int n) {
        if (!n) return;
        const unsigned int c = atomicAdd(n_chunks, 1u);
        if (c < max_chunks) { chunks[c].start = start; chunks[c].n = n; chunks[c].pad = 0; } else *too_big = 1;
        start += n;
    };
    for (int d = 0; d < nd; d++) {
        const unsigned int c = seg_digit_tot[(size_t)seg * nd + d];
        if (c > cap) *too_big = 1;
        if (acc + c > cap) { emit(acc); acc = 0; }
        acc += c;
    }
    emit(acc);
}

// One CTA sorts one chunk (<= kLocalCap keys) completely in shared memory: LSD over all
// `n_pass` 8-bit digits of the key, two key buffers, per-warp digit counters, stable ranks
// from match.any.  512 threads; warp w owns keys [w*per_warp, (w+1)*per_warp) of the chunk.
template <bool WIDE> struct LocalSort { static constexpr int kCap = WIDE ? 5120 : 10240; };
template <bool WIDE>
__global__ void __launch_bounds__(512) k_radix_local(const void* in, void* out, const ChunkDesc* chunks, unsigned int n_chunks, int n_pass) {
    typedef typename Traits<WIDE>::Key Key;
    constexpr int CAP = LocalSort<WIDE>::kCap;
    constexpr int NW = 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Key* bufA = reinterpret_cast<Key*>(smem_raw);
    Key* bufB = bufA + CAP;
    unsigned short* s_rank = reinterpret_cast<unsigned short*>(bufB + CAP);
    unsigned int* s_wc = reinterpret_cast<unsigned int*>(s_rank + CAP);        // [NW][256]
    __shared__ unsigned int s_warp_tot[NW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Key* gin = reinterpret_cast<const Key*>(in);
    Key* gout = reinterpret_cast<Key*>(out);
    for (unsigned int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const unsigned long long start = chunks[c].start;
        const int n = (int)chunks[c].n;
        const int per_warp = ((n + NW - 1) / NW + 31) & ~31;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += 512) bufA[i] = gin[start + i];
        Key* src = bufA; Key* dst = bufB;
        for (int p = 0; p < n_pass; p++) {
            const int shift = 8 * p;
            for (int i = threadIdx.x; i < NW * 256; i += 512) s_wc[i] = 0;
            __syncthreads();
            const int w0 = warp * per_warp, w1 = min(n, w0 + per_warp);
            for (int i0 = w0; i0 < w1; i0 += 32) {
                const int i = i0 + lane;
                const bool ok = i < w1;
                unsigned int d = 256u + (unsigned)lane;
                if (ok) d = digit_of<WIDE>(src[i], shift, 255u);
                // lanes with the same digit: eight ballots (MATCH.ANY serialises the whole SM: ~50 cycles per warp instruction, scripts/microbench)
                // (measured: config 3 count 83.5 -> 71.2 ms; 128-bit keys, one 186 KB CTA per SM, are faster with the single MATCH.ANY: 533 vs 653 ms)
                unsigned int peers;
                if constexpr (WIDE) peers = __match_any_sync(0xFFFFFFFFu, d);
                else {
                    peers = __ballot_sync(0xFFFFFFFFu, ok);
#pragma unroll
                    for (int bit = 0; bit < 8; bit++) {
                        const bool one = (d >> bit) & 1u;
                        const unsigned int bal = __ballot_sync(0xFFFFFFFFu, one);
                        peers &= one ? bal : ~bal;
                    }
                    if (!ok) peers = 1u << lane;
                }
                unsigned int old = 0;
                if (ok) old = s_wc[warp * 256 + d];
                __syncwarp();
                if (ok && (peers & ((1u << lane) - 1u)) == 0u) s_wc[warp * 256 + d] = old + __popc(peers);
                __syncwarp();
                if (ok) s_rank[i] = (unsigned short)(old + __popc(peers & ((1u << lane) - 1u)));
            }
            __syncthreads();
            if (threadIdx.x < 256) {            // digit d = threadIdx.x: prefix over warps, then over digits
                const int d = threadIdx.x;
                unsigned int run = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) { unsigned int t = s_wc[w * 256 + d]; s_wc[w * 256 + d] = run; run += t; }
                unsigned int incl = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
                if (lane == 31) s_warp_tot[warp] = incl;
                asm volatile("bar.sync 1, 256;");
                unsigned int base = incl - run;
                for (int w = 0; w < warp; w++) base += s_warp_tot[w];
#pragma unroll
                for (int w = 0; w < NW; w++) s_wc[w * 256 + d] += base;
            }
            __syncthreads();
            for (int i0 = w0; i0 < w1; i0 += 32) {
                const int i = i0 + lane;
                if (i < w1) {
                    const Key key = src[i];
                    dst[s_wc[warp * 256 + digit_of<WIDE>(key, shift, 255u)] + s_rank[i]] = key;
                }
            }
            __syncthreads();
            Key* t = src; src = dst; dst = t;
        }
        for (int i = threadIdx.x; i < n; i += 512) gout[start + i] = src[i];
    }
}

// ------------------------------------------------------------------ K6b: run-length count of sorted keys
struct RleParams {
    const void* keys;                        // sorted inside each segment
    const unsigned long long* seg_base; const unsigned int* tile_seg; const unsigned int* seg_tile0;
    unsigned int n_tiles; int bin_lo;
    unsigned int* tile_heads;                // [n_tiles+1] heads per tile, then exclusive scan in place
    unsigned long long* bin_distinct;        // [B]
    unsigned long long out_off;              // global entry offset of this batch
    void* out_keys; uint32_t* out_cnt;       // global output
    unsigned long long* first_idx;           // [D_batch+1] scratch: batch-relative index of each run head
    unsigned long long n_keys;               // keys in batch
};
template <bool WIDE, int PASS>   // PASS 0: count heads per tile; PASS 1: write heads
__global__ void __launch_bounds__(256) k_rle(const RleParams P) {
    typedef typename Traits<WIDE>::Key Key;
    __shared__ unsigned int s_warp[8];
    const unsigned int t = blockIdx.x;
    const unsigned int seg = P.tile_seg[t];
    const unsigned long long seg0 = P.seg_base[seg];
    const unsigned long long lo = seg0 + (unsigned long long)(t - P.seg_tile0[seg]) * kSortTile;
    const unsigned long long hi = min(lo + (unsigned long long)kSortTile, P.seg_base[seg + 1]);
    const Key* keys = reinterpret_cast<const Key*>(P.keys);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // thread owns 8 consecutive keys (blocked) so that ranks follow key order
    const unsigned long long i0 = lo + (unsigned long long)threadIdx.x * 8ull;
    unsigned int flags = 0;
    Key prev;
    if (i0 < hi && i0 > seg0) prev = keys[i0 - 1];
    Key mine[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const unsigned long long i = i0 + j;
        if (i < hi) {
            mine[j] = keys[i];
            bool head = (i == seg0) || !key_eq(mine[j], prev);
            if (head) flags |= 1u << j;
            prev = mine[j];
        }
    }
    unsigned int c = __popc(flags), incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned int v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned int wbase = 0, total = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { unsigned int v = s_warp[i]; if (i < warp) wbase += v; total += v; }
    if (PASS == 0) {
        if (threadIdx.x == 0) {
            P.tile_heads[t] = total;
            if (total) atomicAdd(&P.bin_distinct[P.bin_lo + (int)seg], (unsigned long long)total);
        }
    } else {
        unsigned long long o = (unsigned long long)P.tile_heads[t] + wbase + (incl - c);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (flags & (1u << j)) {
                reinterpret_cast<Key*>(P.out_keys)[P.out_off + o] = mine[j];
                P.first_idx[o] = i0 + j;
                o++;
            }
        }
    }
}
// exclusive scan of a u32 array in place by one CTA; a[n] receives the total
__global__ void __launch_bounds__(1024) k_scan_u32(unsigned int* a, unsigned int n) {
    __shared__ unsigned int s_w[32];
    __shared__ unsigned int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (unsigned int base = 0; base < n; base += 1024) {
        unsigned int i = base + threadIdx.x;
        unsigned int v = (i < n) ? a[i] : 0u, incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        unsigned int wb = 0, tot = 0;
        for (int j = 0; j < 32; j++) { unsigned int t = s_w[j]; if (j < warp) wb += t; tot += t; }
        unsigned int carry = s_carry;
        if (i < n) a[i] = carry + wb + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) a[n] = s_carry;
}
// counts of the runs: cnt[j] = first[j+1] - first[j], first[D] = n_keys
__global__ void k_rle_counts(const unsigned long long* first, unsigned long long n_heads, unsigned long long n_keys,
                             uint32_t* out_cnt, unsigned long long out_off) {
    unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_heads) return;
    unsigned long long nxt = (j + 1 < n_heads) ? first[j + 1] : n_keys;
    out_cnt[out_off + j] = (uint32_t)(nxt - first[j]);
}

// exclusive scan over bins [bin_lo, bin_hi) of bin_distinct -> out_base (global
// offsets continue from out_base[bin_lo], which the host has set).  One CTA.
__global__ void __launch_bounds__(256) k_bin_offsets(const unsigned long long* bin_distinct, unsigned long long* out_base,
                                                     int bin_lo, int bin_hi, unsigned long long* batch_total) {
    __shared__ unsigned long long s_w[8];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = out_base[bin_lo];
    __syncthreads();
    const unsigned long long start = s_carry;
    for (int base = bin_lo; base < bin_hi; base += 256) {
        int b = base + threadIdx.x;
        unsigned long long v = (b < bin_hi) ? bin_distinct[b] : 0ull, incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        unsigned long long wb = 0, tot = 0;
        for (int j = 0; j < 8; j++) { unsigned long long t = s_w[j]; if (j < warp) wb += t; tot += t; }
        unsigned long long carry = s_carry;
        if (b < bin_hi) out_base[b + 1] = carry + wb + incl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *batch_total = s_carry - start;
}

// ------------------------------------------------------------------ K7: digest
struct DigestParams {
    const void* keys; const uint32_t* cnt; const unsigned long long* out_base; int B;
    unsigned long long n; unsigned long long origin;  // entries in this chunk, global offset of its first entry
    unsigned long long* acc;                          // [3][64]: sum(h*cnt), xor(mix(h+cnt)), sum(cnt), spread over 64 slots
    int bin_shift;                                    // bin id of an entry = internal bin >> bin_shift
};
template <bool WIDE>
__global__ void __launch_bounds__(256) k_digest(const DigestParams P) {
    typedef typename Traits<WIDE>::Key Key;
    unsigned long long s = 0, x = 0, c = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        int bin = find_bin(P.out_base, 0, P.B, P.origin + i);
        uint64_t hi, lo;
        if constexpr (!WIDE) { hi = 0; lo = reinterpret_cast<const uint64_t*>(P.keys)[i]; }
        else { key128 kk = reinterpret_cast<const key128*>(P.keys)[i]; hi = kk.hi; lo = kk.lo; }
        uint32_t n = P.cnt[i];
        uint64_t h = entry_hash((uint32_t)bin >> P.bin_shift, hi, lo);
        s += h * (uint64_t)n; x ^= mix64(h + n); c += n;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        s += __shfl_xor_sync(0xFFFFFFFFu, s, d); x ^= __shfl_xor_sync(0xFFFFFFFFu, x, d); c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    }
    if ((threadIdx.x & 31) == 0) {
        const int a = (blockIdx.x * 8 + (threadIdx.x >> 5)) & 63;
        atomicAdd(&P.acc[a], s); atomicXor(&P.acc[64 + a], x); atomicAdd(&P.acc[128 + a], c);
    }
}

// ------------------------------------------------------------------ K7: text output (SBKC:579-606, 725-734; UTIL:416-454)
// Lines "<k letters>\t<decimal count>\n" for entries [first, first+n) of a result chunk.  Tile = 256 entries.
// PASS 0: bytes per tile.  PASS 1: format into shared memory, then one coalesced copy to `text`.
struct FmtParams {
    const void* keys; const uint32_t* cnt; unsigned long long first, n; int k;
    unsigned long long* tile_off;     // [n_tiles+1] bytes per tile -> (after the scan) text offset of each tile
    uint8_t* text;
    const unsigned long long* bounds; unsigned long long* bound_off; int n_bounds;   // entry index -> text offset (bin boundaries)
};
__device__ __forceinline__ int dec_digits(uint32_t v) {
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 :
           v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
template <bool WIDE, int PASS>
__global__ void __launch_bounds__(256) k_fmt(const FmtParams P) {
    __shared__ unsigned int s_warp[8];
    __shared__ uint8_t s_text[PASS ? 256 * (64 + 12) : 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    uint32_t c = 0; unsigned int len = 0;
    if (i < P.n) { c = P.cnt[P.first + i]; len = (unsigned)P.k + 2u + (unsigned)dec_digits(c); }
    unsigned int incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { unsigned int t = s_warp[w]; if (w < warp) wbase += t; total += t; }
    if (PASS == 0) { if (threadIdx.x == 0) P.tile_off[blockIdx.x] = total; return; }
    if (i < P.n) {
        uint8_t* o = s_text + wbase + incl - len;
        uint64_t hi = 0, lo;
        if constexpr (!WIDE) lo = reinterpret_cast<const uint64_t*>(P.keys)[P.first + i];
        else { key128 kk = reinterpret_cast<const key128*>(P.keys)[P.first + i]; hi = kk.hi; lo = kk.lo; }
        for (int j = 0; j < P.k; j++) {                       // first base most significant (UTIL:416-454)
            const int bit = 2 * (P.k - 1 - j);
            const unsigned int sym = (unsigned int)((bit >= 64 ? (hi >> (bit - 64)) : (lo >> bit)) & 3ull);
            o[j] = (uint8_t)"ACGT"[sym];
        }
        o[P.k] = (uint8_t)'\t';
        const int nd = dec_digits(c);
        uint32_t v = c;
        for (int j = nd - 1; j >= 0; j--) { o[P.k + 1 + j] = (uint8_t)('0' + v % 10u); v /= 10u; }
        o[P.k + 1 + nd] = (uint8_t)'\n';
    }
    __syncthreads();
    const unsigned long long base = P.tile_off[blockIdx.x];
    for (unsigned int j = threadIdx.x; j < total; j += 256) P.text[base + j] = s_text[j];
}
// text offset of entry bounds[b] (relative to `first`): its tile's offset plus the lines before it in the tile
template <bool WIDE>
__global__ void k_fmt_bounds(const FmtParams P) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.n_bounds) return;
    const unsigned long long e = P.bounds[b];                 // 0 <= e <= n
    const unsigned long long tile = e >> 8;
    unsigned long long off = P.tile_off[tile];
    for (unsigned long long i = tile << 8; i < e; i++) off += (unsigned)P.k + 2u + (unsigned)dec_digits(P.cnt[P.first + i]);
    P.bound_off[b] = off;
}

// ------------------------------------------------------------------ K8: multi-sample distances
// (MSKC:458-520, multiseq/SquaredEuclidean.java:19-27.)  Two per-bin sorted results A and B of the same
// configuration: acc += sum over k-mers present in both of cntA * cntB (a sparse dot product; the squared
// euclidean distance of the count vectors is dot(A,A) + dot(B,B) - 2 dot(A,B)).  One thread per entry of A,
// binary search inside the same bin of B.
struct DotParams {
    const void* keysA; const uint32_t* cntA; const unsigned long long* baseA; unsigned long long nA;
    const void* keysB; const uint32_t* cntB; const unsigned long long* baseB;
    int B; unsigned long long* acc;
};
template <bool WIDE>
__global__ void __launch_bounds__(256) k_sparse_dot(const DotParams P) {
    typedef typename Traits<WIDE>::Key Key;
    const Key* ka = reinterpret_cast<const Key*>(P.keysA); const Key* kb = reinterpret_cast<const Key*>(P.keysB);
    unsigned long long sum = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.nA;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const int bin = find_bin(P.baseA, 0, P.B, i);
        unsigned long long lo = P.baseB[bin], hi = P.baseB[bin + 1];
        const Key key = ka[i];
        while (lo < hi) {
            const unsigned long long mid = (lo + hi) >> 1;
            const Key km = kb[mid];
            bool less;
            if constexpr (!WIDE) less = km < key; else less = key_less(km, key);
            if (less) lo = mid + 1; else hi = mid;
        }
        if (lo < P.baseB[bin + 1] && key_eq(kb[lo], key)) sum += (unsigned long long)P.cntA[i] * (unsigned long long)P.cntB[lo];
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(P.acc, sum);
}

// ------------------------------------------------------------------ synthetic reads
struct SynthParams { SynthSpec S; uint64_t n_pos, n_words; uint64_t* bases; uint32_t* inv; };
// one thread per 32 positions; position p -> read p/(L+1), offset p%(L+1); offset L is the separator
__global__ void __launch_bounds__(256) k_synth(const SynthParams P) {
    uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= P.n_words) return;
    const uint64_t L = P.S.L;
    uint64_t p = wi << 5;
    uint64_t r = p / (L + 1), j = p - r * (L + 1);
    uint64_t pos = 0, strand = 0; bool have = false;
    uint64_t bw = 0; uint32_t iw = 0;
    for (int t = 0; t < 32; t++, p++) {
        uint32_t b = 0; bool bad = true;
        if (p < P.n_pos && j < L) {
            if (!have) { synth_read(P.S, P.S.first_read + r, pos, strand); have = true; }
            b = synth_base(P.S, P.S.first_read + r, j, pos, strand, bad);
            if (bad) b = 0;
        }
        bw = (bw << 2) | b; iw = (iw << 1) | (bad ? 1u : 0u);
        if (++j == L + 1) { j = 0; r++; have = false; }
    }
    P.bases[wi] = bw; P.inv[wi] = iw;
}

// one long sequence, positions [first_pos, first_pos + n_pos - 1) followed by the record separator
struct SynthLongParams { LongSpec S; uint64_t n_pos, n_words; uint64_t* bases; uint32_t* inv; };
__global__ void __launch_bounds__(256) k_synth_long(const SynthLongParams P) {
    uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= P.n_words) return;
    uint64_t bw = 0; uint32_t iw = 0;
    for (int t = 0; t < 32; t++) {
        const uint64_t p = (wi << 5) + t;
        uint32_t b = 0; bool bad = true;
        if (p + 1 < P.n_pos) { b = synth_long_base(P.S, P.S.first_pos + p, bad); if (bad) b = 0; }
        bw = (bw << 2) | b; iw = (iw << 1) | (bad ? 1u : 0u);
    }
    P.bases[wi] = bw; P.inv[wi] = iw;
}

}  // namespace fkm
