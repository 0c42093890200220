// fkm_ingest.cuh — FASTA text -> packed layout, on the device.
//
// Replaces the FASTdoop record readers + `getValue.replaceAll("\n","")` (SBKC:62-65,
// 1009-1012) for the in-memory path: the host only copies the raw text to the GPU.
// Record rule (same as the host packer fkm_pack_fasta): a record
// starts at a '>' that begins a line, its header runs to the end of that line, the
// value is every later byte up to the next header with '\n' dropped; bytes before
// the first header are ignored; one invalid separator position follows each record.
//
// Kernels (tile = 256 threads x 32 bytes):
//   k_ing_lines   last '\n' of each tile, position of the first header, header count
//   k_scan1<max>  last '\n' before each tile            (single-CTA scan over tiles)
//   k_ing_emit<0> kept positions per tile
//   k_scan1<sum>  first output position of each tile
//   k_ing_emit<1> classify again, stage 2-bit codes in shared memory, pack 32 per word
//   k_ing_finish  closing separator + invalid tail of the last word
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fkm {

static constexpr int kIngThreads = 256;
static constexpr int kIngBytesPerThread = 32;
static constexpr int kIngTile = kIngThreads * kIngBytesPerThread;   // 8192 bytes

struct IngestParams {
    const uint8_t* text; unsigned long long n;        // FASTA bytes on the device
    unsigned long long n_tiles;
    long long* tile_nl;                                // [n_tiles+1] last '\n' in tile -> (after scan) last '\n' before tile
    unsigned long long* tile_pos;                      // [n_tiles+1] kept per tile -> (after scan) first position of tile
    unsigned long long* first_hdr;                     // index of the first header byte (n if none)
    unsigned long long* n_hdr;                         // number of headers
    unsigned long long* bases; unsigned int* inv;     // zero-initialised output
};

__device__ __forceinline__ int nt_code_dev(uint8_t c) {   // UTIL:19-22; everything else is invalid (UTIL:697)
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}

__device__ __forceinline__ void load32(const uint8_t* text, unsigned long long n, unsigned long long i0, uint8_t (&b)[32]) {
    if (i0 + 32 <= n) {
        const uint4* p = reinterpret_cast<const uint4*>(text + i0);
        uint4 v0 = p[0], v1 = p[1];
        *reinterpret_cast<uint4*>(&b[0]) = v0; *reinterpret_cast<uint4*>(&b[16]) = v1;
    } else {
#pragma unroll
        for (int j = 0; j < 32; j++) b[j] = (i0 + j < n) ? text[i0 + j] : (uint8_t)'\n';
    }
}

__global__ void __launch_bounds__(kIngThreads) k_ing_lines(const IngestParams P) {
    __shared__ long long s_max[kIngThreads / 32];
    __shared__ unsigned long long s_min[kIngThreads / 32];
    __shared__ unsigned int s_cnt[kIngThreads / 32];
    const unsigned long long i0 = (unsigned long long)blockIdx.x * kIngTile + (unsigned long long)threadIdx.x * kIngBytesPerThread;
    long long last = -1; unsigned long long first = ~0ull; unsigned int nh = 0;
    if (i0 < P.n) {
        __align__(16) uint8_t b[32];
        load32(P.text, P.n, i0, b);
        uint8_t prev = (i0 == 0) ? (uint8_t)'\n' : P.text[i0 - 1];
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if (i0 + j < P.n) {
                if (b[j] == '\n') last = (long long)(i0 + j);
                if (b[j] == '>' && prev == '\n') { nh++; if (first == ~0ull) first = i0 + j; }
            }
            prev = b[j];
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, d));
        first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, d));
        nh += __shfl_xor_sync(0xFFFFFFFFu, nh, d);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_max[warp] = last; s_min[warp] = first; s_cnt[warp] = nh; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kIngThreads / 32; w++) { last = max(last, s_max[w]); first = min(first, s_min[w]); nh += s_cnt[w]; }
        P.tile_nl[blockIdx.x] = last;
        if (first != ~0ull) atomicMin(P.first_hdr, first);
        if (nh) atomicAdd(P.n_hdr, (unsigned long long)nh);
    }
}

// exclusive scan of a[0..n) in place by one CTA; a[n] receives the total.  OP 0: sum (u64), OP 1: max (i64, identity -1)
template <int OP>
__global__ void __launch_bounds__(1024) k_scan1(long long* a, unsigned long long n) {
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ident = OP ? -1ll : 0ll;
    auto op = [](long long x, long long y) { return OP ? (x > y ? x : y) : (long long)((unsigned long long)x + (unsigned long long)y); };
    if (threadIdx.x == 0) s_carry = ident;
    __syncthreads();
    for (unsigned long long base = 0; base < n; base += 1024) {
        unsigned long long i = base + threadIdx.x;
        long long v = (i < n) ? a[i] : ident, incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl = op(incl, t); }
        long long excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1); if (lane == 0) excl = ident;
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        long long wb = ident, tot = ident;
        for (int j = 0; j < 32; j++) { long long t = s_w[j]; if (j < warp) wb = op(wb, t); tot = op(tot, t); }
        long long carry = s_carry;
        if (i < n) a[i] = op(carry, op(wb, excl));
        __syncthreads();
        if (threadIdx.x == 0) s_carry = op(carry, tot);
        __syncthreads();
    }
    if (threadIdx.x == 0) a[n] = s_carry;
}

// WRITE 0: count kept positions per tile.  WRITE 1: emit them.
template <int WRITE>
__global__ void __launch_bounds__(kIngThreads) k_ing_emit(const IngestParams P) {
    __shared__ long long s_wl[kIngThreads / 32];
    __shared__ unsigned int s_wc[kIngThreads / 32];
    __shared__ uint8_t s_code[WRITE ? kIngTile : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long tile0 = (unsigned long long)blockIdx.x * kIngTile;
    const unsigned long long i0 = tile0 + (unsigned long long)threadIdx.x * kIngBytesPerThread;
    const unsigned long long first_hdr = *P.first_hdr;
    __align__(16) uint8_t b[32];
    long long last = -1;
    if (i0 < P.n) {
        load32(P.text, P.n, i0, b);
#pragma unroll
        for (int j = 0; j < 32; j++) if (i0 + j < P.n && b[j] == '\n') last = (long long)(i0 + j);
    }
    // exclusive max-scan of `last` over the block -> last '\n' before this thread's bytes
    long long incl = last;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl = max(incl, t); }
    long long prev_nl = __shfl_up_sync(0xFFFFFFFFu, incl, 1); if (lane == 0) prev_nl = -1;
    if (lane == 31) s_wl[warp] = incl;
    __syncthreads();
    for (int w = 0; w < warp; w++) prev_nl = max(prev_nl, s_wl[w]);
    prev_nl = max(prev_nl, P.tile_nl[blockIdx.x]);                 // carry from earlier tiles
    // classify
    unsigned int kept = 0; unsigned long long codes_lo = 0, codes_hi = 0;   // 3-bit code per byte: 0..3 base, 4 invalid, 7 dropped
    if (i0 < P.n) {
        const unsigned long long ls = (unsigned long long)(prev_nl + 1);    // start of the line holding byte i0
        bool in_hdr = (ls < i0) ? (P.text[ls] == '>') : false;              // ls == i0: decided below at the line start
        bool bol = (ls == i0);
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const unsigned long long i = i0 + j;
            unsigned int code = 7;
            if (i < P.n && i >= first_hdr) {
                const uint8_t c = b[j];
                if (bol) in_hdr = (c == '>');
                if (in_hdr) { if (bol && i != first_hdr) code = 4; }        // separator closing the previous record
                else if (c != '\n') code = (unsigned)nt_code_dev(c);
                bol = (c == '\n');
            } else if (i < P.n) {
                bol = (b[j] == '\n');
            }
            if (code != 7) kept++;
            if (j < 16) codes_lo |= (unsigned long long)code << (3 * j); else codes_hi |= (unsigned long long)code << (3 * (j - 16));
        }
    }
    // exclusive sum-scan of kept over the block
    unsigned int inc = kept;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned int t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_wc[warp] = inc;
    __syncthreads();
    unsigned int wbase = 0, total = 0;
    for (int w = 0; w < kIngThreads / 32; w++) { unsigned int t = s_wc[w]; if (w < warp) wbase += t; total += t; }
    if (WRITE == 0) {
        if (threadIdx.x == 0) P.tile_pos[blockIdx.x] = total;
        return;
    }
    unsigned int o = wbase + inc - kept;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        unsigned int code = (unsigned int)(((j < 16) ? (codes_lo >> (3 * j)) : (codes_hi >> (3 * (j - 16)))) & 7ull);
        if (code != 7) s_code[o++] = (uint8_t)code;
    }
    __syncthreads();
    if (total == 0) return;
    const unsigned long long pos0 = P.tile_pos[blockIdx.x], pos1 = pos0 + total;
    const unsigned long long w0 = pos0 >> 5, w1 = (pos1 - 1) >> 5;
    for (unsigned long long w = w0 + threadIdx.x; w <= w1; w += kIngThreads) {
        unsigned long long bw = 0; unsigned int iw = 0;
        const unsigned long long p0 = w << 5;
#pragma unroll 8
        for (int t = 0; t < 32; t++) {
            const unsigned long long p = p0 + t;
            unsigned int code = 0;
            if (p >= pos0 && p < pos1) code = s_code[p - pos0];
            bw = (bw << 2) | (code & 3u); iw = (iw << 1) | (code >> 2);
        }
        if (p0 >= pos0 && p0 + 32 <= pos1) { P.bases[w] = bw; P.inv[w] = iw; }
        else { if (bw) atomicOr(&P.bases[w], bw); if (iw) atomicOr(&P.inv[w], iw); }
    }
}

// n_kept = tile_pos[n_tiles]; when at least one record exists the last record's
// separator is appended; then the tail of the last word is marked invalid.
__global__ void k_ing_finish(const IngestParams P, unsigned long long* out_npos, unsigned long long* out_nbases) {
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long kept = P.tile_pos[P.n_tiles], nh = *P.n_hdr;
    unsigned long long n_pos = kept + (nh ? 1 : 0);
    if (nh) atomicOr(&P.inv[(n_pos - 1) >> 5], 1u << (31 - ((n_pos - 1) & 31)));
    if (n_pos & 31) atomicOr(&P.inv[n_pos >> 5], 0xFFFFFFFFu >> (n_pos & 31));
    *out_npos = n_pos;
    *out_nbases = kept - (nh ? nh - 1 : 0);
}

}  // namespace fkm
