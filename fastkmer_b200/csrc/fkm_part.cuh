// fkm_part.cuh — the partitioned count path (useHT = 1, count_mode = 2): hash tables that live in one SM's shared
// memory, fed by a second partition level over the canonical k-mers themselves.
//
// Replaces extractKXmersHT (SBKC:664-739: one Object2IntOpenHashMap per bin, addTo per k-mer, dump of the entries).
// A bin of the reference holds millions of k-mers; one SM's shared memory holds a table of 16 K slots.  So every bin
// (its super-k-mer records are already bin-major, the output of the "shuffle" SBKC:1035) is cut into P sub-buckets BY
// A HASH OF THE CANONICAL K-MER — all occurrences of a k-mer meet in one sub-bucket whatever read, strand or record
// they came from, and the sub-buckets of a bin are equally full whatever the minimizer statistics are:
//
//   k_expand_hist   records -> canonical k-mers, written once as plain keys (bin-major), + per-(tile, sub-bucket) histogram
//                   (tile = 512 consecutive records)
//   k_sub_scan      per bin: sub-bucket sizes and offsets, and where every tile's share of a sub-bucket begins
//   k_place_keys    every k-mer gets its exact place: a tile's keys are put in sub-bucket order in shared memory first,
//                   so that they leave in contiguous runs
//   k_count_keys    persistent CTAs, one sub-bucket at a time: its k-mers (plain 8- or 16-byte keys, contiguous) are
//                   inserted into a table in SHARED memory (LDS + ATOMS.CAS on the key, ATOMS.ADD on the count) and the
//                   distinct (k-mer, count) pairs leave once, densely, into the CTA's own region of the output.
//
// No table ever touches L2 or DRAM; DRAM traffic is records (read once) + keys (written twice, read twice) + result.
// A sub-bucket with more distinct k-mers than the table takes (planning is from an estimated distinct / k-mer ratio) is
// redone by the same CTA in a private global-memory table; if that overflows too a flag sends the job to the
// global-table pipeline of fkm_lib.cu.
#pragma once
#include "fkm_smem.cuh"

namespace fkm {

static constexpr int kPartMaxSubs = 2048;                    // sub-buckets per bin (shared-memory histograms of the partition kernels)
template <bool WIDE> struct PartGeom {
    static constexpr int kThreads = WIDE ? 256 : 512;        // threads of k_expand_hist / k_place_keys
    static constexpr int kWarps = kThreads / 32;
    static constexpr int kTileRecs = kThreads;               // records per tile: one per thread
    static constexpr int kBufKeys = WIDE ? 4096 : 5632;      // keys of a tile that are put in order in shared memory (the rest is stored directly)
};
// where bin b (batch-relative) begins in keys_lin: every tile's share is padded to an even number of keys, every bin begins at an even index
__device__ __forceinline__ unsigned long long part_lin_base(unsigned long long key_base_b, uint32_t tile_first_b, int b) {
    return (key_base_b + tile_first_b + (unsigned long long)b + 1ull) & ~1ull;
}

struct PartParams {
    const void* records;                        // bin-major super-k-mer records
    const unsigned long long* bin_rec_base;     // [B+1] record offset of every bin (global bin ids)
    const unsigned long long* bin_rec_cnt;      // [B] records of every bin when the bins' regions have gaps (speculative scatter), else NULL
    int bin_lo, bin_hi;                         // bins of this batch; every array below is indexed by b - bin_lo
    const uint32_t* tile_first;                 // [nb+1] first tile of the bin
    const uint32_t* sub_first;                  // [nb+1] first sub-bucket of the bin
    const unsigned long long* hist_off;         // [nb] where the bin's [tiles][subs] matrices begin in tile_hist / tile_base
    const unsigned long long* key_base;         // [nb+1] first key of the bin in `keys`
    uint32_t n_tiles, n_sub;
    uint32_t* tile_hist;                        // k-mers of (tile, sub-bucket)
    uint32_t* tile_base;                        // first key of (tile, sub-bucket), relative to the bin's first key
    uint32_t* bin_key_cursor;                   // [nb] keys of the bin written so far by k_expand_hist (zeroed)
    unsigned long long* tile_key_off; uint32_t* tile_nkeys;   // [n_tiles] where the tile's k-mers sit in keys_lin (even), and how many
    void* keys_lin;                             // the batch's canonical k-mers, bin-major, tile by tile
    void* keys;                                 // ... and sub-bucket-major
    unsigned long long* mid_key_base;           // [n_sub+1] first key of every sub-bucket
    uint32_t* mid_bin;                          // [n_sub] its bin (global id)
    int k;
};

// The warp's 32 records sit in shared memory at byte address rec_smem (16 or 32 bytes each); nk is the k-mer count of this
// lane's record (0 beyond the end).  The warp's k-mers form a pool; lane l takes k-mers l, l + 32, ... of it, so every
// lane is busy whatever the lengths of the records.  F(key, active) is called by the converged warp once per 32 k-mers.
template <bool WIDE, typename F>
__device__ __forceinline__ void warp_each_kmer(uint32_t rec_smem, uint32_t nk, int k, F&& f) {
    typedef typename SmTraits<WIDE>::Key Key;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t RBY = WIDE ? 32u : 16u;
    const uint32_t le_mask = (lane == 31) ? 0xFFFFFFFFu : ((2u << lane) - 1u);
    const RecShared recs{rec_smem};
    uint32_t incl = nk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
    const uint32_t excl = incl - nk;
    const uint32_t T = __shfl_sync(FULL, incl, 31);
    uint32_t n_before = 0;                                   // records that start before the current block of 32 k-mers
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
            if constexpr (!WIDE) key = kmer_of_record(recs, (uint32_t)ri * RBY, j, k); else key = kmer_of_record_wide(recs, (uint32_t)ri * RBY, j, k);
        }
        f(key, active);
    }
}

__device__ __forceinline__ uint64_t kc_load(const uint64_t* p) { return __ldcs(reinterpret_cast<const unsigned long long*>(p)); }
__device__ __forceinline__ key128 kc_load(const key128* p) { const ulonglong2 q = __ldcs(reinterpret_cast<const ulonglong2*>(p)); key128 r; r.lo = q.x; r.hi = q.y; return r; }

// asks the bulk-copy engine to bring [p, p + bytes) into L2 (both multiples of 16)
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Sub-bucket of a k-mer among the Pn sub-buckets of its bin.  ORDERED = false (hash path): by hash.  ORDERED = true (sort path):
// Pn is a power of two and the sub-bucket is the k-mer's top log2(Pn) bits, so the sub-buckets of a bin are key ranges in order.
template <bool ORDERED> __device__ __forceinline__ uint32_t part_sub(uint64_t key, uint32_t Pn, int k) {
    if constexpr (!ORDERED) return umulhi32(part_hash(key), Pn);
    else { const uint32_t top = (uint32_t)((key << (64 - 2 * k)) >> 32); return Pn > 1u ? top >> (__clz(Pn) + 1) : 0u; }
}
template <bool ORDERED> __device__ __forceinline__ uint32_t part_sub(key128 key, uint32_t Pn, int k) {
    if constexpr (!ORDERED) return umulhi32(part_hash(key), Pn);
    else { const int s = 128 - 2 * k; const uint32_t top = (uint32_t)(((s ? ((key.hi << s) | (key.lo >> (64 - s))) : key.hi)) >> 32); return Pn > 1u ? top >> (__clz(Pn) + 1) : 0u; }
}

// tile -> bin of the batch (largest b with tile_first[b] <= tile), by thread 0 into shared memory
__device__ __forceinline__ int part_find_bin(const uint32_t* tile_first, int nb, uint32_t tile) {
    int lo = 0, hi = nb;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tile_first[mid] <= tile) lo = mid; else hi = mid; }
    return lo;
}

// loads this lane's record of the tile into the warp's staging area; returns its k-mer count
template <bool WIDE>
__device__ __forceinline__ uint32_t part_stage_record(const void* records, unsigned long long r, bool in, unsigned char* stage_lane) {
    uint32_t nk = 0;
    if (in) {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(records) + (WIDE ? 2 : 1) * r;
        const ulonglong2 a = __ldcs(src);
        *reinterpret_cast<ulonglong2*>(stage_lane) = a;
        if constexpr (WIDE) {
            const ulonglong2 b2 = __ldcs(src + 1);
            *reinterpret_cast<ulonglong2*>(stage_lane + 16) = b2;
            nk = (uint32_t)(b2.y & 0xFFull);
        } else nk = (uint32_t)(a.y & 0xFFull);
    }
    return nk;
}

// ------------------------------------------------------------------ pass 1: records -> canonical k-mers (bin-major) + per-(tile, sub-bucket) histogram
// One tile = kTileRecs consecutive records of a bin, one record per thread.  The tile's k-mers are written contiguously
// (a per-bin cursor hands out the place: the order of the tiles inside a bin is whatever the scheduling makes it, which
// nobody observes) in the order of the warps' k-mer pools, so every store instruction writes 32 consecutive keys.
template <bool WIDE, bool ORDERED>
__global__ void __launch_bounds__(PartGeom<WIDE>::kThreads) k_expand_hist(const PartParams P) {
    typedef typename SmTraits<WIDE>::Key Key;
    constexpr int kPartThreads = PartGeom<WIDE>::kThreads, kPartWarps = PartGeom<WIDE>::kWarps;
    constexpr int RB = WIDE ? 32 : 16;
    constexpr int TR = PartGeom<WIDE>::kTileRecs;
    __shared__ uint32_t s_hist[kPartMaxSubs];
    __shared__ __align__(16) unsigned char s_rec[kPartWarps][32 * RB];
    __shared__ uint32_t s_wtot[kPartWarps];
    __shared__ uint32_t s_key0;
    __shared__ int s_bin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = P.bin_hi - P.bin_lo;
    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        if (threadIdx.x == 0) s_bin = part_find_bin(P.tile_first, nb, tile);
        __syncthreads();
        const int b = s_bin;
        const uint32_t Pn = P.sub_first[b + 1] - P.sub_first[b];
        const uint32_t tl = tile - P.tile_first[b];
        const unsigned long long r_lo = P.bin_rec_base[P.bin_lo + b] + (unsigned long long)tl * TR;
        const unsigned long long r_end = P.bin_rec_cnt ? P.bin_rec_base[P.bin_lo + b] + P.bin_rec_cnt[P.bin_lo + b] : P.bin_rec_base[P.bin_lo + b + 1];
        const unsigned long long r_hi = min(r_lo + (unsigned long long)TR, r_end);
        for (uint32_t i = threadIdx.x; i < Pn; i += kPartThreads) s_hist[i] = 0u;
        const unsigned long long r = r_lo + threadIdx.x;
        const uint32_t nk = part_stage_record<WIDE>(P.records, r, r < r_hi, &s_rec[warp][lane * RB]);
        const uint32_t wt = __reduce_add_sync(0xFFFFFFFFu, nk);
        if (lane == 0) s_wtot[warp] = wt;
        __syncthreads();
        // exclusive prefix of the warps' k-mer counts: one shuffle scan over the (at most 16) totals
        const uint32_t wv = lane < kPartWarps ? s_wtot[lane] : 0u;
        uint32_t winc = wv;
#pragma unroll
        for (int d = 1; d < kPartWarps; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, d); if (lane >= d) winc += t; }
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, winc, kPartWarps - 1);
        const uint32_t wb = __shfl_sync(0xFFFFFFFFu, winc - wv, warp);
        if (threadIdx.x == 0) {
            const uint32_t k0 = atomicAdd(&P.bin_key_cursor[b], (tot + 1u) & ~1u);
            s_key0 = k0; P.tile_key_off[tile] = part_lin_base(P.key_base[b], P.tile_first[b], b) + k0; P.tile_nkeys[tile] = tot;
        }
        __syncthreads();
        Key* const out = reinterpret_cast<Key*>(P.keys_lin) + part_lin_base(P.key_base[b], P.tile_first[b], b) + s_key0 + wb;
        uint32_t t0 = 0;
        warp_each_kmer<WIDE>(smem_addr(&s_rec[warp][0]), nk, P.k, [&](Key key, bool active) {
            if (active) {
                out[t0 + lane] = key;
                atomicAdd(&s_hist[part_sub<ORDERED>(key, Pn, P.k)], 1u);
            }
            t0 += 32;
        });
        __syncthreads();
        uint32_t* dst = P.tile_hist + P.hist_off[b] + (unsigned long long)tl * Pn;
        for (uint32_t i = threadIdx.x; i < Pn; i += kPartThreads) dst[i] = s_hist[i];
        __syncthreads();                                     // s_bin, s_hist, s_rec are reused by the next tile
    }
}

// ------------------------------------------------------------------ pass 2: offsets.  One CTA per bin of the batch.
__global__ void __launch_bounds__(256) k_sub_scan(const PartParams P) {
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = P.bin_hi - P.bin_lo;
    const int b = blockIdx.x;
    if (b == nb - 1 && threadIdx.x == 0) P.mid_key_base[P.n_sub] = P.key_base[nb];
    const uint32_t Pn = P.sub_first[b + 1] - P.sub_first[b];
    const uint32_t nt = P.tile_first[b + 1] - P.tile_first[b];
    if (Pn == 0) return;
    const uint32_t* hist = P.tile_hist + P.hist_off[b];
    uint32_t* base = P.tile_base + P.hist_off[b];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t s0 = 0; s0 < Pn; s0 += 256) {
        const uint32_t s = s0 + threadIdx.x;
        uint32_t total = 0;
        if (s < Pn) {
            uint32_t t = 0;
            for (; t + 4 <= nt; t += 4) {
                const uint32_t a0 = hist[(size_t)t * Pn + s], a1 = hist[(size_t)(t + 1) * Pn + s], a2 = hist[(size_t)(t + 2) * Pn + s], a3 = hist[(size_t)(t + 3) * Pn + s];
                total += a0 + a1 + a2 + a3;
            }
            for (; t < nt; t++) total += hist[(size_t)t * Pn + s];
        }
        uint32_t incl = total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint32_t wb = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { const uint32_t t = s_w[i]; if (i < warp) wb += t; tot += t; }
        const uint32_t excl = s_carry + wb + incl - total;
        if (s < Pn) {
            P.mid_key_base[P.sub_first[b] + s] = P.key_base[b] + (unsigned long long)excl;
            P.mid_bin[P.sub_first[b] + s] = (uint32_t)(P.bin_lo + b);
            uint32_t run = excl;
            for (uint32_t t = 0; t < nt; t++) { const uint32_t v = hist[(size_t)t * Pn + s]; base[(size_t)t * Pn + s] = run; run += v; }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ pass 3: k-mers -> their sub-buckets
// dynamic shared memory: keys[kBufKeys] | sub[kBufKeys] (u16).  A tile's k-mers (plain keys, contiguous in keys_lin; the
// bulk-copy engine pulls the CTA's next tile into L2 while the current one is worked on) are put in sub-bucket order in
// shared memory — the place of every k-mer is exact: the tile's share of a sub-bucket begins at tile_base, ATOMS.ADD
// hands out the ranks — and leave in contiguous runs.  Three CTAs per SM, two barriers per phase change.
template <bool WIDE, bool ORDERED>
__global__ void __launch_bounds__(PartGeom<WIDE>::kThreads) k_place_keys(const PartParams P) {
    typedef typename SmTraits<WIDE>::Key Key;
    constexpr int kPartThreads = PartGeom<WIDE>::kThreads, kPartWarps = PartGeom<WIDE>::kWarps;
    constexpr uint32_t BUF = PartGeom<WIDE>::kBufKeys;
    extern __shared__ __align__(128) unsigned char part_raw[];
    Key* s_keys = reinterpret_cast<Key*>(part_raw);
    unsigned short* s_sub = reinterpret_cast<unsigned short*>(s_keys + BUF);
    __shared__ uint32_t s_cur[kPartMaxSubs];                 // next place (in the tile's sub-bucket order) of every sub-bucket
    __shared__ uint32_t s_delta[kPartMaxSubs];               // key index (relative to the bin) minus that place
    __shared__ uint32_t s_w[2][kPartWarps];
    __shared__ int s_bin[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = P.bin_hi - P.bin_lo;
    const Key* const lin = reinterpret_cast<const Key*>(P.keys_lin);
    auto prefetch = [&](uint32_t tile) {                     // thread 0: the tile's keys -> L2
        const uint32_t bytes = (uint32_t)((((size_t)P.tile_nkeys[tile] * sizeof(Key)) + 15u) & ~15u);
        if (bytes) bulk_prefetch_l2(lin + P.tile_key_off[tile], bytes);
    };
    if (threadIdx.x == 0 && blockIdx.x < P.n_tiles) s_bin[0] = part_find_bin(P.tile_first, nb, blockIdx.x);
    __syncthreads();
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, it++) {
        const int b = s_bin[it & 1];
        if (threadIdx.x == 0 && tile + gridDim.x < P.n_tiles) {      // the next tile: its bin, and its keys on their way to L2
            s_bin[(it + 1) & 1] = part_find_bin(P.tile_first, nb, tile + gridDim.x);
            prefetch(tile + gridDim.x);
        }
        const uint32_t Pn = P.sub_first[b + 1] - P.sub_first[b];
        const uint32_t tl = tile - P.tile_first[b];
        const uint32_t* hist = P.tile_hist + P.hist_off[b] + (unsigned long long)tl * Pn;
        const uint32_t* base = P.tile_base + P.hist_off[b] + (unsigned long long)tl * Pn;
        const uint32_t n_tile = P.tile_nkeys[tile];
        const Key* const in = lin + P.tile_key_off[tile];
        // the first keys are requested before the offsets are built
        Key k0[2]; bool h0[2];
#pragma unroll
        for (int j = 0; j < 2; j++) { const uint32_t i = j * kPartThreads + threadIdx.x; h0[j] = i < n_tile; if (h0[j]) k0[j] = kc_load(in + i); }
        // exclusive scan of the tile's counts over its sub-buckets: the tile's own sub-bucket order
        uint32_t carry = 0;
        for (uint32_t s0 = 0, r = 0; s0 < Pn; s0 += kPartThreads, r ^= 1u) {
            const uint32_t sb = s0 + threadIdx.x;
            const uint32_t c = sb < Pn ? hist[sb] : 0u;
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
            if (lane == 31) s_w[r][warp] = incl;
            __syncthreads();
            uint32_t wb = 0, tot = 0;
#pragma unroll
            for (int i = 0; i < kPartWarps; i++) { const uint32_t t = s_w[r][i]; if (i < warp) wb += t; tot += t; }
            const uint32_t excl = carry + wb + incl - c;
            if (sb < Pn) { s_cur[sb] = excl; s_delta[sb] = base[sb] - excl; }
            carry += tot;
        }
        __syncthreads();
        Key* const out = reinterpret_cast<Key*>(P.keys) + P.key_base[b];
        auto place = [&](Key key) {
            const uint32_t sub = part_sub<ORDERED>(key, Pn, P.k);
            const uint32_t pos = atomicAdd(&s_cur[sub], 1u);
            if (pos < BUF) { s_keys[pos] = key; s_sub[pos] = (unsigned short)sub; }
            else out[(uint32_t)(s_delta[sub] + pos)] = key;                  // a tile with unusually many k-mers: the tail is stored directly
        };
#pragma unroll
        for (int j = 0; j < 2; j++) if (h0[j]) place(k0[j]);
        for (uint32_t i0 = 2 * kPartThreads; i0 < n_tile; i0 += 4 * kPartThreads) {
            Key kk[4]; bool hv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { const uint32_t i = i0 + j * kPartThreads + threadIdx.x; hv[j] = i < n_tile; if (hv[j]) kk[j] = kc_load(in + i); }
#pragma unroll
            for (int j = 0; j < 4; j++) if (hv[j]) place(kk[j]);
        }
        __syncthreads();
        const uint32_t n_buf = min(n_tile, BUF);
        for (uint32_t pos = threadIdx.x; pos < n_buf; pos += kPartThreads) out[(uint32_t)(s_delta[s_sub[pos]] + pos)] = s_keys[pos];
        __syncthreads();                                     // s_keys, s_cur, s_delta, s_bin are reused by the next tile
    }
}

// ------------------------------------------------------------------ sort path: sub-buckets (key ranges in order) -> chunks for k_radix_local
// One thread per bin: consecutive sub-buckets are merged into chunks of at most `cap` keys; a single sub-bucket above cap sets *too_big.
__global__ void k_form_chunks_sub(const unsigned long long* mid_key_base, const uint32_t* sub_first, int nb, unsigned int cap,
                                  ChunkDesc* chunks, unsigned int max_chunks, unsigned int* n_chunks, int* too_big) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    unsigned long long start = mid_key_base[sub_first[b]];
    unsigned int acc = 0;
    auto emit = [&](unsigned int n) {
        if (!n) return;
        const unsigned int c = atomicAdd(n_chunks, 1u);
        if (c < max_chunks) { chunks[c].start = start; chunks[c].n = n; chunks[c].pad = 0; } else *too_big = 1;
        start += n;
    };
    for (uint32_t m = sub_first[b]; m < sub_first[b + 1]; m++) {
        const unsigned long long c64 = mid_key_base[m + 1] - mid_key_base[m];
        if (c64 > cap) { *too_big = 1; return; }
        const unsigned int c = (unsigned int)c64;
        if (acc + c > cap) { emit(acc); acc = 0; }
        acc += c;
    }
    emit(acc);
}

// ------------------------------------------------------------------ the count kernel
struct KeyCountParams {
    const void* keys;                           // sub-bucket-major canonical k-mers of the batch
    const unsigned long long* mid_key_base;     // [n_sub+1]
    const uint32_t* mid_bin;                    // [n_sub] bin of every sub-bucket
    const uint32_t* sub_first;                  // [nb+1] first sub-bucket of every bin of the batch
    int bin_lo;
    uint32_t n_sub;
    void* out_keys; uint32_t* out_cnt;          // CTA c writes its entries densely from index c * region_cap
    unsigned long long region_cap;
    unsigned long long* cta_total;              // [gridDim.x] entries every CTA wrote
    uint32_t* bin_cta; unsigned long long* bin_off;   // [B] where a bin's entries begin: CTA and offset inside its region
    unsigned long long* acc;                    // [3][64] digest accumulators (sum, xor, count)
    uint32_t cap_slots;                         // slots of the shared-memory table (a power of two)
    uint32_t max_fill;                          // distinct k-mers it may take
    void* slow_keys; uint32_t* slow_cnt;        // private global table of every CTA (slow path)
    unsigned long long slow_slots; unsigned long long slow_max_fill;
    int* flags;                                 // [0] slow table overflow (job must be redone), [1] output region too small, [2] a 32-bit count wrapped
    unsigned long long* counters;               // [0] sub-buckets that took the slow path
    int k;                                      // k_count_keys_ordered: k-mer length (the slot is cut from the k-mer's top bits)
    uint32_t tail_slots;                        // k_count_keys_ordered: slots behind cap_slots (its probing never wraps around)
    int bin_shift;                              // the digest's bin id = internal bin >> bin_shift
};

static constexpr int kKcThreads = 1024;
static constexpr uint32_t kKcMaxProbe = 192;                 // probes before the shared-memory table counts as full


__device__ __forceinline__ uint64_t kc_lds(const uint64_t* p) {
    uint64_t r;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(r) : "r"(smem_addr(p)) : "memory");
    return r;
}
__device__ __forceinline__ key128 kc_lds(const key128* p) {
    key128 r;
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "r"(smem_addr(p)) : "memory");
    return r;
}
// one probe of a shared-memory slot.  hit: the key is (now) in the slot; returns true when this call put it there
__device__ __forceinline__ bool kc_probe(uint64_t* slot, uint64_t key, bool& hit) {
    const uint64_t cur = kc_lds(slot);
    hit = cur == key;
    if (cur != ~0ull) return false;
    const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(slot), ~0ull, (unsigned long long)key);
    hit = old == ~0ull || old == key;
    return old == ~0ull;
}
__device__ __forceinline__ bool kc_probe(key128* slot, key128 key, bool& hit) {
    key128 cur = kc_lds(slot);
    hit = key_eq(cur, key);
    if (cur.lo != ~0ull && cur.hi != ~0ull) return false;
    // (a 16-byte shared-memory load of an aligned slot is not torn, but a half equal to all ones is handed to the CAS anyway)
    const key128 empty = sm_empty(key);
    cur = atoms_cas128(slot, empty, key);
    const bool won = sm_is_empty(cur);
    hit = won || key_eq(cur, key);
    return won;
}

// dynamic shared memory: keys[cap_slots] | cnt[cap_slots] | list[cap_slots] (u16: the claimed slots in claim order)
//
// Every CTA owns a contiguous range of sub-buckets (equal shares of the k-mers), hence a contiguous range of keys and a
// contiguous region of the output: no CTA ever waits for another.  Inside a sub-bucket thread t takes keys t, t + 1024,
// ... (every warp load is one contiguous 256- or 512-byte piece; two loads per thread are in flight), and a thread whose
// key is done moves on to its next one while its neighbours still probe: the warp makes one probe per lane per round
// whatever the lengths of the probe chains.
template <bool WIDE, int NT, bool GUARD>
__global__ void __launch_bounds__(NT, 1024 / NT) k_count_keys(const KeyCountParams P) {
    typedef typename SmTraits<WIDE>::Key Key;
    extern __shared__ __align__(128) unsigned char kc_raw[];
    Key* const t_keys = reinterpret_cast<Key*>(kc_raw);
    uint32_t* const t_cnt = reinterpret_cast<uint32_t*>(t_keys + P.cap_slots);
    unsigned short* const t_list = reinterpret_cast<unsigned short*>(t_cnt + P.cap_slots);
    __shared__ unsigned int s_nclaim, s_pos, s_ovf;
    __shared__ unsigned int s_range[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Key empty = sm_empty(Key());
    const uint32_t mask = P.cap_slots - 1u;
    for (uint32_t i = threadIdx.x; i < P.cap_slots; i += NT) { t_keys[i] = empty; t_cnt[i] = 0; }
    if (threadIdx.x == 0) {
        s_nclaim = 0; s_pos = 0; s_ovf = 0;
        // this CTA's sub-buckets: [first whose key offset >= c * total / G, the same for c + 1)
        const unsigned long long total = P.mid_key_base[P.n_sub];
        for (int e = 0; e < 2; e++) {
            const unsigned long long want = (total / gridDim.x) * (blockIdx.x + e) + min((unsigned long long)(blockIdx.x + e), total % gridDim.x);
            uint32_t lo = 0, hi = P.n_sub;
            if (blockIdx.x + e >= gridDim.x) lo = P.n_sub;
            else while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (P.mid_key_base[mid] >= want) hi = mid; else lo = mid + 1; }
            s_range[e] = lo;
        }
    }
    __syncthreads();
    const uint32_t m_lo = s_range[0], m_hi = s_range[1];
    Key* const okeys = reinterpret_cast<Key*>(P.out_keys) + (size_t)blockIdx.x * P.region_cap;
    uint32_t* const ocnt = P.out_cnt + (size_t)blockIdx.x * P.region_cap;
    const Key* const keys = reinterpret_cast<const Key*>(P.keys);
    unsigned long long off = 0;                              // entries this CTA has written (uniform)
    unsigned long long dsum = 0, dxor = 0, dcnt = 0;
    bool slow_ready = false, out_full = false;
    int wrapped = 0;

    for (uint32_t m = m_lo; m < m_hi; m++) {
        const uint32_t bin = P.mid_bin[m];
        if (threadIdx.x == 0 && P.sub_first[bin - (uint32_t)P.bin_lo] == m) { P.bin_cta[bin] = blockIdx.x; P.bin_off[bin] = off; }
        const unsigned long long kb = P.mid_key_base[m], ke = P.mid_key_base[m + 1];
        if (ke == kb) continue;
        const Key* const kp = keys + kb;
        const unsigned long long K = ke - kb;
        // ---- insert (one key at work, the next two on their way; the sub-bucket after this one is pulled into L2 meanwhile)
        if (threadIdx.x == 0 && m + 1 < m_hi) {
            const unsigned long long nb0 = (ke * sizeof(Key)) & ~15ull, nb1 = (P.mid_key_base[m + 2] * sizeof(Key) + 15ull) & ~15ull;
            if (nb1 > nb0) bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(keys) + nb0, (uint32_t)min(nb1 - nb0, (unsigned long long)(1u << 20)));
        }
        {
            const uint32_t K32 = (uint32_t)K;               // (a bin, hence a sub-bucket, has fewer than 2^32 k-mers)
            uint32_t p = threadIdx.x;
            bool have = p < K32; Key key = Key(); if (have) key = kc_load(kp + p); p += NT;
            bool hn = p < K32; Key nxt = Key(); if (hn) nxt = kc_load(kp + p); p += NT;
            bool hn2 = p < K32; Key nxt2 = Key(); if (hn2) nxt2 = kc_load(kp + p); p += NT;     // two loads in flight behind the key at work
            // The table never fills up: a lane whose claim is number max_fill or later (max_fill = slots - 2 * threads: every thread
            // has at most one claim in flight) raises the overflow flag and stops, so a probe sequence always meets its key or an
            // empty slot and the miss path is one increment.
            // (Tables of fewer than 4096 slots — test knob — can fill up under 1024 concurrent claims: GUARD counts the probes.)
            uint32_t slot = part_slot(part_hash(key), mask), probes = 0;
            while (have) {
                bool hit;
                if (kc_probe(&t_keys[slot], key, hit)) {     // this lane claimed the slot: it joins the list the dump walks
                    unsigned int pos;                        // (plain per-lane ATOMS: the compiler's warp-aggregated form costs 15 more instructions a round)
                    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(smem_addr(&s_nclaim)) : "memory");
                    t_list[pos] = (unsigned short)slot;
                    if (pos >= P.max_fill) { s_ovf = 1; have = false; hit = false; }
                }
                if (hit) {
                    atomicAdd(&t_cnt[slot], 1u);             // (cannot wrap: a sub-bucket has fewer than 2^32 k-mers)
                    have = hn; key = nxt; hn = hn2; nxt = nxt2; hn2 = p < K32; if (hn2) nxt2 = kc_load(kp + p); p += NT;
                    slot = part_slot(part_hash(key), mask); probes = 0;
                } else {
                    slot = (slot + 1u) & mask;
                    if constexpr (GUARD) { if (++probes > kKcMaxProbe) { s_ovf = 1; have = false; } }
                }
            }
        }
        __syncthreads();                                     // all inserts of the sub-bucket are done
        const bool slow = s_ovf != 0u;                      // (a table fuller than planned but not full is only slower)
        Key* gkeys = nullptr; uint32_t* gcnt = nullptr;
        if (slow) {
            // too many distinct k-mers for the shared-memory table: redo the sub-bucket in this CTA's private global table
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < P.cap_slots; i += NT) { t_keys[i] = empty; t_cnt[i] = 0; }
            gkeys = reinterpret_cast<Key*>(P.slow_keys) + (size_t)blockIdx.x * P.slow_slots;
            gcnt = P.slow_cnt + (size_t)blockIdx.x * P.slow_slots;
            if (!slow_ready) {                               // first use by this CTA (the dump below leaves it empty again)
                for (unsigned long long i = threadIdx.x; i < P.slow_slots; i += NT) { gkeys[i] = empty; gcnt[i] = 0; }
                slow_ready = true;
            }
            if (threadIdx.x == 0) { s_nclaim = 0; s_ovf = 0; atomicAdd(&P.counters[0], 1ull); }
            __syncthreads();
            for (unsigned long long c0 = 0; c0 < K; c0 += 65536ull) {
                const bool full = s_nclaim > P.slow_max_fill;                        // uniform: read between two barriers
                unsigned int cl = 0, failed = 0;
                if (!full)
                    for (unsigned long long i = c0 + threadIdx.x; i < min(K, c0 + 65536ull); i += NT) {
                        const int c = gm_insert(gkeys, gcnt, P.slow_slots, kc_load(kp + i), &wrapped);
                        if (c > 0) cl++; else if (c < 0) failed = 1;
                    }
                cl = __reduce_add_sync(0xFFFFFFFFu, cl);
                failed = __reduce_or_sync(0xFFFFFFFFu, failed);
                if (lane == 0 && cl) atomicAdd(&s_nclaim, cl);
                if (lane == 0 && failed) P.flags[0] = 1;
                __syncthreads();
            }
            if (threadIdx.x == 0 && s_nclaim > P.slow_max_fill) P.flags[0] = 1;     // even the global table is too small: the job is redone elsewhere
        }
        const unsigned int D = s_nclaim;
        const bool skip = off + D > P.region_cap;
        if (skip) out_full = true;
        // the claimed slots leave in claim order (dense, every lane busy) and are emptied; the slow path scans its global table
        const uint64_t hbin = mix64((uint64_t)(bin >> P.bin_shift));
        const uint64_t hpre = mix64(hbin);                   // entry_hash's inner term when hi == 0 (64-bit keys)
        auto emit = [&](unsigned long long o, Key kk, uint32_t n) { if (!skip) { okeys[o] = kk; ocnt[o] = n; } };
        if (!slow) {
            for (unsigned int i = threadIdx.x; i < D; i += NT) {
                const uint32_t sl = t_list[i];
                const Key kk = t_keys[sl]; const uint32_t n = t_cnt[sl];
                t_keys[sl] = empty; t_cnt[sl] = 0;
                if (!skip) {
                    okeys[off + i] = kk; ocnt[off + i] = n;
                    uint64_t h;                              // == entry_hash(bin, hi, lo)
                    if constexpr (!WIDE) h = mix64(kk ^ hpre); else h = mix64(kk.lo ^ mix64(kk.hi ^ hbin));
                    dsum += h * (uint64_t)n; dxor ^= mix64(h + n); dcnt += n;
                }
            }
        } else {
            for (unsigned long long s0 = 0; s0 < P.slow_slots; s0 += NT) {
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
        __syncthreads();                                     // the sub-bucket's entries are in the output region (visible to the whole CTA)
        // digest over the dense entries: every lane busy (the lines were just written and sit in L2)
        if (!skip && slow) {
            for (unsigned int i = threadIdx.x; i < D; i += NT) {
                Key kk; const uint32_t n = __ldcg(&ocnt[off + i]);
                uint64_t h;                                  // == entry_hash(bin, hi, lo)
                if constexpr (!WIDE) { kk = __ldcg(&okeys[off + i]); h = mix64(kk ^ hpre); }
                else { const ulonglong2 q = __ldcg(reinterpret_cast<const ulonglong2*>(&okeys[off + i])); kk.lo = q.x; kk.hi = q.y; h = mix64(kk.lo ^ mix64(kk.hi ^ hbin)); }
                dsum += h * (uint64_t)n; dxor ^= mix64(h + n); dcnt += n;
            }
        }
        if (!skip) off += D;
        if (threadIdx.x == 0) { s_nclaim = 0; s_pos = 0; s_ovf = 0; }     // (everybody read them before the barrier above)
        __syncthreads();
    }
    if (threadIdx.x == 0) P.cta_total[blockIdx.x] = off;
    if (out_full && threadIdx.x == 0) P.flags[1] = 1;
    if (wrapped) P.flags[2] = 1;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        dsum += __shfl_xor_sync(0xFFFFFFFFu, dsum, o); dxor ^= __shfl_xor_sync(0xFFFFFFFFu, dxor, o); dcnt += __shfl_xor_sync(0xFFFFFFFFu, dcnt, o);
    }
    if (lane == 0 && dcnt) {
        const int a = (blockIdx.x * (NT / 32) + warp) & 63;
        atomicAdd(&P.acc[a], dsum); atomicXor(&P.acc[64 + a], dxor); atomicAdd(&P.acc[128 + a], dcnt);
    }
}

// ------------------------------------------------------------------ the count kernel of the sort path (useHT = 0)
// The sub-buckets of a bin are KEY RANGES in order (k_expand_hist / k_place_keys in ORDERED mode), and so is the table: a
// k-mer's home slot is cut from the bits that follow its sub-bucket's prefix, which is monotone in the k-mer; linear
// probing never wraps (tail_slots spare slots behind the table; running past them sends the job to the older sort
// kernels).  After the inserts a k-mer sits between its home slot and the end of its cluster (the run of occupied slots
// around it), so k-mers of different clusters are already in order; every cluster is sorted in place by one thread (the
// clusters are a few slots long), and the table, read from slot 0 upwards, is the sub-bucket's (k-mer, count) list in
// ascending order — extractKXmers' sort + heap merge + run-length count (SBKC:540-597, UTIL:642-681) without a sort.
// dynamic shared memory: keys[cap_slots + tail_slots] | cnt[cap_slots + tail_slots]
template <bool WIDE>
__global__ void __launch_bounds__(kKcThreads, 1) k_count_keys_ordered(const KeyCountParams P) {
    typedef typename SmTraits<WIDE>::Key Key;
    extern __shared__ __align__(128) unsigned char kc_raw[];
    const uint32_t n_slots = P.cap_slots + P.tail_slots;
    Key* const t_keys = reinterpret_cast<Key*>(kc_raw);
    uint32_t* const t_cnt = reinterpret_cast<uint32_t*>(t_keys + n_slots);
    __shared__ unsigned int s_ovf;
    __shared__ unsigned int s_wsum[kKcThreads / 32];
    __shared__ unsigned int s_range[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Key empty = sm_empty(Key());
    const int cap_bits = 31 - __clz(P.cap_slots);
    for (uint32_t i = threadIdx.x; i < n_slots; i += kKcThreads) { t_keys[i] = empty; t_cnt[i] = 0; }
    if (threadIdx.x == 0) {
        s_ovf = 0;
        const unsigned long long total = P.mid_key_base[P.n_sub];
        for (int e = 0; e < 2; e++) {
            const unsigned long long want = (total / gridDim.x) * (blockIdx.x + e) + min((unsigned long long)(blockIdx.x + e), total % gridDim.x);
            uint32_t lo = 0, hi = P.n_sub;
            if (blockIdx.x + e >= gridDim.x) lo = P.n_sub;
            else while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (P.mid_key_base[mid] >= want) hi = mid; else lo = mid + 1; }
            s_range[e] = lo;
        }
    }
    __syncthreads();
    const uint32_t m_lo = s_range[0], m_hi = s_range[1];
    Key* const okeys = reinterpret_cast<Key*>(P.out_keys) + (size_t)blockIdx.x * P.region_cap;
    uint32_t* const ocnt = P.out_cnt + (size_t)blockIdx.x * P.region_cap;
    const Key* const keys = reinterpret_cast<const Key*>(P.keys);
    unsigned long long off = 0;
    unsigned long long dsum = 0, dxor = 0, dcnt = 0;
    bool out_full = false;
    const uint32_t spt = (n_slots + kKcThreads - 1) / kKcThreads;               // slots per thread in the ordered passes (<= 32)
    const uint32_t s_lo = min(n_slots, threadIdx.x * spt), s_hi = min(n_slots, s_lo + spt);
    auto home = [&](Key key, int lb) -> uint32_t {          // monotone in the key inside one sub-bucket
        uint32_t top;
        if constexpr (!WIDE) top = (uint32_t)((key << (64 - 2 * P.k)) >> 32);
        else { const int sh = 128 - 2 * P.k; top = (uint32_t)(((sh ? ((key.hi << sh) | (key.lo >> (64 - sh))) : key.hi)) >> 32); }
        return (lb ? top << lb : top) >> (32 - cap_bits);
    };
    auto less = [&](Key a, Key b) -> bool { if constexpr (!WIDE) return a < b; else return key_less(a, b); };

    for (uint32_t m = m_lo; m < m_hi; m++) {
        const uint32_t bin = P.mid_bin[m];
        const uint32_t bl = bin - (uint32_t)P.bin_lo;
        if (threadIdx.x == 0 && P.sub_first[bl] == m) { P.bin_cta[bin] = blockIdx.x; P.bin_off[bin] = off; }
        const unsigned long long kb = P.mid_key_base[m], ke = P.mid_key_base[m + 1];
        if (ke == kb) continue;
        const int lb = 31 - __clz(P.sub_first[bl + 1] - P.sub_first[bl]);          // the bin has 2^lb sub-buckets
        const Key* const kp = keys + kb;
        if (threadIdx.x == 0 && m + 1 < m_hi) {
            const unsigned long long nb0 = (ke * sizeof(Key)) & ~15ull, nb1 = (P.mid_key_base[m + 2] * sizeof(Key) + 15ull) & ~15ull;
            if (nb1 > nb0) bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(keys) + nb0, (uint32_t)min(nb1 - nb0, (unsigned long long)(1u << 20)));
        }
        {
            const uint32_t K32 = (uint32_t)(ke - kb);       // (a bin, hence a sub-bucket, has fewer than 2^32 k-mers)
            uint32_t p = threadIdx.x;
            bool have = p < K32; Key key = Key(); if (have) key = kc_load(kp + p); p += kKcThreads;
            bool hn = p < K32; Key nxt = Key(); if (hn) nxt = kc_load(kp + p); p += kKcThreads;
            bool hn2 = p < K32; Key nxt2 = Key(); if (hn2) nxt2 = kc_load(kp + p); p += kKcThreads;
            uint32_t slot = home(key, lb);
            while (have) {
                bool hit;
                kc_probe(&t_keys[slot], key, hit);
                if (hit) {
                    atomicAdd(&t_cnt[slot], 1u);
                    have = hn; key = nxt; hn = hn2; nxt = nxt2; hn2 = p < K32; if (hn2) nxt2 = kc_load(kp + p); p += kKcThreads;
                    slot = home(key, lb);
                } else if (++slot >= n_slots) { s_ovf = 1; have = false; }
            }
        }
        __syncthreads();                                     // all inserts of the sub-bucket are done
        if (s_ovf) {                                         // ran past the table: the job is redone by the older sort kernels
            __syncthreads();
            if (threadIdx.x == 0) { P.flags[0] = 1; s_ovf = 0; }
            for (uint32_t i = threadIdx.x; i < n_slots; i += kKcThreads) { t_keys[i] = empty; t_cnt[i] = 0; }
            __syncthreads();
            continue;
        }
        // ---- cluster heads in this thread's slots (read-only pass), then one insertion sort per cluster
        uint32_t heads = 0, occ = 0;
        for (uint32_t sl = s_lo; sl < s_hi; sl++) {
            const bool o = !sm_is_empty(t_keys[sl]);
            if (o) { occ |= 1u << (sl - s_lo); if (sl == 0 || sm_is_empty(t_keys[sl - 1])) heads |= 1u << (sl - s_lo); }
        }
        __syncthreads();
        while (heads) {
            const uint32_t h0 = s_lo + (uint32_t)(__ffs((int)heads) - 1);
            heads &= heads - 1u;
            for (uint32_t j = h0 + 1; j < n_slots; j++) {
                const Key kj = t_keys[j];
                if (sm_is_empty(kj)) break;
                const uint32_t cj = t_cnt[j];
                uint32_t i = j;
                while (i > h0 && less(kj, t_keys[i - 1])) { t_keys[i] = t_keys[i - 1]; t_cnt[i] = t_cnt[i - 1]; i--; }
                if (i != j) { t_keys[i] = kj; t_cnt[i] = cj; }
            }
        }
        // ---- the table in slot order is the sorted list: dense output, every thread its own run of slots
        const unsigned int c = (unsigned)__popc(occ);
        unsigned int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();                                     // (also: every cluster is sorted)
        unsigned int wb = 0, D = 0;
#pragma unroll
        for (int i = 0; i < kKcThreads / 32; i++) { const unsigned int t = s_wsum[i]; if (i < warp) wb += t; D += t; }
        const bool skip = off + D > P.region_cap;
        if (skip) out_full = true;
        const uint64_t hbin = mix64((uint64_t)(bin >> P.bin_shift));
        const uint64_t hpre = mix64(hbin);
        unsigned long long o = off + wb + incl - c;
        for (uint32_t sl = s_lo; sl < s_hi; sl++) {
            if (!((occ >> (sl - s_lo)) & 1u)) continue;
            const Key kk = t_keys[sl]; const uint32_t n = t_cnt[sl];
            t_keys[sl] = empty; t_cnt[sl] = 0;
            if (!skip) {
                okeys[o] = kk; ocnt[o] = n; o++;
                uint64_t h;
                if constexpr (!WIDE) h = mix64(kk ^ hpre); else h = mix64(kk.lo ^ mix64(kk.hi ^ hbin));
                dsum += h * (uint64_t)n; dxor ^= mix64(h + n); dcnt += n;
            }
        }
        if (!skip) off += D;
        __syncthreads();
    }
    if (threadIdx.x == 0) P.cta_total[blockIdx.x] = off;
    if (out_full && threadIdx.x == 0) P.flags[1] = 1;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        dsum += __shfl_xor_sync(0xFFFFFFFFu, dsum, o); dxor ^= __shfl_xor_sync(0xFFFFFFFFu, dxor, o); dcnt += __shfl_xor_sync(0xFFFFFFFFu, dcnt, o);
    }
    if (lane == 0 && dcnt) {
        const int a = (blockIdx.x * (kKcThreads / 32) + warp) & 63;
        atomicAdd(&P.acc[a], dsum); atomicXor(&P.acc[64 + a], dxor); atomicAdd(&P.acc[128 + a], dcnt);
    }
}

}  // namespace fkm
