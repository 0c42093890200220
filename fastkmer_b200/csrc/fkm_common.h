// fkm_common.h — small integer helpers shared by host code (g++) and kernels (nvcc).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define FKM_HD __host__ __device__ __forceinline__
#else
#define FKM_HD inline
#endif

namespace fkm {

// UTIL:686-695 hash_to_bucket: Wang/Jenkins 32-bit mix on JVM Ints (wrapping,
// logical shifts), sign bit cleared, modulo the number of bins.
FKM_HD uint32_t hash_to_bucket(uint32_t key, uint32_t B) {
    key = (key ^ 61u) ^ (key >> 16);
    key = key + (key << 3);
    key = key ^ (key >> 4);
    key = key * 0x27d4eb2du;
    key = key ^ (key >> 15);
    return (key & 0x7FFFFFFFu) % B;
}

// Internal bins.  When a bin would hold more k-mers than the partitioned count stage likes (very deep inputs, few bins, a rank
// of a multi-GPU job that owns B/N bins of N times the data), every bin is cut into 2^split internal bins by a second hash of
// the signature: a canonical k-mer has ONE signature, so all its occurrences still meet in one internal bin, and the
// internal bins of a bin are consecutive, so the bin's entries stay contiguous in the (unordered) hash-path output.
FKM_HD uint32_t split_bin(uint32_t sig, uint32_t B, int split) {
    const uint32_t b = hash_to_bucket(sig, B);
    return split ? (b << split) | ((sig * 0x9E3779B1u) >> (32 - split)) : b;
}

// splitmix64 step applied to a counter: the synthetic-data generator of SURVEY
// §8(d) and the result digest use it (neither exists in the reference).
FKM_HD uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// digest term of one result entry (bin, k-mer as hi:lo, count)
FKM_HD uint64_t entry_hash(uint32_t bin, uint64_t hi, uint64_t lo) { return mix64(lo ^ mix64(hi ^ mix64((uint64_t)bin))); }

struct SynthSpec { uint64_t seedG, seedR, seedE, G, R, L, first_read; };

// read r (global index): start position and strand
FKM_HD void synth_read(const SynthSpec& S, uint64_t r, uint64_t& pos, uint64_t& strand) {
    pos = mix64(S.seedR + 2 * r) % (S.G - S.L + 1);
    strand = mix64(S.seedR + 2 * r + 1) & 1ull;
}
// base j of read r; returns 0..3, sets invalid for an 'N'
FKM_HD uint32_t synth_base(const SynthSpec& S, uint64_t r, uint64_t j, uint64_t pos, uint64_t strand, bool& invalid) {
    uint64_t gi = strand ? (pos + (S.L - 1 - j)) : (pos + j);
    uint32_t b = (uint32_t)(mix64(S.seedG + gi) >> 62);
    if (strand) b = 3u - b;
    uint64_t e = mix64(S.seedE + r * S.L + j);
    invalid = (e % 1000ull) == 0ull;
    if (!invalid && (e % 100ull) == 1ull) b = (b + 1u + (uint32_t)((e >> 32) % 3ull)) & 3u;
    return b;
}

// Long-sequence generator (BASELINE config 3): position i of one synthetic genome.
// iid bases; 5 % of the 5-kb blocks are copies of one of 1000 repeat templates; 0.5 % of the
// 1-kb blocks are runs of 'N'.  Counter-based, so any shard [first, first+n) can be generated alone.
struct LongSpec { uint64_t seedG, seedRep, seedN, first_pos; };
FKM_HD uint32_t synth_long_base(const LongSpec& S, uint64_t i, bool& invalid) {
    invalid = (mix64(S.seedN + i / 1000ull) % 200ull) == 0ull;
    const uint64_t blk = i / 5000ull;
    const uint64_t r = mix64(S.seedRep + blk);
    if ((r % 20ull) == 0ull) {
        const uint64_t t = (r >> 20) % 1000ull;
        return (uint32_t)(mix64(S.seedG ^ 0x5bd1e995ull ^ (t * 5000ull + i % 5000ull) * 0x9E3779B97F4A7C15ull) >> 62);
    }
    return (uint32_t)(mix64(S.seedG + i) >> 62);
}

}  // namespace fkm
