// fkm_math.h — the pure integer arithmetic of the kernels: reverse complements, the closed-form norm of an m-mer,
// canonical records, table hashing, cutting canonical k-mers out of a super-k-mer record.  Device code under nvcc;
// plain inline functions under a host compiler, so that tests/test_device_math.py can run the very functions the
// kernels use on the CPU (tests/host_math_check.cpp) against string-level definitions.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FKM_DEV __device__ __forceinline__
#else
#define FKM_DEV static inline
#endif

namespace fkm {

struct alignas(16) key128 { uint64_t lo, hi; };

// bit reversal and high halves of products: hardware instructions on the device
FKM_DEV uint64_t brev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
    x = ((x & 0x3333333333333333ull) << 2) | ((x >> 2) & 0x3333333333333333ull);
    x = ((x & 0x0F0F0F0F0F0F0F0Full) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0Full);
    x = ((x & 0x00FF00FF00FF00FFull) << 8) | ((x >> 8) & 0x00FF00FF00FF00FFull);
    x = ((x & 0x0000FFFF0000FFFFull) << 16) | ((x >> 16) & 0x0000FFFF0000FFFFull);
    return (x << 32) | (x >> 32);
#endif
}
FKM_DEV uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    return (uint32_t)(brev64((uint64_t)x) >> 32);
#endif
}
FKM_DEV uint32_t umulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
FKM_DEV unsigned long long umulhi64(unsigned long long a, unsigned long long b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (unsigned long long)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

FKM_DEV uint64_t swap_pairs(uint64_t x) {
    return ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
}
// reverse complement of a right-aligned len-mer (len <= 32)
FKM_DEV uint64_t revcomp64(uint64_t x, int len) {
#if defined(__CUDA_ARCH__)
    // two BREVs, then per word ONE LOP3 that swaps the two bits of every pair and complements:
    // ~(((v >> 1) & 0x55555555) | ((v << 1) & 0xAAAAAAAA))   (LUT 0x1B over a = v >> 1, b = v << 1, c = 0x55555555)
    const uint32_t a = __brev((uint32_t)x), b = __brev((uint32_t)(x >> 32));
    uint32_t ra, rb;
    asm("lop3.b32 %0, %1, %2, 0x55555555, 0x1B;" : "=r"(ra) : "r"(a >> 1), "r"(a << 1));
    asm("lop3.b32 %0, %1, %2, 0x55555555, 0x1B;" : "=r"(rb) : "r"(b >> 1), "r"(b << 1));
    return (((uint64_t)ra << 32) | rb) >> (64 - 2 * len);
#else
    return swap_pairs(brev64(~x)) >> (64 - 2 * len);
#endif
}
FKM_DEV key128 revcomp128(key128 x, int len) {          // 32 < len <= 64
    uint64_t rh = swap_pairs(brev64(~x.lo)), rl = swap_pairs(brev64(~x.hi));
    int s = 128 - 2 * len;                                                  // 0..62
    key128 r;
    r.lo = s ? ((rl >> s) | (rh << (64 - s))) : rl;
    r.hi = rh >> s;
    return r;
}
FKM_DEV bool key_less(key128 a, key128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
FKM_DEV bool key_eq(key128 a, key128 b) { return a.hi == b.hi && a.lo == b.lo; }
FKM_DEV bool key_eq(uint64_t a, uint64_t b) { return a == b; }

// norm of an m-mer given the value v and its reverse complement r (UTIL:46-100,
// closed form of SURVEY App. A.5: allowed <=> no "AA" inside and prefix != "ACA").
FKM_DEV bool mmer_allowed(uint32_t v, int m, uint32_t mmask) {
    uint32_t nz = (v | (v >> 1)) & 0x55555555u;
    uint32_t a = ~nz & 0x55555555u & mmask;
    return ((a & (a >> 2)) == 0u) && ((v >> (2 * m - 6)) != 4u);
}
FKM_DEV uint32_t mmer_norm(uint32_t v, uint32_t r, int m, uint32_t mmask) {
    uint32_t dflt = mmask + 1u;
    uint32_t a = mmer_allowed(v, m, mmask) ? v : dflt;
    uint32_t b = mmer_allowed(r, m, mmask) ? r : dflt;
    return a < b ? a : b;
}


FKM_DEV uint32_t revcomp32(uint32_t v, int len) {       // len <= 15
    uint32_t x = brev32(~v);
    x = ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
    return x >> (32 - 2 * len);
}


// Record folding (hash path, NARROW records): a record and its reverse complement hold the same canonical
// k-mers, and so do two records that differ only behind their n+k-1 bases.  Zero the unused tail and keep the
// smaller of the string and its reverse complement (left-aligned, so the order is lexicographic): identical
// super-k-mers of different reads, either strand, become bit-identical records.  w1's n byte must be clear.
FKM_DEV void canon_record_narrow(uint64_t& w0, uint64_t& w1, int len) {
    const int bits = 2 * len;                                               // 2..120
    w0 &= bits >= 64 ? ~0ull : ~0ull << (64 - bits);
    w1 &= bits <= 64 ? 0ull : ~0ull << (128 - bits);
    // complement, reverse all 64 base positions: the string's reverse complement lands in the low `bits` bits
    const uint64_t c_hi = swap_pairs(brev64(~w1)), c_lo = swap_pairs(brev64(~w0));
    const int sft = 128 - bits;                                             // 8..126: shift it back to the top
    uint64_t r0, r1;
    if (sft >= 64) { r0 = c_lo << (sft - 64); r1 = 0ull; }
    else { r0 = (c_hi << sft) | (c_lo >> (64 - sft)); r1 = c_lo << sft; }
    if (r0 < w0 || (r0 == w0 && r1 < w1)) { w0 = r0; w1 = r1; }
}


// 32-bit table hash (murmur3 finaliser over the folded key); the slot is mulhi(hash, size)
FKM_DEV uint32_t fmix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
FKM_DEV uint32_t key_hash(uint64_t key) {
    return fmix32((uint32_t)key * 0x9E3779B1u ^ (uint32_t)(key >> 32) * 0x85EBCA77u);
}
FKM_DEV uint32_t key_hash(key128 key) {
    return fmix32(((uint32_t)key.lo * 0x9E3779B1u ^ (uint32_t)(key.lo >> 32) * 0x85EBCA77u) +
                  ((uint32_t)key.hi * 0xC2B2AE3Du ^ (uint32_t)(key.hi >> 32) * 0x27D4EB2Fu));
}
// Hash of the partitioned count path (fkm_part.cuh): the sub-bucket of a k-mer is mulhi(h, sub-buckets) (the high bits of h),
// its table slot part_slot(h) (low bits, folded once more), so that the k-mers of one sub-bucket spread over the whole table.
FKM_DEV uint32_t part_hash(uint64_t key) {
    uint32_t x = (uint32_t)key * 0x9E3779B1u ^ (uint32_t)(key >> 32) * 0x85EBCA77u;
    x ^= x >> 16; x *= 0xC2B2AE3Du;
    return x;
}
FKM_DEV uint32_t part_hash(key128 key) {
    uint32_t x = ((uint32_t)key.lo * 0x9E3779B1u ^ (uint32_t)(key.lo >> 32) * 0x85EBCA77u) + ((uint32_t)key.hi * 0x27D4EB2Fu ^ (uint32_t)(key.hi >> 32) * 0x165667B1u);
    x ^= x >> 16; x *= 0xC2B2AE3Du;
    return x;
}
FKM_DEV uint32_t part_slot(uint32_t h, uint32_t mask) { return (h ^ (h >> 13)) & mask; }
FKM_DEV unsigned long long slot_of(uint32_t h, unsigned long long size) {
    return (size <= 0xFFFFFFFFull) ? (unsigned long long)umulhi32(h, (uint32_t)size)
                                   : umulhi64(((unsigned long long)h << 32) | fmix32(h), size);
}


// canonical k-mer number j of a record whose words sit in shared memory (n byte already cleared)
FKM_DEV uint64_t kmer_at_narrow(const uint64_t* rec, int j, int k) {
    const int q = j >> 5, sh = 2 * (j & 31);
    const uint64_t a0 = rec[q], a1 = q ? 0ull : rec[1];
    const uint64_t hi = sh ? ((a0 << sh) | (a1 >> (64 - sh))) : a0;
    const uint64_t fwd = hi >> (64 - 2 * k);
    const uint64_t rc = revcomp64(fwd, k);
    return fwd < rc ? fwd : rc;
}
FKM_DEV key128 kmer_at_wide(const uint64_t* rec, int j, int k) {
    const int q = j >> 5, sh = 2 * (j & 31);
    const uint64_t a0 = rec[q], a1 = (q + 1 < 4) ? rec[q + 1] : 0ull, a2 = (q + 2 < 4) ? rec[q + 2] : 0ull;
    const uint64_t h0 = sh ? ((a0 << sh) | (a1 >> (64 - sh))) : a0;
    const uint64_t h1 = sh ? ((a1 << sh) | (a2 >> (64 - sh))) : a1;
    const int s = 128 - 2 * k;
    key128 fwd;
    fwd.hi = h0 >> s; fwd.lo = s ? ((h0 << (64 - s)) | (h1 >> s)) : h1;
    const key128 rc = revcomp128(fwd, k);
    return key_less(rc, fwd) ? rc : fwd;
}


}  // namespace fkm
