// Runs the kernels' own integer helpers (fastkmer_b200/csrc/fkm_math.h, fkm_common.h) on the CPU and compares them
// with string-level definitions.  Built and run by tests/test_device_math.py (g++, no CUDA).  Test infrastructure.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#include <random>
#include "fkm_common.h"
#include "fkm_math.h"

using namespace fkm;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } g_fail++; } } while (0)

static const char* ACGT = "ACGT";
static int code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }
static std::string rc_str(const std::string& s) {
    std::string r(s.rbegin(), s.rend());
    for (char& c : r) c = ACGT[3 - code(c)];
    return r;
}
static std::string rnd(std::mt19937_64& g, int n) { std::string s(n, 'A'); for (char& c : s) c = ACGT[g() & 3]; return s; }
// right-aligned value of a string of <= 64 bases as (hi, lo)
static key128 pack_right(const std::string& s) {
    key128 v{0, 0};
    for (char c : s) { v.hi = (v.hi << 2) | (v.lo >> 62); v.lo = (v.lo << 2) | (uint64_t)code(c); }
    return v;
}
// left-aligned 128-bit image of a string of <= 64 bases: w0 = first 32 bases, w1 = next 32
static void pack_left(const std::string& s, uint64_t& w0, uint64_t& w1) {
    w0 = w1 = 0;
    for (size_t i = 0; i < s.size(); i++) { const uint64_t c = (uint64_t)code(s[i]); if (i < 32) w0 |= c << (62 - 2 * i); else w1 |= c << (62 - 2 * (i - 32)); }
}
static bool allowed_str(const std::string& s) {                    // UTIL:46-75 as SURVEY App. A.5 states it
    return s.find("AA") == std::string::npos && s.compare(0, 3, "ACA") != 0;
}

int main() {
    std::mt19937_64 g(12345);
    // reverse complements
    for (int it = 0; it < 20000; it++) {
        const int len = 1 + (int)(g() % 32);
        const std::string s = rnd(g, len);
        CHECK(revcomp64(pack_right(s).lo, len) == pack_right(rc_str(s)).lo, "revcomp64 len %d", len);
        const int len2 = 33 + (int)(g() % 32);
        const std::string t = rnd(g, len2);
        const key128 r = revcomp128(pack_right(t), len2), want = pack_right(rc_str(t));
        CHECK(r.hi == want.hi && r.lo == want.lo, "revcomp128 len %d", len2);
        const int m = 3 + (int)(g() % 13);
        const std::string u = rnd(g, m);
        CHECK(revcomp32((uint32_t)pack_right(u).lo, m) == (uint32_t)pack_right(rc_str(u)).lo, "revcomp32 m %d", m);
    }
    // closed-form norm == min(allowed(v) ? v : 4^m, allowed(rc v) ? rc v : 4^m): exhaustive for m <= 9, sampled above
    for (int m = 3; m <= 15; m++) {
        const uint32_t mmask = (uint32_t)((1ull << (2 * m)) - 1);
        const uint64_t total = 1ull << (2 * m);
        const uint64_t step = m <= 9 ? 1 : total / 200000 + 1;
        for (uint64_t v = 0; v < total; v += step) {
            std::string s(m, 'A');
            for (int i = 0; i < m; i++) s[i] = ACGT[(v >> (2 * (m - 1 - i))) & 3];
            const std::string r = rc_str(s);
            const uint32_t rv = (uint32_t)pack_right(r).lo;
            const uint32_t dflt = mmask + 1u;
            const uint32_t a = allowed_str(s) ? (uint32_t)v : dflt, b = allowed_str(r) ? rv : dflt;
            CHECK(mmer_allowed((uint32_t)v, m, mmask) == allowed_str(s), "allowed m %d v %llu", m, (unsigned long long)v);
            CHECK(mmer_norm((uint32_t)v, rv, m, mmask) == std::min(a, b), "norm m %d v %llu", m, (unsigned long long)v);
        }
    }
    // canonical records: tail zeroed, min(string, reverse complement), left-aligned
    for (int it = 0; it < 50000; it++) {
        const int len = 1 + (int)(g() % 60);
        const std::string s = rnd(g, len), tail = rnd(g, 64 - len);
        uint64_t w0, w1; pack_left(s + tail, w0, w1); w1 &= ~0xFFull;
        canon_record_narrow(w0, w1, len);
        uint64_t e0, e1; pack_left(std::min(s, rc_str(s)), e0, e1);
        CHECK(w0 == e0 && w1 == e1 && (w1 & 0xFF) == 0, "canon len %d", len);
    }
    // k-mer j of a record == canonical value of the substring
    for (int it = 0; it < 3000; it++) {
        const int k = 3 + (int)(g() % 30);                        // 3..32, NARROW record: 60 bases
        const std::string s = rnd(g, 60);
        uint64_t rec[2]; pack_left(s + "AAAA", rec[0], rec[1]); rec[1] &= ~0xFFull;
        for (int j = 0; j + k <= 60; j++) {
            const std::string w = s.substr(j, k);
            CHECK(kmer_at_narrow(rec, j, k) == std::min(pack_right(w).lo, pack_right(rc_str(w)).lo), "kmer_at_narrow k %d j %d", k, j);
        }
        const int kw = 33 + (int)(g() % 32);                      // 33..64, WIDE record: 124 bases
        const std::string t = rnd(g, 124);
        uint64_t recw[4]; pack_left(t.substr(0, 64), recw[0], recw[1]); pack_left(t.substr(64) + "AAAA", recw[2], recw[3]); recw[3] &= ~0xFFull;
        for (int j = 0; j + kw <= 124; j++) {
            const std::string w = t.substr(j, kw);
            const key128 a = pack_right(w), b = pack_right(rc_str(w));
            const key128 want = key_less(b, a) ? b : a, got = kmer_at_wide(recw, j, kw);
            CHECK(got.hi == want.hi && got.lo == want.lo, "kmer_at_wide k %d j %d", kw, j);
        }
    }
    // table slots: mulhi(hash, size) stays inside the table; the 32-bit form is floor(h * size / 2^32)
    for (int it = 0; it < 200000; it++) {
        const uint32_t h = (uint32_t)g();
        const unsigned long long small = 1 + g() % 0xFFFFFFFFull, big = (1ull << 32) + g() % (1ull << 36);
        CHECK(slot_of(h, small) == (((unsigned long long)h * small) >> 32) && slot_of(h, small) < small, "slot_of small");
        CHECK(slot_of(h, big) < big, "slot_of big");
    }
    // SURVEY App. C.1 known answers of hash_to_bucket (UTIL:686-695)
    CHECK(hash_to_bucket(0, 2048) == 362 && hash_to_bucket(12345, 2048) == 1043 && hash_to_bucket(1048576, 2048) == 1821 &&
          hash_to_bucket(4194303, 4096) == 382, "hash_to_bucket KAT");
    // internal bins: split_bin keeps the configuration's bin in the high bits (a bin's internal bins are consecutive) and adds a
    // second hash of the signature below it; split 0 is hash_to_bucket itself
    {
        std::vector<unsigned> seen(16, 0);
        for (int it = 0; it < 200000; it++) {
            const uint32_t sig = (uint32_t)(g() & 0x3FFFFFFFu), B = 1 + (uint32_t)(g() % 5000);
            const int sp = (int)(g() % 7);
            const uint32_t ib = split_bin(sig, B, sp);
            CHECK((ib >> sp) == hash_to_bucket(sig, B) && ib < (B << sp), "split_bin B %u split %d", B, sp);
            CHECK(split_bin(sig, B, 0) == hash_to_bucket(sig, B), "split_bin split 0");
            if (sp == 4) seen[ib & 15u]++;
        }
        unsigned lo = ~0u, hi = 0;
        for (unsigned c : seen) { lo = std::min(lo, c); hi = std::max(hi, c); }
        CHECK(hi < 2 * lo, "split_bin spreads signatures over the internal bins (%u..%u)", lo, hi);
    }
    // the partitioned count stage: sub-bucket (high bits of part_hash) and table slot (folded low bits) are independent enough that the
    // k-mers of ONE sub-bucket of 111 fill a 16384-slot table evenly (no 64-slot block more than 4x as full as the average)
    {
        std::vector<unsigned> blocks(256, 0);
        unsigned n_in = 0;
        for (int it = 0; it < 3000000; it++) {
            const uint64_t key = g() >> 8;                                    // a 56-bit canonical 28-mer stand-in
            const uint32_t h = part_hash(key);
            if (umulhi32(h, 111u) != 5u) continue;
            blocks[part_slot(h, 16383u) >> 6]++; n_in++;
        }
        unsigned hi = 0;
        for (unsigned c : blocks) hi = std::max(hi, c);
        CHECK(n_in > 20000 && hi * 256u < 4u * n_in * 2u, "part_slot spread inside a sub-bucket (max block %u of %u keys)", hi, n_in);
        key128 w; w.lo = 0x0123456789ABCDEFull; w.hi = 0x0FEDCBA987654321ull;
        key128 w2 = w; w2.hi ^= 1ull << 40;
        CHECK(part_hash(w) != part_hash(w2), "part_hash(key128) depends on the high word");
    }
    if (g_fail) { printf("%d checks failed\n", g_fail); return 1; }
    printf("ok\n");
    return 0;
}
