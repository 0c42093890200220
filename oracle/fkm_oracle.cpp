// fkm_oracle.cpp — CPU restatement ("oracle") of fastkmer's exact k-mer counting path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under fastkmer_b200/ may link, import or call
// this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and there only as the checker / timed CPU baseline.
//
// PARITY STATUS: "parity unpinned" in the strict sense — the reference
// (/root/reference, Scala on Spark) ships no tests, fixtures or golden vectors and
// cannot be compiled or run in this image (no JVM / Spark / FASTdoop jar).  The
// pins that exist are (a) the known-answer vectors of SURVEY.md App. C (derived
// from a line-by-line transliteration of the Scala), (b) oracle/literal.py, an
// independent literal transliteration kept in this repo, and (c) the clean
// specification in tests/clean_spec.py.  tests/test_oracle_*.py hold all three
// against this file.
//
// Every function cites the reference lines it follows.  Shorthands:
//   UTIL = src/main/scala/skc/package.scala
//   SBKC = src/main/scala/skc/SparkBinKmerCounter.scala
//   TCFG = src/main/scala/skc/test/package.scala
//
// Structure kept from the reference (so that the timed CPU baseline is the
// reference's algorithm, not a different one):
//   map stage     getSuperKmers state machine, O(k) invalid scan per window,
//                 O(k) signature rescan when the minimizer leaves the window
//   shuffle       per-bin concatenation of super-k-mers
//   reduce, HT    per k-mer orientation test + canonical k-mer + open hash map
//   reduce, sort  (k,x)-mer runs, sort of x+1 arrays, k-way heap merge of cursors
//
// Build: see oracle/Makefile (g++ -O2 -shared).  C ABI at the bottom.

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <string>
#include <thread>
#include <vector>

namespace {

typedef unsigned __int128 u128;

// ---------------------------------------------------------------- UTIL:17-41
// A=0 C=1 G=2 T=3, complement = 3 - c.
inline int nt_code(uint8_t c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    return -1;
}
// UTIL:697 notANucleotide — only uppercase ACGT are nucleotides.
inline bool not_a_nucleotide(uint8_t c) { return nt_code(c) < 0; }

// ---------------------------------------------------------------- UTIL:686-695
// JVM Int arithmetic: wrapping 32-bit, >>> is a logical shift.
int32_t hash_to_bucket(int32_t s, int32_t B) {
    uint32_t key = (uint32_t)s;
    const uint32_t c2 = 0x27d4eb2dU;
    key = (key ^ 61U) ^ (key >> 16);
    key = key + (key << 3);
    key = key ^ (key >> 4);
    key = key * c2;
    key = key ^ (key >> 15);
    return (int32_t)((key & 0x7FFFFFFFU) % (uint32_t)B);
}

// ---------------------------------------------------------------- UTIL:46-75
bool is_allowed(int32_t mmer, int length) {
    for (int j = 0; j < length - 3; j++) {
        if ((mmer & 0xf) == 0) return false;   // AA inside
        mmer >>= 2;
    }
    if (mmer == 0) return false;               // AAA prefix
    if (mmer == 0x04) return false;            // ACA prefix
    if ((mmer & 0x3c) == 0) return false;      // AA* prefix
    if ((mmer & 0xf) == 0) return false;       // *AA prefix
    return true;
}

// ---------------------------------------------------------------- UTIL:103-115
int64_t reverse_complement(int64_t seq, int length) {
    int64_t cur = seq, rev = 0;
    int shift = length * 2 - 2;
    for (int i = 0; i < length; i++) {
        rev += (3 - (cur & 3)) << shift;
        cur >>= 2;
        shift -= 2;
    }
    return rev;
}

// ---------------------------------------------------------------- UTIL:77-100
std::vector<int32_t> fill_norm(int sigLen) {
    const int32_t default_signature = 1 << (sigLen * 2);
    std::vector<int32_t> norm((size_t)default_signature);
    for (int32_t i = 0; i < default_signature; i++) {
        int32_t rev = (int32_t)reverse_complement(i, sigLen);
        int32_t str_val = is_allowed(i, sigLen) ? i : default_signature;
        int32_t rev_val = is_allowed(rev, sigLen) ? rev : default_signature;
        norm[(size_t)i] = std::min(str_val, rev_val);
    }
    return norm;
}

// ---------------------------------------------------------------- UTIL:739-754
// first/last offset (relative to start) of a non-ACGT byte in s[start,end), or (-1,-1).
inline void first_last_invalid(const uint8_t* s, int64_t start, int64_t end, int64_t& first, int64_t& last) {
    first = -1; last = -1;
    for (int64_t i = start; i < end; i++) {
        if (not_a_nucleotide(s[i])) {
            if (first == -1) first = i - start;
            last = i - start;
        }
    }
}

// ---------------------------------------------------------------- UTIL:337-357
// Kmer.getSignature on the window cur[i, i+k): minimum of norm over its k-m+1
// m-mers, leftmost position on ties (strict '<' update).
inline void get_signature(const uint8_t* w, int k, int m, const int32_t* norm, int32_t& sig, int& pos) {
    const int32_t mask = (int32_t)((1u << (m * 2)) - 1);
    int32_t data = 0;
    for (int i = 0; i < m; i++) data = ((data << 2) + nt_code(w[i])) & mask;   // Mmer.insert UTIL:552-558
    sig = norm[data]; pos = 0;
    for (int i = m; i < k; i++) {
        data = ((data << 2) + nt_code(w[i])) & mask;
        int32_t v = norm[data];
        if (v < sig) { sig = v; pos = i - m + 1; }
    }
}

// ---------------------------------------------------------------- UTIL:310-326
inline int32_t last_m(const uint8_t* w, int k, int m, const int32_t* norm) {
    int32_t res = 0;
    for (int i = k - m; i < k; i++) res = (res << 2) | nt_code(w[i]);
    return norm[res];
}

// A super-k-mer as shipped through the "shuffle": length + 2-bit symbols.
// (The reference ships Kmer objects: Int length + Array[Long] with 31 nt per
// Long, UTIL:138-172; the word layout is not observable, so the oracle keeps one
// symbol per byte after unpacking.)
struct BinStore {
    std::vector<uint8_t> bytes;      // records: u32 length, then ceil(len/4) packed bytes
    uint64_t n_kmers = 0;            // SBKC:347,363,382,407 binSizes
    void push(const uint8_t* s, int64_t len, int k) {
        uint32_t L = (uint32_t)len;
        size_t at = bytes.size();
        bytes.resize(at + 4 + (L + 3) / 4);
        memcpy(&bytes[at], &L, 4);
        uint8_t* p = &bytes[at + 4];
        for (uint32_t i = 0; i < L; i += 4) {
            uint8_t b = 0;
            for (uint32_t j = 0; j < 4; j++) { b <<= 2; if (i + j < L) b |= (uint8_t)nt_code(s[i + j]); }
            *p++ = b;
        }
        n_kmers += (uint64_t)(len - k + 1);
    }
};

struct MapStats { uint64_t n_bases = 0, n_kmers = 0, n_superkmers = 0, superkmer_bases = 0, n_records = 0; };

// ---------------------------------------------------------------- SBKC:67-157
// getSuperKmers / getSuperKmersWithBinSizes on one record `cur` of `len` bytes.
void scan_record(const uint8_t* cur, int64_t len, int k, int m, int B, const int32_t* norm,
                 std::vector<BinStore>& out, MapStats& st) {
    st.n_bases += (uint64_t)len; st.n_records++;
    if (len < k) return;                                             // SBKC:67
    int32_t min_value = -1; int64_t min_pos = -1;                    // Signature(-1,-1) SBKC:69
    int64_t super_kmer_start = 0, i = 0;
    auto flush = [&](int64_t start, int64_t length) {
        out[(size_t)hash_to_bucket(min_value, B)].push(cur + start, length, k);
        st.n_superkmers++; st.superkmer_bases += (uint64_t)length; st.n_kmers += (uint64_t)(length - k + 1);
    };
    while (i < len - k + 1) {                                        // SBKC:75
        int64_t nf, nl;
        first_last_invalid(cur, i, i + k, nf, nl);                   // SBKC:78
        if (nf != -1) {                                              // SBKC:81-97
            if (super_kmer_start < i) flush(super_kmer_start, i - 1 + k - super_kmer_start);
            super_kmer_start = i + nl + 1;
            i += nl + 1;
        } else {
            if (i > min_pos) {                                       // SBKC:102-114
                if (super_kmer_start < i) { flush(super_kmer_start, i - 1 + k - super_kmer_start); super_kmer_start = i; }
                int32_t sig; int pos;
                get_signature(cur + i, k, m, norm, sig, pos);
                min_value = sig; min_pos = pos + i;
            } else {                                                 // SBKC:115-136
                int32_t last = last_m(cur + i, k, m, norm);
                if (last < min_value) {
                    if (super_kmer_start < i) { flush(super_kmer_start, i - 1 + k - super_kmer_start); super_kmer_start = i; }
                    min_value = last; min_pos = i + k - m;
                }
            }
            i += 1;
        }
    }
    if (len - super_kmer_start >= k) {                               // SBKC:142-157
        int64_t nf, nl;
        first_last_invalid(cur, i, len, nf, nl);
        if (nf == -1) flush(super_kmer_start, len - super_kmer_start);
        else if (i + nf >= super_kmer_start + k) flush(super_kmer_start, i + nf);   // SBKC:152-156 (unreachable)
    }
}

// ---------------------------------------------------------------- UTIL:721-728
// getOrientation(Kmer,i,j): 0 if the k-mer is < its reverse complement, else 1
// (palindromes -> 1).  s = one symbol per byte.
inline int get_orientation(const uint8_t* s, int i, int j) {
    for (;;) {
        int start = s[i], endc = 3 - s[j];
        if (start < endc) return 0;
        if (start > endc || i >= j) return 1;
        i++; j--;
    }
}

// symbols [a,b] of s as a right-aligned 2-bit integer; reverse-complemented if rc.
// Stands in for Kmer.readFromKmer UTIL:174-295 (word layout unobservable).
inline u128 read_from(const uint8_t* s, int a, int b, int rc) {
    u128 v = 0;
    if (!rc) for (int i = a; i <= b; i++) v = (v << 2) | s[i];
    else     for (int i = b; i >= a; i--) v = (v << 2) | (uint8_t)(3 - s[i]);
    return v;
}

struct Entry { int32_t bin; uint64_t hi, lo; uint32_t cnt; };

inline uint64_t mix64(uint64_t x) {                // splitmix64 finaliser (not from the reference)
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// ---------------------------------------------------------------- SBKC:664-739
// extractKXmersHT on one bin.  Open-addressing map pre-sized by the k-mer upper
// bound, like Object2IntOpenHashMap(expected) (fastutil 7.2.0, load factor .75).
void reduce_bin_ht(int32_t bin, const BinStore& bs, int k, std::vector<Entry>& out) {
    if (bs.n_kmers == 0) return;
    size_t cap = 16; while (cap * 3 / 4 < bs.n_kmers + 1) cap <<= 1;
    std::vector<u128> keys(cap); std::vector<uint32_t> cnt(cap, 0);
    std::vector<uint8_t> sk;
    const uint8_t* p = bs.bytes.data(); const uint8_t* e = p + bs.bytes.size();
    while (p < e) {
        uint32_t L; memcpy(&L, p, 4); p += 4;
        sk.resize(L);
        for (uint32_t i = 0; i < L; i++) sk[i] = (p[i >> 2] >> (6 - 2 * (i & 3))) & 3;
        p += (L + 3) / 4;
        for (int i = 0; i + k <= (int)L; i++) {                              // SBKC:698
            int o = get_orientation(sk.data(), i, i + k - 1);                 // SBKC:700
            u128 km = read_from(sk.data(), i, i + k - 1, o);                 // SBKC:701
            size_t h = (size_t)(mix64((uint64_t)km ^ mix64((uint64_t)(km >> 64)))) & (cap - 1);
            while (cnt[h] != 0 && keys[h] != km) h = (h + 1) & (cap - 1);
            keys[h] = km; cnt[h]++;                                          // SBKC:703 addTo(kmer,1)
        }
    }
    for (size_t h = 0; h < cap; h++)
        if (cnt[h]) out.push_back(Entry{bin, (uint64_t)(keys[h] >> 64), (uint64_t)keys[h], cnt[h]});
}

// ---------------------------------------------------------------- UTIL:562-601
struct RIndex {            // cursor over sorted array r, sub-range [pos,end), k-mer at `shift`
    const std::vector<u128>* arr; size_t pos, end; int shift, r; u128 cur;
};
// ---------------------------------------------------------------- SBKC:428-660
// extractKXmers on one bin: (k,x)-mer runs, x+1 sorts, k-way merge through a heap.
void reduce_bin_sort(int32_t bin, const BinStore& bs, int k, int x, std::vector<Entry>& out) {
    if (bs.n_kmers == 0) return;
    std::vector<std::vector<u128>> R((size_t)x + 1);
    std::vector<uint8_t> sk;
    const uint8_t* p = bs.bytes.data(); const uint8_t* e = p + bs.bytes.size();
    while (p < e) {
        uint32_t L; memcpy(&L, p, 4); p += 4;
        sk.resize(L);
        for (uint32_t i = 0; i < L; i++) sk[i] = (p[i >> 2] >> (6 - 2 * (i & 3))) & 3;
        p += (L + 3) / 4;
        int lastOrientation = -1, orientation = -1, runLength = 0, runStart = 0;   // SBKC:474-476
        for (int i = 0; i + k <= (int)L; i++) {                                       // SBKC:484
            orientation = get_orientation(sk.data(), i, i + k - 1);
            if (orientation == lastOrientation) {
                runLength++;
                if (runLength == x + 1) {                                             // SBKC:495-502
                    R[(size_t)runLength - 1].push_back(read_from(sk.data(), runStart, runStart + k + runLength - 2, orientation));
                    runLength = 0; runStart = i; lastOrientation = -1;
                }
            } else {
                if (lastOrientation != -1)                                            // SBKC:507-511
                    R[(size_t)runLength - 1].push_back(read_from(sk.data(), runStart, runStart + k + runLength - 2, lastOrientation));
                runLength = 1; runStart = i; lastOrientation = orientation;
            }
        }
        if (runLength > 0)                                                            // SBKC:520-524
            R[(size_t)runLength - 1].push_back(read_from(sk.data(), runStart, runStart + k + runLength - 2, lastOrientation));
    }
    for (auto& a : R) std::sort(a.begin(), a.end());                                  // SBKC:540-542
    // priorityQueueWithIndexes UTIL:642-681
    const u128 kmask = (k == 64) ? ~(u128)0 : (((u128)1 << (2 * k)) - 1);
    auto kmer_at = [&](const RIndex& c) -> u128 { return ((*c.arr)[c.pos] >> (2 * (c.r - c.shift))) & kmask; };
    auto cmp = [](const RIndex& a, const RIndex& b) { return a.cur > b.cur; };        // PointedMinOrder UTIL:604-614
    std::priority_queue<RIndex, std::vector<RIndex>, decltype(cmp)> heap(cmp);
    auto push_idx = [&](const std::vector<u128>& a, size_t s, size_t en, int shift, int r) {
        RIndex c{&a, s, en, shift, r, 0}; c.cur = kmer_at(c); heap.push(c);
    };
    for (int r = 0; r <= x; r++) {
        const std::vector<u128>& a = R[(size_t)r];
        if (a.empty()) continue;
        std::vector<size_t> starts((size_t)r, 0);
        push_idx(a, 0, a.size(), 0, r);                                               // UTIL:658
        auto firstM = [&](size_t j, int mm) -> u128 { return a[j] >> (2 * (k + r - mm)); };   // UTIL:329-334
        for (size_t j = 1; j < a.size(); j++)
            for (int i = 0; i < r; i++)
                if (firstM(j - 1, i + 1) != firstM(j, i + 1)) { push_idx(a, starts[(size_t)i], j, i + 1, r); starts[(size_t)i] = j; }
        for (int i = 0; i < r; i++) push_idx(a, starts[(size_t)i], a.size(), i + 1, r);
    }
    bool have = false; u128 last = 0; uint32_t last_cnt = 0;                          // SBKC:560-597
    while (!heap.empty()) {
        RIndex c = heap.top(); heap.pop();
        if (have && c.cur == last) last_cnt++;
        else {
            if (have) out.push_back(Entry{bin, (uint64_t)(last >> 64), (uint64_t)last, last_cnt});
            last = c.cur; last_cnt = 1; have = true;
        }
        c.pos++;
        if (c.pos < c.end) { c.cur = kmer_at(c); heap.push(c); }
    }
    if (have) out.push_back(Entry{bin, (uint64_t)(last >> 64), (uint64_t)last, last_cnt});
}

// ---------------------------------------------------------------- records
// SURVEY App. A.1 / SBKC:62-65: a record's value is its sequence lines with '\n'
// removed and nothing else stripped.  FASTdoop (not vendored) does the record
// split; the oracle defines: a record starts at a '>' that begins a line, its
// header runs to the end of that line, bytes before the first header are ignored.
struct Rec { std::vector<uint8_t> seq; };
void parse_fasta(const uint8_t* t, size_t n, std::vector<std::pair<size_t, size_t>>& recs) {
    // recs: [begin,end) byte range of the sequence lines of each record
    size_t i = 0;
    bool bol = true;
    while (i < n && !(bol && t[i] == '>')) { bol = (t[i] == '\n'); i++; }
    while (i < n) {
        while (i < n && t[i] != '\n') i++;       // header line
        if (i < n) i++;
        size_t b = i; bol = true;
        while (i < n && !(bol && t[i] == '>')) { bol = (t[i] == '\n'); i++; }
        recs.emplace_back(b, i);
    }
}

struct Result {
    std::vector<Entry> entries;      // sorted by (bin, key) when sorted==1
    MapStats st;
    uint64_t n_distinct = 0, digest_sum = 0, digest_xor = 0, total_count = 0;
    double ms_map = 0, ms_reduce = 0;
};

inline double now_ms() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

Result* run(const uint8_t* fasta, size_t n, int k, int m, int x, int max_b, int use_ht, int threads, int sorted) {
    Result* res = new Result();
    const int B = (int)std::min<int64_t>((int64_t)1 << (2 * m), (int64_t)max_b);     // TCFG:32
    std::vector<std::pair<size_t, size_t>> recs;
    parse_fasta(fasta, n, recs);
    if (threads < 1) threads = 1;
    double t0 = now_ms();
    std::vector<int32_t> norm = fill_norm(m);                                         // SBKC:47
    std::vector<std::vector<BinStore>> per_thread((size_t)threads, std::vector<BinStore>((size_t)B));
    std::vector<MapStats> stats((size_t)threads);
    {
        std::atomic<size_t> next(0);
        size_t chunk = 256;
        // map tasks: whole records, except that a long record is cut into pieces of kPiece window starts
        // (k-1 bytes of look-ahead), the way FASTdoop's FASTAlongInputFormat hands one PartialSequence per
        // input split to getSuperKmers (SBKC:62-63,1012); every k-window is still seen exactly once.
        const size_t kPiece = (size_t)4 << 20;
        struct Task { size_t rec, first, last; };          // byte range [first,last) of the record's text
        std::vector<Task> tasks;
        for (size_t r = 0; r < recs.size(); r++) {
            const size_t len = recs[r].second - recs[r].first;
            if (len <= 2 * kPiece) tasks.push_back(Task{r, 0, len});
            else for (size_t a = 0; a < len; a += kPiece) tasks.push_back(Task{r, a, std::min(len, a + kPiece)});
        }
        chunk = std::max<size_t>(1, std::min<size_t>(256, tasks.size() / ((size_t)threads * 8 + 1)));
        auto work = [&](int t) {
            std::vector<uint8_t> cur;
            for (;;) {
                size_t b = next.fetch_add(chunk);
                if (b >= tasks.size()) break;
                size_t e = std::min(tasks.size(), b + chunk);
                for (size_t ti = b; ti < e; ti++) {
                    const Task& T = tasks[ti];
                    const uint8_t* txt = fasta + recs[T.rec].first;
                    const size_t rec_len = recs[T.rec].second - recs[T.rec].first;
                    cur.clear();
                    for (size_t i = T.first; i < T.last; i++) if (txt[i] != '\n') cur.push_back(txt[i]);   // SBKC:63-64 replaceAll("\n","")
                    if (T.last < rec_len) {                 // look-ahead: the next k-1 sequence bytes
                        size_t need = (size_t)k - 1;
                        for (size_t i = T.last; i < rec_len && need; i++) if (txt[i] != '\n') { cur.push_back(txt[i]); need--; }
                    }
                    const uint64_t nb0 = stats[(size_t)t].n_bases;
                    scan_record(cur.data(), (int64_t)cur.size(), k, m, B, norm.data(), per_thread[(size_t)t], stats[(size_t)t]);
                    if (T.last < rec_len) {                 // do not count the look-ahead bytes twice
                        size_t la = 0, need = (size_t)k - 1;
                        for (size_t i = T.last; i < rec_len && need; i++) if (txt[i] != '\n') { la++; need--; }
                        stats[(size_t)t].n_bases = nb0 + (uint64_t)(cur.size() - la);
                    }
                    if (T.first != 0) stats[(size_t)t].n_records--;
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(work, t);
        work(0);
        for (auto& q : th) q.join();
    }
    for (auto& s : stats) {
        res->st.n_bases += s.n_bases; res->st.n_kmers += s.n_kmers; res->st.n_superkmers += s.n_superkmers;
        res->st.superkmer_bases += s.superkmer_bases; res->st.n_records += s.n_records;
    }
    double t1 = now_ms();
    res->ms_map = t1 - t0;
    // shuffle (SBKC:1035,1042 reduceByKey(_ ++ _)) + reduce
    std::vector<std::vector<Entry>> outs((size_t)threads);
    {
        std::atomic<int> next(0);
        auto work = [&](int t) {
            BinStore merged;
            for (;;) {
                int b = next.fetch_add(1);
                if (b >= B) break;
                merged.bytes.clear(); merged.n_kmers = 0;
                for (int s = 0; s < threads; s++) {
                    BinStore& src = per_thread[(size_t)s][(size_t)b];
                    merged.bytes.insert(merged.bytes.end(), src.bytes.begin(), src.bytes.end());
                    merged.n_kmers += src.n_kmers;
                    std::vector<uint8_t>().swap(src.bytes);
                }
                if (use_ht) reduce_bin_ht(b, merged, k, outs[(size_t)t]);
                else        reduce_bin_sort(b, merged, k, x, outs[(size_t)t]);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(work, t);
        work(0);
        for (auto& q : th) q.join();
    }
    res->ms_reduce = now_ms() - t1;
    size_t total = 0; for (auto& o : outs) total += o.size();
    res->entries.reserve(total);
    for (auto& o : outs) { res->entries.insert(res->entries.end(), o.begin(), o.end()); std::vector<Entry>().swap(o); }
    for (const Entry& en : res->entries) {
        uint64_t h = mix64(en.lo ^ mix64(en.hi ^ mix64((uint64_t)(uint32_t)en.bin)));
        res->digest_sum += h * (uint64_t)en.cnt;
        res->digest_xor ^= mix64(h + en.cnt);
        res->total_count += en.cnt;
    }
    res->n_distinct = res->entries.size();
    if (sorted)
        std::sort(res->entries.begin(), res->entries.end(), [](const Entry& a, const Entry& b) {
            if (a.bin != b.bin) return a.bin < b.bin;
            if (a.hi != b.hi) return a.hi < b.hi;
            return a.lo < b.lo;
        });
    return res;
}

}  // namespace

// ======================================================================== C ABI
extern "C" {

int32_t fkmo_hash_to_bucket(int32_t s, int32_t B) { return hash_to_bucket(s, B); }
int32_t fkmo_is_allowed(int32_t mmer, int32_t length) { return is_allowed(mmer, length) ? 1 : 0; }
int64_t fkmo_reverse_complement(int64_t seq, int32_t length) { return reverse_complement(seq, length); }
// out must hold 4^m int32
void fkmo_fill_norm(int32_t m, int32_t* out) {
    std::vector<int32_t> n = fill_norm(m);
    memcpy(out, n.data(), n.size() * sizeof(int32_t));
}
// signature (min norm, leftmost pos) of the k-window w (ASCII ACGT)
void fkmo_signature(const uint8_t* w, int32_t k, int32_t m, int32_t* sig, int32_t* pos) {
    std::vector<int32_t> n = fill_norm(m);
    int p; int32_t s; get_signature(w, k, m, n.data(), s, p); *sig = s; *pos = p;
}
int32_t fkmo_orientation(const uint8_t* ascii, int32_t k) {
    std::vector<uint8_t> s((size_t)k);
    for (int i = 0; i < k; i++) s[(size_t)i] = (uint8_t)nt_code(ascii[i]);
    return get_orientation(s.data(), 0, k - 1);
}

// super-k-mers of one record under the reference's cutting rule, for the worked
// example of SURVEY App. C.3: writes up to cap tuples (bin, sig_unused, start, len).
int32_t fkmo_superkmers(const uint8_t* rec, int64_t len, int32_t k, int32_t m, int32_t max_b,
                        int32_t* bins, int32_t* lens, int32_t cap) {
    const int B = (int)std::min<int64_t>((int64_t)1 << (2 * m), (int64_t)max_b);
    std::vector<int32_t> norm = fill_norm(m);
    std::vector<BinStore> out((size_t)B);
    MapStats st;
    scan_record(rec, len, k, m, B, norm.data(), out, st);
    int n = 0;
    for (int b = 0; b < B; b++) {
        const uint8_t* p = out[(size_t)b].bytes.data(); const uint8_t* e = p + out[(size_t)b].bytes.size();
        while (p < e) { uint32_t L; memcpy(&L, p, 4); p += 4 + (L + 3) / 4; if (n < cap) { bins[n] = b; lens[n] = (int32_t)L; } n++; }
    }
    return n;
}

// bin of every k-window of one record (ASCII), -1 where the window holds a non-ACGT
// byte: hash_to_bucket(getSignature(window)) — SBKC:99-112 applied to each window.
void fkmo_window_bins(const uint8_t* rec, int64_t len, int32_t k, int32_t m, int32_t max_b, int32_t* out) {
    const int B = (int)std::min<int64_t>((int64_t)1 << (2 * m), (int64_t)max_b);
    std::vector<int32_t> norm = fill_norm(m);
    for (int64_t i = 0; i + k <= len; i++) {
        int64_t nf, nl; first_last_invalid(rec, i, i + k, nf, nl);
        if (nf != -1) { out[i] = -1; continue; }
        int32_t sig; int pos; get_signature(rec + i, k, m, norm.data(), sig, pos);
        out[i] = hash_to_bucket(sig, B);
    }
}

void* fkmo_count(const uint8_t* fasta, uint64_t n, int32_t k, int32_t m, int32_t x, int32_t max_b,
                 int32_t use_ht, int32_t threads, int32_t sorted) {
    if (k < m || m < 3 || m > 15 || k > 64 || max_b < 1) return nullptr;
    if (!use_ht && (x < 1 || k + x > 64)) return nullptr;            // x=0 crashes the sort path
    return run(fasta, (size_t)n, k, m, x, max_b, use_ht, threads, sorted);
}
uint64_t fkmo_result_size(void* r) { return ((Result*)r)->entries.size(); }
// copies entries into caller arrays (each of fkmo_result_size elements)
void fkmo_result_copy(void* r, int32_t* bin, uint64_t* hi, uint64_t* lo, uint32_t* cnt) {
    Result* R = (Result*)r;
    for (size_t i = 0; i < R->entries.size(); i++) {
        bin[i] = R->entries[i].bin; hi[i] = R->entries[i].hi; lo[i] = R->entries[i].lo; cnt[i] = R->entries[i].cnt;
    }
}
// stats[0..11]: n_bases n_kmers n_superkmers superkmer_bases n_records n_distinct total_count digest_sum digest_xor ms_map ms_reduce(both as integer microseconds)
void fkmo_result_stats(void* r, uint64_t* s) {
    Result* R = (Result*)r;
    s[0] = R->st.n_bases; s[1] = R->st.n_kmers; s[2] = R->st.n_superkmers; s[3] = R->st.superkmer_bases;
    s[4] = R->st.n_records; s[5] = R->n_distinct; s[6] = R->total_count; s[7] = R->digest_sum; s[8] = R->digest_xor;
    s[9] = (uint64_t)(R->ms_map * 1000.0); s[10] = (uint64_t)(R->ms_reduce * 1000.0);
}
void fkmo_result_free(void* r) { delete (Result*)r; }

// ---- the synthetic inputs of SURVEY §8(d) (counter-based splitmix64), restated here so that the CPU baseline of bench.py
// builds its input without loading the product library.  Must produce the text of fkm_synth_fasta_host /
// fkm_synth_long_fasta_host byte for byte (tests/test_oracle_kat.py::test_oracle_generators_match_the_library).
static inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
// reads first_read .. first_read + R - 1 as '>r<global index>\n<L bases>\n'; returns the bytes (out == NULL sizes)
uint64_t fkmo_synth_fasta(uint64_t seedG, uint64_t seedR, uint64_t seedE, uint64_t G, uint64_t R, uint64_t L, uint64_t first_read,
                          uint8_t* out, uint64_t cap, int32_t threads) {
    // bytes of the reads with global index in [g0, g1): '>r' + decimal index + '\n' + L bases + '\n'
    auto bytes_of = [&](uint64_t g0, uint64_t g1) {
        uint64_t total = 0, lo_d = 0, hi_d = 10;
        for (uint64_t d = 1; d <= 20; d++) {
            const uint64_t a = std::max(g0, lo_d), e = (d == 20) ? g1 : std::min(g1, hi_d);
            if (e > a) total += (e - a) * (d + L + 4);
            lo_d = hi_d; hi_d = (d < 19) ? hi_d * 10 : ~0ull;
        }
        return total;
    };
    const unsigned nt = (unsigned)std::max<int32_t>(1, std::min<int32_t>(threads, 256));
    std::vector<uint64_t> lo(nt + 1), start(nt + 1, 0);
    for (unsigned t = 0; t <= nt; t++) lo[t] = R * t / nt;
    for (unsigned t = 0; t < nt; t++) start[t + 1] = start[t] + bytes_of(first_read + lo[t], first_read + lo[t + 1]);
    if (!out) return start[nt];
    if (cap < start[nt]) return 0;
    auto work = [&](unsigned t) {
        uint8_t* p = out + start[t];
        char num[32];
        for (uint64_t r = lo[t]; r < lo[t + 1]; r++) {
            const uint64_t g = first_read + r;
            const int n = snprintf(num, sizeof num, ">r%llu\n", (unsigned long long)g);
            memcpy(p, num, (size_t)n); p += n;
            const uint64_t pos = sm64(seedR + 2 * g) % (G - L + 1), strand = sm64(seedR + 2 * g + 1) & 1ull;
            for (uint64_t j = 0; j < L; j++) {
                const uint64_t gi = strand ? pos + (L - 1 - j) : pos + j;
                uint32_t b = (uint32_t)(sm64(seedG + gi) >> 62);
                if (strand) b = 3u - b;
                const uint64_t e = sm64(seedE + g * L + j);
                if (e % 1000ull == 0ull) { *p++ = 'N'; continue; }
                if (e % 100ull == 1ull) b = (b + 1u + (uint32_t)((e >> 32) % 3ull)) & 3u;
                *p++ = (uint8_t)"ACGT"[b];
            }
            *p++ = '\n';
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& q : th) q.join();
    return start[nt];
}
// one long record (BASELINE config 3): '>chr1 synthetic\n' + 70-column lines of bases [first_pos, first_pos + n)
uint64_t fkmo_synth_long_fasta(uint64_t seedG, uint64_t seedRep, uint64_t seedN, uint64_t first_pos, uint64_t n, uint8_t* out, uint64_t cap, int32_t threads) {
    const char* hdr = ">chr1 synthetic\n";
    const uint64_t hl = strlen(hdr), W = 70, lines = (n + W - 1) / W, total = hl + n + lines;
    if (!out) return total;
    if (cap < total) return 0;
    memcpy(out, hdr, hl);
    const unsigned nt = (unsigned)std::max<int32_t>(1, std::min<int32_t>(threads, 256));
    auto work = [&](unsigned t) {
        for (uint64_t ln = lines * t / nt; ln < lines * (t + 1) / nt; ln++) {
            uint8_t* p = out + hl + ln * (W + 1);
            const uint64_t a = ln * W, e = std::min(n, a + W);
            for (uint64_t q = a; q < e; q++) {
                const uint64_t i = first_pos + q;
                if (sm64(seedN + i / 1000ull) % 200ull == 0ull) { *p++ = 'N'; continue; }
                const uint64_t r = sm64(seedRep + i / 5000ull);
                uint32_t b;
                if (r % 20ull == 0ull) { const uint64_t t5 = (r >> 20) % 1000ull; b = (uint32_t)(sm64(seedG ^ 0x5bd1e995ull ^ (t5 * 5000ull + i % 5000ull) * 0x9E3779B97F4A7C15ull) >> 62); }
                else b = (uint32_t)(sm64(seedG + i) >> 62);
                *p++ = (uint8_t)"ACGT"[b];
            }
            *p = '\n';
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& q : th) q.join();
    return total;
}

// SURVEY App. C.4 sequential-LCG read set as FASTA text ('>r<i>\n<seq>\n').
// Returns bytes written (call with out==NULL to size).
uint64_t fkmo_gen_lcg_fasta(uint64_t seed, uint64_t G, uint64_t R, uint64_t L, uint8_t* out, uint64_t cap) {
    uint64_t x = seed;
    auto nxt = [&]() -> uint64_t { x = x * 6364136223846793005ULL + 1442695040888963407ULL; return x >> 33; };
    static const char* ACGT = "ACGT";
    std::string genome(G, 'A');
    for (uint64_t i = 0; i < G; i++) genome[i] = ACGT[nxt() & 3];
    std::string o; o.reserve(R * (L + 12));
    std::string s(L, 'A');
    for (uint64_t r = 0; r < R; r++) {
        uint64_t pos = nxt() % (G - L + 1);
        uint64_t strand = nxt() & 1;
        for (uint64_t j = 0; j < L; j++) s[j] = genome[pos + j];
        if (strand) {
            std::reverse(s.begin(), s.end());
            for (auto& c : s) c = (c == 'A') ? 'T' : (c == 'C') ? 'G' : (c == 'G') ? 'C' : 'A';
        }
        for (uint64_t j = 0; j < L; j++) {
            uint64_t e = nxt() % 100;
            if (e == 0) s[j] = 'N';
            else if (e == 1) s[j] = ACGT[nxt() & 3];
        }
        o += ">r"; o += std::to_string(r); o += "\n"; o += s; o += "\n";
    }
    if (out && cap >= o.size()) memcpy(out, o.data(), o.size());
    return o.size();
}

}  // extern "C"
