// fkm_host.cpp — host-side helpers: FASTA record split + 2-bit packing (the host twin of the
// device ingest; replaces the FASTdoop record readers, SBKC:62-65,1009-1012), directory
// creation for the bin files, and the synthetic FASTA generators of SURVEY §8(d).
#include "fkm_host.h"
#include "fkm_common.h"
#include "../../include/fastkmer_b200.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

// Record split: a record starts at a '>' that begins a line; its
// header runs to the end of that line; the value is every following byte up to
// the next header with '\n' removed (SBKC:63-64 replaceAll("\n","")) and nothing
// else stripped.  Bytes before the first header are ignored.  Each record is laid
// out as its bytes followed by ONE invalid separator position.
static int pack_fasta_serial(const uint8_t* t, uint64_t n, uint64_t* bases, uint32_t* invalid,
                             uint64_t cap_positions, uint64_t* n_positions, uint64_t* n_bases) {
    static int8_t code[256]; static bool init = false;
    if (!init) { for (int i = 0; i < 256; i++) code[i] = -1; code['A'] = 0; code['C'] = 1; code['G'] = 2; code['T'] = 3; init = true; }
    uint64_t p = 0, nb = 0;
    uint64_t bw = 0; uint32_t iw = 0;
    const bool write = bases != nullptr;
    auto put = [&](int c) {            // c in 0..3, or -1 for an invalid position
        if (write) {
            bw = (bw << 2) | (uint64_t)(c < 0 ? 0 : c); iw = (iw << 1) | (c < 0 ? 1u : 0u);
            if ((p & 31) == 31) { bases[p >> 5] = bw; invalid[p >> 5] = iw; bw = 0; iw = 0; }
        }
        p++;
    };
    uint64_t i = 0; bool bol = true;
    while (i < n && !(bol && t[i] == '>')) { bol = (t[i] == '\n'); i++; }
    while (i < n) {
        while (i < n && t[i] != '\n') i++;          // header line
        if (i < n) i++;
        bol = true;
        while (i < n && !(bol && t[i] == '>')) {
            uint8_t c = t[i];
            bol = (c == '\n');
            if (c != '\n') {
                if (write && p >= cap_positions) return fkm_set_error(FKM_EINVAL, "packed buffer too small");
                put(code[c]); nb++;
            }
            i++;
        }
        if (write && p >= cap_positions) return fkm_set_error(FKM_EINVAL, "packed buffer too small");
        put(-1);                                     // record separator
    }
    if (write && (p & 31)) {                         // flush the partial word, tail positions invalid
        unsigned rem = 32 - (unsigned)(p & 31);
        bases[p >> 5] = bw << (2 * rem); invalid[p >> 5] = (iw << rem) | ((1u << rem) - 1u);
    }
    if (n_positions) *n_positions = p;
    if (n_bases) *n_bases = nb;
    return FKM_OK;
}

// The same layout from `threads` host threads: the text is cut at record starts, a first pass counts every range's positions
// (so that every range knows its first position), a second pass packs the ranges in place; the words two ranges share are
// merged afterwards.  threads <= 0: one per hardware thread (small inputs stay on the calling thread).
extern "C" int fkm_pack_fasta_mt(const uint8_t* t, uint64_t n, uint64_t* bases, uint32_t* invalid,
                                 uint64_t cap_positions, uint64_t* n_positions, uint64_t* n_bases, int32_t threads) {
    unsigned T = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
    T = (unsigned)std::min<uint64_t>(T, std::max<uint64_t>(1, n >> 16));              // at least 64 KB of text per thread
    if (T <= 1) return pack_fasta_serial(t, n, bases, invalid, cap_positions, n_positions, n_bases);
    static int8_t code[256]; static bool init = false;
    if (!init) { for (int i = 0; i < 256; i++) code[i] = -1; code['A'] = 0; code['C'] = 1; code['G'] = 2; code['T'] = 3; init = true; }
    auto next_header = [&](uint64_t from) -> uint64_t {                                // first '>' at the start of a line at or after `from`
        while (from < n) {
            const uint8_t* q = (const uint8_t*)memchr(t + from, '>', (size_t)(n - from));
            if (!q) return n;
            const uint64_t at = (uint64_t)(q - t);
            if (at == 0 || t[at - 1] == '\n') return at;
            from = at + 1;
        }
        return n;
    };
    std::vector<uint64_t> cut(T + 1, n);
    cut[0] = next_header(0);                                                           // bytes before the first header are ignored
    for (unsigned i = 1; i < T; i++) cut[i] = next_header(std::max(cut[i - 1], n / T * i));
    // walks the records of [lo, hi) line by line: line(p, len) for every line of a record's value (without its '\n'), sep() after every record
    auto walk = [&](uint64_t lo, uint64_t hi, auto&& line, auto&& sep) {
        uint64_t i = lo;
        while (i < hi) {
            const uint8_t* e = (const uint8_t*)memchr(t + i, '\n', (size_t)(hi - i));  // header line
            i = e ? (uint64_t)(e - t) + 1 : hi;
            while (i < hi && t[i] != '>') {                                            // (a '>' starts a record only at the start of a line)
                e = (const uint8_t*)memchr(t + i, '\n', (size_t)(hi - i));
                const uint64_t le = e ? (uint64_t)(e - t) : hi;
                if (le > i) line(t + i, le - i);
                i = e ? le + 1 : hi;
            }
            sep();
        }
    };
    std::vector<uint64_t> cnt(T, 0), nbs(T, 0), first(T + 1, 0);
    {
        std::vector<std::thread> th;
        for (unsigned i = 0; i < T; i++)
            th.emplace_back([&, i]() {
                uint64_t c = 0, nb = 0;
                walk(cut[i], cut[i + 1], [&](const uint8_t*, uint64_t len) { c += len; nb += len; }, [&]() { c++; });
                cnt[i] = c; nbs[i] = nb;
            });
        for (auto& x : th) x.join();
    }
    uint64_t nb_total = 0;
    for (unsigned i = 0; i < T; i++) { first[i + 1] = first[i] + cnt[i]; nb_total += nbs[i]; }
    const uint64_t P = first[T];
    if (n_positions) *n_positions = P;
    if (n_bases) *n_bases = nb_total;
    if (!bases) return FKM_OK;
    if (P > cap_positions) return fkm_set_error(FKM_EINVAL, "packed buffer too small");
    struct Part { uint64_t word; uint64_t b; uint32_t v; };                            // bits of a word shared with a neighbouring range
    std::vector<std::vector<Part>> parts(T);
    {
        std::vector<std::thread> th;
        for (unsigned i = 0; i < T; i++)
            th.emplace_back([&, i]() {
                const uint64_t p0 = first[i], p1 = first[i + 1];
                // bw / iw hold the positions of the current word seen so far, right-aligned; `fill` of them (the range's first word
                // starts at position p0 & 31)
                uint64_t p = p0, bw = 0; uint32_t iw = 0;
                auto flush = [&](uint64_t word, uint64_t b, uint32_t v) {               // word `word` as far as this range goes, bits in place
                    const bool mine = (word << 5) >= p0 && ((word + 1) << 5) <= p1;    // all 32 positions belong to this range
                    if (mine) { bases[word] = b; invalid[word] = v; } else parts[i].push_back(Part{word, b, v});
                };
                auto put = [&](int c) {
                    bw = (bw << 2) | (uint64_t)(c < 0 ? 0 : c); iw = (iw << 1) | (c < 0 ? 1u : 0u);
                    if ((p & 31) == 31) { flush(p >> 5, bw, iw); bw = 0; iw = 0; }
                    p++;
                };
                walk(cut[i], cut[i + 1],
                     [&](const uint8_t* q, uint64_t len) { for (uint64_t j = 0; j < len; j++) put((int)code[q[j]]); },
                     [&]() { put(-1); });
                if (p & 31) {                                                          // the range ends inside a word: shift its bits into place
                    const unsigned rem = 32u - (unsigned)(p & 31);
                    flush(p >> 5, bw << (2 * rem), iw << rem);
                }
            });
        for (auto& x : th) x.join();
    }
    // shared words: zero them, then OR every range's share in; the unused tail of the last word is invalid
    for (unsigned i = 0; i < T; i++) for (const Part& q : parts[i]) { bases[q.word] = 0; invalid[q.word] = 0; }
    for (unsigned i = 0; i < T; i++) for (const Part& q : parts[i]) { bases[q.word] |= q.b; invalid[q.word] |= q.v; }
    if (P & 31) invalid[P >> 5] |= (1u << (32 - (unsigned)(P & 31))) - 1u;
    return FKM_OK;
}

extern "C" int fkm_pack_fasta(const uint8_t* t, uint64_t n, uint64_t* bases, uint32_t* invalid,
                              uint64_t cap_positions, uint64_t* n_positions, uint64_t* n_bases) {
    return fkm_pack_fasta_mt(t, n, bases, invalid, cap_positions, n_positions, n_bases, 0);
}

static int mkdir_p(const std::string& dir) {
    std::string cur;
    for (size_t i = 0; i <= dir.size(); i++) {
        if (i == dir.size() || dir[i] == '/') {
            if (!cur.empty() && cur != "/") {
                if (mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return -1;
            }
        }
        if (i < dir.size()) cur += dir[i];
    }
    return 0;
}

int fkm_make_dirs(const char* dir) {
    if (mkdir_p(dir) != 0) return fkm_set_error(FKM_EIO, "cannot create %s: %s", dir, strerror(errno));
    return FKM_OK;
}

// SURVEY §8(d) synthetic reads as FASTA text: '>r<global index>\n<seq>\n'
extern "C" int fkm_synth_fasta_host(const fkm_synth* s, uint8_t* out, uint64_t cap, uint64_t* n_bytes) {
    if (!s || s->genome_len < s->read_len || s->read_len == 0) return fkm_set_error(FKM_EINVAL, "bad synthetic spec");
    fkm::SynthSpec S{s->seed_genome, s->seed_reads, s->seed_errors, s->genome_len, s->n_reads, s->read_len, s->first_read};
    auto hdr_len = [](uint64_t r) { uint64_t d = 1; while (r >= 10) { r /= 10; d++; } return 3 + d; };   // '>' 'r' digits '\n'
    // sizes: prefix of header lengths is needed for parallel fill
    const uint64_t R = S.R, L = S.L;
    unsigned nt = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    std::vector<uint64_t> start(nt + 1, 0);
    std::vector<uint64_t> first(nt + 1, 0);
    for (unsigned t = 0; t <= nt; t++) first[t] = R * t / nt;
    for (unsigned t = 0; t < nt; t++) {
        uint64_t bytes = 0;
        for (uint64_t r = first[t]; r < first[t + 1]; ) {       // runs of equal digit count
            uint64_t g = S.first_read + r, d = hdr_len(g);
            uint64_t lim = 10; while (lim <= g) lim *= 10;      // first index with one more digit
            uint64_t e = std::min<uint64_t>(first[t + 1], r + (lim - g));
            bytes += (e - r) * (d + L + 1);
            r = e;
        }
        start[t + 1] = start[t] + bytes;
    }
    if (n_bytes) *n_bytes = start[nt];
    if (!out) return FKM_OK;
    if (cap < start[nt]) return fkm_set_error(FKM_EINVAL, "FASTA buffer too small");
    auto work = [&](unsigned t) {
        uint8_t* p = out + start[t];
        char num[32];
        for (uint64_t r = first[t]; r < first[t + 1]; r++) {
            uint64_t g = S.first_read + r, pos, strand;
            int n = snprintf(num, sizeof num, ">r%llu\n", (unsigned long long)g);
            memcpy(p, num, (size_t)n); p += n;
            fkm::synth_read(S, g, pos, strand);
            for (uint64_t j = 0; j < L; j++) {
                bool bad; uint32_t b = fkm::synth_base(S, g, j, pos, strand, bad);
                *p++ = bad ? 'N' : (uint8_t)"ACGT"[b];
            }
            *p++ = '\n';
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& q : th) q.join();
    return FKM_OK;
}

// BASELINE config 3: one long record, 70-column lines
extern "C" int fkm_synth_long_fasta_host(const fkm_synth_long* s, uint8_t* out, uint64_t cap, uint64_t* n_bytes) {
    if (!s) return fkm_set_error(FKM_EINVAL, "bad synthetic spec");
    const uint64_t n = s->n_bases, W = 70;
    const char* hdr = ">chr1 synthetic\n";
    const uint64_t hl = strlen(hdr);
    const uint64_t total = hl + n + (n + W - 1) / W;
    if (n_bytes) *n_bytes = total;
    if (!out) return FKM_OK;
    if (cap < total) return fkm_set_error(FKM_EINVAL, "FASTA buffer too small");
    memcpy(out, hdr, hl);
    fkm::LongSpec S{s->seed_genome, s->seed_repeats, s->seed_n, s->first_pos};
    unsigned nt = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    const uint64_t lines = (n + W - 1) / W;
    auto work = [&](unsigned t) {
        for (uint64_t ln = lines * t / nt; ln < lines * (t + 1) / nt; ln++) {
            uint8_t* p = out + hl + ln * (W + 1);
            const uint64_t a = ln * W, e = std::min(n, a + W);
            for (uint64_t i = a; i < e; i++) { bool bad; uint32_t b = fkm::synth_long_base(S, S.first_pos + i, bad); *p++ = bad ? 'N' : (uint8_t)"ACGT"[b]; }
            *p = '\n';
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& q : th) q.join();
    return FKM_OK;
}
