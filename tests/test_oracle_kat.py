"""Pins the CPU oracle (oracle/fkm_oracle.cpp) against the known-answer vectors of
SURVEY.md App. C and against the clean spec (tests/clean_spec.py).  CPU only."""
import random

import pytest

import clean_spec
import oracle_lib

GOLD = {  # id: (seed G R L, k m x B, N_k, distinct, bins, maxcount, superkmers, superkmer bases, sha256)  App. C.4
    "G1": ((42, 2000, 200, 100), (28, 10, 3, 2048), 10628, 3706, 274, 14, 1334, 46646,
           "a53c138c5a9fe1638f6fa005e383821af030c1935d095f8813adb9dcc077d90f"),
    "G2": ((43, 3000, 120, 150), (31, 11, 3, 4096), 10376, 4837, 337, 7, 1130, 44276,
           "d69568e6a4cc2777b5af4cc66497737ad978b0ec53b9eded510a8a6a9b8d1940"),
    "G3": ((44, 3000, 120, 150), (55, 13, 3, 2048), 6071, 4261, 160, 6, 418, 28643,
           "f227fe83008813dd1fba1ebb0e837aa9c81de7918a0e014b0ed6a0a42d707920"),
    "G4": ((45, 300, 60, 40), (5, 3, 1, 64), 2061, 252, 22, 36, 1113, 6513,
           "1496d3bae854c9bb75bd0f8bfebe714d205f763b0d8a4bfccb6ebbeda62ad793"),
}


def test_hash_to_bucket_kat(oracle):                      # App. C.1 (UTIL:686-695)
    kat2048 = {0: 362, 1: 1181, 2: 1397, 61: 0, 12345: 1043, 1048575: 2011, 1048576: 1821, 4194304: 414, 67108864: 452}
    kat4096 = {0: 2410, 1: 3229, 123456: 2125, 4194303: 382, 4194304: 414}
    for s, b in kat2048.items():
        assert oracle.hash_to_bucket(s, 2048) == b == clean_spec.hash_to_bucket(s, 2048)
    for s, b in kat4096.items():
        assert oracle.hash_to_bucket(s, 4096) == b == clean_spec.hash_to_bucket(s, 4096)


def test_fill_norm_kat(oracle):                           # App. C.2 (UTIL:77-100)
    want3 = [int(t) for t in ("63 47 31 15 59 5 6 7 8 9 10 7 12 13 14 15 62 17 18 14 20 21 22 10 24 25 22 6 28 29 18 31 "
                              "61 33 29 13 36 37 25 9 40 37 21 5 44 33 17 47 60 44 28 12 52 40 24 8 52 36 20 59 60 61 62 63").split()]
    assert oracle.fill_norm(3).tolist() == want3
    n4 = oracle.fill_norm(4)
    assert len(set(n4.tolist())) == 135 and int((n4 == 256).sum()) == 2
    n10 = oracle.fill_norm(10)
    assert (n10[0], n10[1], n10[5], n10[727835]) == (1048575, 786431, 720895, 111025)
    n11 = oracle.fill_norm(11)
    assert (n11[0], n11[1], n11[1776411]) == (4194303, 3145727, 444102)
    n13 = oracle.fill_norm(13)
    assert (n13[0], n13[1], n13[1776411]) == (67108863, 50331647, 7105647)


@pytest.mark.parametrize("m", [3, 4, 5, 6, 7])
def test_norm_closed_form(oracle, m):                     # App. A.5: is_allowed == no "AA" and not "ACA" prefix
    n = oracle.fill_norm(m)
    for v in range(4 ** m):
        s = "".join("ACGT"[(v >> (2 * (m - 1 - i))) & 3] for i in range(m))
        assert n[v] == clean_spec.norm(s)


def test_worked_example(oracle):                          # App. C.3
    read = b"ACGTTGCANGGCTTAACCGGTA"
    sk = oracle.superkmers(read, 5, 3, 64)
    assert sorted(l for _, l in sk) == sorted([5, 5, 6, 6, 5, 5, 7, 6])
    res = oracle.count(b">r0\n" + read + b"\n", 5, 3, 1, 64, 1)
    got = {}
    for b, h, l, c in zip(res["bin"], res["hi"], res["lo"], res["cnt"]):
        got.setdefault(int(b), []).append((oracle_lib.kmer_str(h, l, 5), int(c)))
    want = {12: [("GCAAC", 1), ("TGCAA", 1)], 13: [("AACCG", 1), ("ACCGG", 2), ("CGGTA", 1), ("GGTTA", 1)],
            18: [("AACGT", 1), ("CAACG", 1)], 55: [("CTTAA", 1)], 58: [("AAGCC", 1), ("GCTTA", 1)], 60: [("GTTAA", 1)]}
    assert got == want
    assert res["stats"]["n_kmers"] == 13


@pytest.mark.parametrize("gid", sorted(GOLD))
@pytest.mark.parametrize("use_ht", [0, 1])
def test_golden_digests(oracle, gid, use_ht):             # App. C.4
    (seed, G, R, L), (k, m, x, B), nk, dist, nbins, maxc, nsk, skb, sha = GOLD[gid]
    fasta = oracle.gen_lcg_fasta(seed, G, R, L)
    if gid == "G1":
        assert fasta.split(b"\n")[1][:30] == b"TCCTACACGACGGCTCTCGACCAAATCGGC"
    res = oracle.count(fasta, k, m, x, B, use_ht)
    st = res["stats"]
    assert (st["n_kmers"], st["n_distinct"], st["n_superkmers"], st["superkmer_bases"]) == (nk, dist, nsk, skb)
    assert len(set(res["bin"].tolist())) == nbins and int(res["cnt"].max()) == maxc
    assert st["total_count"] == nk
    assert oracle_lib.digest_of(res, k) == sha


def _rand_fasta(rng, n_reads, lo, hi, alphabet="ACGT", p_bad=0.02, bad="NnacgtRY\r ", width=None):
    out = []
    for r in range(n_reads):
        L = rng.randint(lo, hi)
        s = "".join(rng.choice(bad) if rng.random() < p_bad else rng.choice(alphabet) for _ in range(L))
        if width:
            s = "\n".join(s[i:i + width] for i in range(0, len(s), width))
        out.append(">r%d some header\n%s\n" % (r, s))
    return "".join(out)


@pytest.mark.parametrize("k,m,x,B", [(5, 3, 1, 64), (12, 4, 2, 100), (20, 5, 3, 2000), (28, 10, 3, 2048), (31, 11, 3, 4096),
                                      (32, 7, 2, 333), (33, 8, 3, 512), (55, 13, 3, 2048), (60, 9, 4, 77), (61, 6, 3, 4096)])
def test_oracle_vs_clean_spec(oracle, k, m, x, B):       # HT == sort == clean spec (SURVEY §4 plan (2))
    rng = random.Random(k * 1000 + m)
    for alphabet, width in (("ACGT", None), ("AC", 17), ("ACGT", 70), ("A", None)):
        fasta = _rand_fasta(rng, 25, 0, 3 * k + 40, alphabet=alphabet, width=width)
        want = clean_spec.count(fasta, k, m, B)
        for use_ht in (0, 1):
            res = oracle.count(fasta.encode(), k, m, x, B, use_ht, threads=3)
            got = {(int(b), oracle_lib.kmer_str(h, l, k)): int(c)
                   for b, h, l, c in zip(res["bin"], res["hi"], res["lo"], res["cnt"])}
            assert got == dict(want)
            assert res["stats"]["total_count"] == sum(want.values()) == res["stats"]["n_kmers"]


def test_strand_symmetry_and_bins(oracle):                # sig(w) == sig(rc w): every k-mer re-hashes to its bin
    rng = random.Random(7)
    seq = "".join(rng.choice("ACGT") for _ in range(3000))
    a = oracle.count((">a\n%s\n" % seq).encode(), 28, 10, 3, 2048, 1)
    b = oracle.count((">a\n%s\n" % clean_spec.revcomp(seq)).encode(), 28, 10, 3, 2048, 0)
    assert oracle_lib.digest_of(a, 28) == oracle_lib.digest_of(b, 28)
    for bn, h, l in zip(a["bin"][:200], a["hi"][:200], a["lo"][:200]):
        s = oracle_lib.kmer_str(h, l, 28).encode()
        sig, _ = oracle.signature(s, 10)
        assert oracle.hash_to_bucket(sig, 2048) == int(bn)


def test_rejects_x0_on_sort_path(oracle):                 # App. A.8(3)
    with pytest.raises(ValueError):
        oracle.count(b">a\nACGTACGTACGT\n", 5, 3, 0, 64, 0)


def test_oracle_generators_match_the_library(oracle):
    """bench.py's CPU baseline builds its input with the oracle's own restatement of the SURVEY §8(d) generators (so that it never loads
    the product library); the text must equal the product's host generator byte for byte (which the GPU tests hold equal to the
    device generator)."""
    import numpy as np
    import fastkmer_b200 as fk
    for spec in (dict(seeds=(2001, 2002, 2003), genome_len=100000, n_reads=20011, read_len=150),
                 dict(seeds=(7, 8, 9), genome_len=3000, n_reads=1234, read_len=100, first_read=99_999_000),
                 dict(seeds=(1, 2, 3), genome_len=500, n_reads=5, read_len=40, first_read=7)):
        for threads in (1, 8):
            assert np.array_equal(oracle.synth_fasta(spec, threads=threads), fk.synth_fasta(spec))
    for spec in (dict(seeds=(3001, 3002, 3003), n_bases=300_001), dict(seeds=(5, 6, 7), n_bases=12_345, first_pos=4_999_000)):
        assert np.array_equal(oracle.synth_long_fasta(spec, threads=3), fk.synth_long_fasta(spec))
