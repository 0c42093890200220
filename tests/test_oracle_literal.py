"""The literal transliteration of the Scala (oracle/literal.py: Kmer with 31-nt Long slices, readFromKmer,
RIndex heap merge, ...) against the fast C++ oracle and the SURVEY App. C digests.  CPU only."""
import hashlib
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import literal
import oracle_lib
from test_oracle_kat import GOLD, _rand_fasta


def _oracle_triples(res, k):
    return [(int(b), oracle_lib.kmer_str(h, l, k), int(c)) for b, h, l, c in zip(res["bin"], res["hi"], res["lo"], res["cnt"])]


def test_literal_helpers_match_oracle(oracle):
    for m in (3, 4, 5, 7):
        assert literal.fillNorm(m) == oracle.fill_norm(m).tolist()
    for s, B in ((0, 2048), (1, 2048), (12345, 2048), (1048576, 2048), (4194303, 4096), (67108864, 2048)):
        assert literal.hash_to_bucket(s, B) == oracle.hash_to_bucket(s, B)
    rng = random.Random(1)
    for L in (5, 30, 31, 32, 61, 62, 63, 93, 94):              # slice layout and string round trip (UTIL:144-172, 416-454)
        s = "".join(rng.choice("ACGT") for _ in range(L)).encode()
        km = literal.Kmer.fromBytes(L, s, 0)
        assert km.toByteArray() == s
        assert all(0 <= w < (1 << 62) for w in km._data)
        for a in range(0, L, 7):                               # every readFromKmer branch, both orientations (UTIL:174-295)
            for b in range(a, L, 5):
                sub = literal.Kmer.fromKmer(b - a + 1, km, a, b, 0)
                assert sub.toByteArray() == s[a:b + 1]
                rc = literal.Kmer.fromKmer(b - a + 1, km, a, b, 1)
                comp = bytes(b"TGCA"[b"ACGT".index(c)] for c in reversed(s[a:b + 1]))
                assert rc.toByteArray() == comp


@pytest.mark.parametrize("gid", sorted(GOLD))
@pytest.mark.parametrize("use_ht", [0, 1])
def test_literal_reproduces_golden_digests(oracle, gid, use_ht):
    """The Scala, transliterated line by line, gives the SHA-256 digests of SURVEY App. C.4 on both reduce paths —
    and, entry for entry, what the fast C++ oracle gives."""
    (seed, G, R, L), (k, m, x, B), nk, dist, nbins, maxc, nsk, skb, sha = GOLD[gid]
    fasta = oracle.gen_lcg_fasta(seed, G, R, L)
    triples, per_bin = literal.count(fasta, k, m, x, B, use_ht)
    assert triples == _oracle_triples(oracle.count(fasta, k, m, x, B, use_ht), k)
    lines = sorted(b"%d\t%s\t%d\n" % (b, s.encode(), c) for b, s, c in triples)
    assert hashlib.sha256(b"".join(lines)).hexdigest() == sha
    assert (sum(c for _, _, c in triples), len(triples), len(per_bin), max(c for _, _, c in triples)) == (nk, dist, nbins, maxc)
    if gid == "G4" and not use_ht:
        assert per_bin[0][:5] == [("CTGAC", 3), ("CTGCA", 10), ("CTGGA", 13), ("CTGTC", 18), ("GGAAC", 16)]
    if not use_ht:                                            # the sort path writes ascending k-mers (SBKC:566-597)
        for lines_ in per_bin.values():
            assert [s for s, _ in lines_] == sorted(s for s, _ in lines_)


@pytest.mark.parametrize("k,m,x", [(5, 3, 1), (12, 4, 2), (28, 10, 3), (31, 11, 3), (32, 7, 2), (33, 8, 3), (55, 13, 3),
                                   (60, 9, 4), (61, 6, 3)])
def test_literal_vs_oracle_random(oracle, k, m, x):
    rng = random.Random(k * 31 + m)
    fasta = _rand_fasta(rng, 8, 0, 2 * k + 50, width=23).encode()
    for use_ht in (0, 1):
        got, _ = literal.count(fasta, k, m, x, 777, use_ht)
        want = _oracle_triples(oracle.count(fasta, k, m, x, 777, use_ht), k)
        assert got == want


def test_literal_beyond_the_oracles_width():
    """k + x > 64 (reference supports any k): literal HT path == literal sort path == clean spec."""
    import clean_spec
    rng = random.Random(9)
    fasta = _rand_fasta(rng, 5, 60, 260, p_bad=0.01)
    for k, m, x in ((62, 10, 3), (63, 12, 3), (70, 9, 2), (93, 13, 1)):
        a, _ = literal.count(fasta.encode(), k, m, x, 500, 1)
        b, _ = literal.count(fasta.encode(), k, m, x, 500, 0)
        want = sorted((bn, s, c) for (bn, s), c in clean_spec.count(fasta, k, m, 500).items())
        assert a == b == want
