"""tests/golden/*.tsv — expected bin files produced by the line-by-line transliteration of the reference's Scala
(tests/golden/make_golden.py) — against the C++ oracle (CPU) and the CUDA library through its C ABI (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "MANIFEST.json")) as f:
    VECTORS = json.load(f)["vectors"]
IDS = [v["expected"][:-4] for v in VECTORS]


def load(v):
    with open(os.path.join(GOLDEN, v["fasta"]), "rb") as f:
        fasta = f.read()
    with open(os.path.join(GOLDEN, v["expected"]), "rb") as f:
        raw = f.read()
    want = [(int(b), s.decode(), int(c)) for b, s, c in (line.split(b"\t") for line in raw.splitlines())]
    return fasta, raw, want


def triples(res, k):
    return [(int(b), oracle_lib.kmer_str(h, l, k), int(c)) for b, h, l, c in zip(res["bin"], res["hi"], res["lo"], res["cnt"])]


@pytest.mark.parametrize("v", VECTORS, ids=IDS)
def test_fixture_is_intact(v):
    fasta, raw, want = load(v)
    assert hashlib.sha256(raw).hexdigest() == v["sha256"]
    assert (sum(c for _, _, c in want), len(want)) == (v["n_kmers"], v["n_distinct"])
    assert want == sorted(want)                               # bins ascending, k-mers ascending inside a bin (SBKC:566-597)
    assert all(len(s) == v["k"] and set(s) <= set("ACGT") for _, s, _ in want)


@pytest.mark.parametrize("v", VECTORS, ids=IDS)
def test_oracle_reproduces_golden_files(oracle, v):
    fasta, raw, want = load(v)
    for use_ht in (0, 1):
        if not use_ht and v["k"] + v["x"] > 64:
            continue                                          # the oracle's (k,x)-mers are limited to 64 symbols; the literal and CUDA paths are not
        res = oracle.count(fasta, v["k"], v["m"], v["x"], v["max_b"], use_ht)
        assert triples(res, v["k"]) == want
        assert (res["stats"]["n_kmers"], res["stats"]["n_distinct"]) == (v["n_kmers"], v["n_distinct"])


@pytest.mark.gpu
@pytest.mark.parametrize("v", VECTORS, ids=IDS)
def test_cuda_reproduces_golden_files(v, tmp_path):
    import fastkmer_b200 as fk
    fasta, raw, want = load(v)
    ctx = fk.Context(0)
    try:
        for use_ht in (0, 1):
            cfg = fk.TestConfiguration("", "", v["k"], v["m"], v["x"], max_b=v["max_b"], useHT=bool(use_ht), write=False)
            res, st = ctx.count_fasta(cfg, fasta)
            assert triples(res.sorted_arrays(), v["k"]) == want, "useHT=%d" % use_ht
            assert (st["n_kmers"], st["n_distinct"], st["total_count"]) == (v["n_kmers"], v["n_distinct"], v["n_kmers"])
            if not use_ht:
                assert triples(res.arrays(), v["k"]) == want  # device order == file order
        # the hash path with its tables in shared memory (k_count_smem) instead of global memory
        for mode in (2, 1):      # 2: hash-partitioned k-mers (k_count_keys), 1: dual-minimizer mid bins (k_count_smem)
            ctx.set("count_mode", mode)
            cfg = fk.TestConfiguration("", "", v["k"], v["m"], v["x"], max_b=v["max_b"], useHT=True, write=False)
            res, st = ctx.count_fasta(cfg, fasta)
            assert triples(res.sorted_arrays(), v["k"]) == want, "shared-memory tables, mode %d" % mode
            assert (st["n_kmers"], st["n_distinct"], st["total_count"]) == (v["n_kmers"], v["n_distinct"], v["n_kmers"]) and st["n_mid_bins"] > 0
        ctx.set("count_mode", 0)
        # the job as the reference runs it: dataset file in, <outputDir>/bin<id> files out (sorted + "EOF", SBKC:550-606)
        src = tmp_path / "in.fasta"
        src.write_bytes(fasta)
        tc = fk.TestConfiguration(str(src), str(tmp_path) + "/", v["k"], v["m"], v["x"], max_b=v["max_b"], useHT=False, write=True, prefix="g_")
        fk.SparkBinKmerCounter.executeJob(ctx, tc)
        by_bin = {}
        for b, s, c in want:
            by_bin.setdefault(b, []).append("%s\t%d\n" % (s, c))
        out_dir = tc.outputDir
        assert sorted(os.listdir(out_dir)) == sorted("bin%d" % b for b in by_bin)
        for b, lines in by_bin.items():
            with open(os.path.join(out_dir, "bin%d" % b)) as f:
                assert f.read() == "".join(lines) + "EOF"     # no newline after the trailer (SBKC:598-606)
    finally:
        ctx.close()
