"""Parity of the hash path under a context knob:  python scripts/knob_check.py cas_first=1 [name=value ...]
Counts a few synthetic inputs (64- and 128-bit keys, deep and shallow coverage) with the knobs set and compares
every (bin, k-mer, count) with the CPU oracle (test infrastructure).  Needs a GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fastkmer_b200 as fk   # noqa: E402
import oracle_lib            # noqa: E402


def main():
    oracle = oracle_lib.load()
    ctx = fk.Context(0)
    for kv in sys.argv[1:]:
        name, value = kv.split("=")
        ctx.set(name, float(value))
    cases = [("G1", oracle.gen_lcg_fasta(42, 2000, 200, 100)),
             ("deep", fk.synth_fasta(dict(seeds=(7, 8, 9), genome_len=30000, n_reads=60000, read_len=100)).tobytes()),
             ("shallow", fk.synth_fasta(dict(seeds=(17, 18, 19), genome_len=2000000, n_reads=100000, read_len=150)).tobytes())]
    for label, fasta in cases:
        for k, m, B in ((28, 10, 2048), (55, 13, 2048), (31, 11, 64), (64, 15, 1)):
            cfg = fk.TestConfiguration("", "", k, m, 3, max_b=B, useHT=True, write=False)
            want = oracle.count(fasta, k, m, 3, B, 1, threads=8)
            res, st = ctx.count_fasta(cfg, fasta)
            got = res.sorted_arrays()
            for key in ("bin", "hi", "lo", "cnt"):
                assert np.array_equal(got[key], want[key]), "%s k=%d: differs from the oracle in %s" % (label, k, key)
            assert st["digest_sum"] == want["stats"]["digest_sum"] and st["n_fallbacks"] == 0
            print("ok %-8s k=%d m=%d B=%d: %d k-mers, %d distinct, %d batches" % (label, k, m, B, st["n_kmers"], st["n_distinct"], st["n_batches"]))
    ctx.close()


if __name__ == "__main__":
    main()
