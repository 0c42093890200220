"""BASELINE config 5 shape on one GPU: S samples x R reads x 100 bp, k=28 m=10 B=2048, squared-euclidean distances
between the samples' k-mer count vectors through fkm_multiseq_fasta (host FASTA in, S x S matrix out).

Samples here are independent read sets (own read / error seeds) over one synthetic genome; headers are 'S<s>.<r>'.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastkmer_b200 as fk


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    L = 100
    parts = []
    for s_ in range(S):
        t = fk.synth_fasta(dict(seeds=(5001, 5100 + s_, 5200 + s_), genome_len=5_000_000, n_reads=R, read_len=L)).tobytes()
        parts.append(t.replace(b">r", b">S%d." % s_))
    fasta = np.frombuffer(b"".join(parts), dtype=np.uint8)
    ctx = fk.Context(0)
    cfg = fk.TestConfiguration("", "", 28, 10, 3, max_b=2048, useHT=False, write=False)
    times = []
    for it in range(4):
        t0 = time.perf_counter()
        names, dist, _, st = ctx.multiseq_fasta(cfg, fasta, max_samples=64)
        times.append(time.perf_counter() - t0)
    dt = min(times[1:])
    print(json.dumps({"workload": "config 5 shape: %d samples x %d x %d bp, k=28 m=10 B=2048" % (S, R, L), "seconds": dt,
                      "bases_per_sec": S * R * L / dt, "kmers_per_sec": st["n_kmers"] / dt, "n_samples": len(names),
                      "dist_0_1": dist[0, 1], "dist_max": float(dist.max()), "launches": st["gpu_launches"]}))
    ctx.close()


if __name__ == "__main__":
    main()
