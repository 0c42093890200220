"""N-rank check of the multi-sample distances (BASELINE config 5 shape) under torchrun: the sharded job must give the
matrix of the single-GPU fkm_multiseq_fasta on the whole input.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/mg_multiseq_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastkmer_b200 as fk
from fastkmer_b200 import multigpu


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = fk.Context(local, stream.cuda_stream)
    S, R, L = 4, 20000, 100
    parts = []
    for s_ in range(S):
        t = fk.synth_fasta(dict(seeds=(5001, 5100 + s_, 5200 + s_), genome_len=200000, n_reads=R, read_len=L)).tobytes()
        parts.append(t.replace(b">r", b">S%d." % s_))
    whole = b"".join(parts)
    recs = whole.split(b">")[1:]
    mine = b"".join(b">" + r for i, r in enumerate(recs) if i % world == rank)       # reads dealt round-robin to the ranks
    ok = True
    for k, m in ((28, 10), (55, 13)):
        cfg = fk.TestConfiguration("", "", k, m, 3, max_b=2048, useHT=False, write=False)
        names, d = multigpu.multiseq_sharded(ctx, cfg, dist, rank, world, mine)
        if rank == 0:
            want_names, want, _, _ = ctx.multiseq_fasta(cfg, whole)
            order = [want_names.index(n) for n in names]
            good = np.array_equal(d, want[np.ix_(order, order)]) and d.max() > 0
            print("mg_multiseq_check world=%d k=%d: %s (dist[0,1]=%.0f)" % (world, k, "OK" if good else "MISMATCH", d[0, 1]))
            ok &= bool(good)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
