"""N-rank parity check under torchrun (NCCL): every rank counts its shard through ShardedJob, rank 0 merges the
per-rank results and compares them with the CPU oracle on the whole read set.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/mg_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fastkmer_b200 as fk
from fastkmer_b200 import multigpu


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = fk.Context(local, stream.cuda_stream)
    ok = True
    for (k, m, ht, reads) in ((28, 10, 1, 40000), (28, 10, 0, 40000), (55, 13, 1, 20000), (31, 11, 0, 20000)):
        spec = dict(seeds=(51, 52, 53), genome_len=300000, n_reads=reads, read_len=150)
        per = reads // world
        mine = dict(spec, n_reads=per if rank < world - 1 else reads - per * (world - 1), first_read=rank * per)
        cfg = fk.TestConfiguration("", "", k, m, 3, max_b=2048, useHT=bool(ht), write=False)
        job = multigpu.ShardedJob(ctx, cfg, dist, rank, world)
        d_b, d_i, n_pos = ctx.synth_packed_device(mine)
        res, st = job.count_packed_device(d_b, d_i, n_pos, want_result=True)
        a = res.arrays()
        fasta = fk.synth_fasta(mine)
        res2, st2 = job.count_fasta(fasta, want_result=True)          # host-FASTA path of the same shard
        same_paths = st2["digest_sum_global"] == st["digest_sum_global"] and st2["n_kmers_global"] == st["n_kmers_global"]
        gathered = [None] * world if rank == 0 else None
        dist.gather_object({k_: v for k_, v in a.items()}, gathered, dst=0)
        if rank == 0:
            import oracle_lib
            oracle = oracle_lib.load()
            want = oracle.count(fk.synth_fasta(spec).tobytes(), k, m, 3, 2048, ht, threads=os.cpu_count())
            got = {k_: np.concatenate([g[k_] for g in gathered]) for k_ in ("bin", "hi", "lo", "cnt")}
            order = np.lexsort((got["lo"], got["hi"], got["bin"]))
            good = all(np.array_equal(got[k_][order], want[k_]) for k_ in got)
            good &= st["n_kmers_global"] == want["stats"]["n_kmers"] and st["digest_sum_global"] == want["stats"]["digest_sum"]
            good &= st["digest_xor_global"] == want["stats"]["digest_xor"] and same_paths
            print("mg_check world=%d k=%d m=%d useHT=%d: %s (%d distinct, exchange %.3f ms)"
                  % (world, k, m, ht, "OK" if good else "MISMATCH", want["stats"]["n_distinct"], st["exchange_ms"]))
            ok &= bool(good)
        ctx.free_device(d_b)
        ctx.free_device(d_i)
    # long-sequence mode (BASELINE config 3): one genome split by position range with a (k-1)-base halo per shard
    for (k, m, ht) in ((31, 11, 0), (31, 11, 1)):
        n_bases = 3_000_000
        seeds = (3001, 3002, 3003)
        per = n_bases // world
        first = rank * per
        n = (per if rank < world - 1 else n_bases - first) + (k - 1 if rank < world - 1 else 0)
        cfg = fk.TestConfiguration("", "", k, m, 3, max_b=4096, useHT=bool(ht), write=False, sequenceType=1)
        job = multigpu.ShardedJob(ctx, cfg, dist, rank, world)
        d_b, d_i, n_pos = ctx.synth_long_packed_device(dict(seeds=seeds, first_pos=first, n_bases=n))
        res, st = job.count_packed_device(d_b, d_i, n_pos, want_result=True)
        a = res.arrays()
        gathered = [None] * world if rank == 0 else None
        dist.gather_object({k_: v for k_, v in a.items()}, gathered, dst=0)
        if rank == 0:
            import oracle_lib
            oracle = oracle_lib.load()
            want = oracle.count(fk.synth_long_fasta(dict(seeds=seeds, n_bases=n_bases)).tobytes(), k, m, 3, 4096, ht, threads=os.cpu_count())
            got = {k_: np.concatenate([g[k_] for g in gathered]) for k_ in ("bin", "hi", "lo", "cnt")}
            order = np.lexsort((got["lo"], got["hi"], got["bin"]))
            good = all(np.array_equal(got[k_][order], want[k_]) for k_ in got) and st["n_kmers_global"] == want["stats"]["n_kmers"]
            print("mg_check world=%d long sequence k=%d useHT=%d: %s (%d k-mers)" % (world, k, ht, "OK" if good else "MISMATCH", want["stats"]["n_kmers"]))
            ok &= bool(good)
        ctx.free_device(d_b)
        ctx.free_device(d_i)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
