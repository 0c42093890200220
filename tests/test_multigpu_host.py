"""CPU tests of the multi-GPU host logic (fastkmer_b200/multigpu.py): owner assignment and the exchange plan,
run as a real 2-rank job over the gloo backend with fake records standing in for the GPU stages."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastkmer_b200 import multigpu as mg


def test_assign_owners_is_lpt_and_deterministic():
    kmers = np.array([5, 100, 7, 7, 50, 0, 49, 1], dtype=np.uint64)
    owner = mg.assign_owners(kmers, 2)
    assert owner.tolist() == mg.assign_owners(kmers, 2).tolist()
    load = [int(kmers[owner == g].sum()) for g in range(2)]
    assert abs(load[0] - load[1]) <= 7 and sum(load) == int(kmers.sum())
    assert set(mg.assign_owners(kmers, 1).tolist()) == {0}
    big = np.random.default_rng(0).integers(0, 10**6, 2048).astype(np.uint64)
    o8 = mg.assign_owners(big, 8)
    l8 = np.array([big[o8 == g].sum() for g in range(8)], dtype=np.float64)
    assert l8.max() / l8.mean() < 1.01                       # LPT balances 2048 bins over 8 GPUs to < 1 %


def test_internal_bins_share_an_owner():
    """A job with internal bins (fkm_job_bins: `split` consecutive internal bins per bin of the configuration) exchanges per-internal-bin
    histograms; ownership is decided per bin of the configuration, by LPT over the bins' totals."""
    rng = np.random.default_rng(5)
    G, B, split = 4, 96, 8
    Hr = rng.integers(0, 50, (G, B * split))
    Hk = Hr * rng.integers(1, 30, (G, B * split))
    plans = [mg.plan_exchange(Hr, Hk, r, G, split=split) for r in range(G)]
    owner = plans[0]["owner"]
    assert owner.shape == (B * split,) and all((p["owner"] == owner).all() for p in plans)
    assert (owner.reshape(B, split) == owner.reshape(B, split)[:, :1]).all()          # one owner per bin
    assert owner.reshape(B, split)[:, 0].tolist() == mg.assign_owners(Hk.sum(axis=0).reshape(B, split).sum(axis=1), G).tolist()
    for r, p in enumerate(plans):                                                     # what r sends to g is what g expects from r
        for g in range(G):
            assert p["send_splits"][g] == plans[g]["recv_splits"][r]
        assert p["bin_rec"].sum() == Hr[:, owner == r].sum()
    fixed = np.arange(B, dtype=np.int32) % G                                          # a fixed per-bin map is expanded the same way
    assert (mg.plan_exchange(Hr, Hk, 0, G, owner=fixed, split=split)["owner"] == np.repeat(fixed, split)).all()


def test_cpp_plan_of_the_multi_gpu_job_equals_the_python_plan():
    """fkm_execute_job_multi plans its exchange in C++ (owners by LPT, owner-major send offsets, source-major receive offsets,
    segments back to bin-major); the hook fkm_debug_multi_plan needs no GPU: every field must equal multigpu.plan_exchange's."""
    import ctypes as C
    from fastkmer_b200 import api
    lib = api.load_library()
    rng = np.random.default_rng(11)
    for G, B, split in ((2, 64, 1), (3, 17, 1), (4, 96, 8), (8, 2048, 8), (5, 1, 4)):
        Bi = B * split
        Hr = rng.integers(0, 40, (G, Bi)).astype(np.uint64)
        Hr[:, rng.integers(0, Bi, max(1, Bi // 4))] = 0
        Hk = Hr * rng.integers(1, 30, (G, Bi)).astype(np.uint64)
        for r in range(G):
            want = mg.plan_exchange(Hr, Hk, r, G, split=split)
            owner = np.zeros(Bi, dtype=np.int32)
            send_base = np.zeros(Bi + 1, dtype=np.uint64); send_off = np.zeros(G + 1, dtype=np.uint64); recv_off = np.zeros(G + 1, dtype=np.uint64)
            bin_rec = np.zeros(Bi, dtype=np.uint64); bin_kmer = np.zeros(Bi, dtype=np.uint64)
            cap = G * Bi + 1
            seg_src = np.zeros(cap + 1, dtype=np.uint64); seg_dst = np.zeros(cap, dtype=np.uint64)
            n_seg = C.c_uint64()
            hr, hk = np.ascontiguousarray(Hr), np.ascontiguousarray(Hk)
            rc = lib.fkm_debug_multi_plan(G, B, split, hr.ctypes.data, hk.ctypes.data, r, owner.ctypes.data, send_base.ctypes.data, send_off.ctypes.data,
                                          recv_off.ctypes.data, bin_rec.ctypes.data, bin_kmer.ctypes.data, seg_src.ctypes.data, seg_dst.ctypes.data, cap, C.byref(n_seg))
            assert rc == 0
            what = "G=%d B=%d split=%d rank %d" % (G, B, split, r)
            assert owner.tolist() == want["owner"].tolist(), what
            assert send_base.tolist() == want["send_base"].tolist(), what
            assert np.diff(send_off).tolist() == want["send_splits"] and np.diff(recv_off).tolist() == want["recv_splits"], what
            assert bin_rec.tolist() == want["bin_rec"].tolist() and bin_kmer.tolist() == want["bin_kmer"].tolist(), what
            n = n_seg.value
            assert n == want["seg_dst"].size and seg_dst[:n].tolist() == want["seg_dst"].tolist() and seg_src[:n + 1].tolist() == want["seg_src"].tolist(), what


def test_plan_is_consistent_across_ranks():
    rng = np.random.default_rng(3)
    for G, B in ((2, 64), (3, 17), (8, 2048)):
        Hr = rng.integers(0, 40, (G, B))
        Hr[:, rng.integers(0, B, B // 4)] = 0                # empty bins
        Hk = Hr * rng.integers(1, 30, (G, B))
        plans = [mg.plan_exchange(Hr, Hk, r, G) for r in range(G)]
        for r, p in enumerate(plans):
            assert p["n_send"] == Hr[r].sum() and sum(p["send_splits"]) == p["n_send"]
            assert p["recv_splits"] == [plans[s]["send_splits"][r] for s in range(G)]
            assert p["seg_src"][-1] == p["n_recv"] == p["bin_rec"].sum()
            assert (p["owner"] == plans[0]["owner"]).all()
        assert sum(int(p["bin_kmer"].sum()) for p in plans) == Hk.sum()


def test_plan_matches_straightforward_loops():
    """plan_exchange is array arithmetic; the same plan written as loops over ranks and bins must give identical arrays."""
    rng = np.random.default_rng(5)
    for world, B in ((1, 7), (2, 64), (4, 300), (8, 2048)):
        H_rec = rng.integers(0, 50, size=(world, B)).astype(np.uint64) * (rng.random((world, B)) < 0.7)
        H_kmer = H_rec * rng.integers(1, 9, size=(world, B)).astype(np.uint64)
        for rank in range(world):
            p = mg.plan_exchange(H_rec, H_kmer, rank, world)
            owner = p["owner"]
            order = sorted(range(B), key=lambda b: (owner[b], b))
            off, send_base = 0, np.zeros(B + 1, dtype=np.uint64)
            for b in order:
                send_base[b] = off
                off += int(H_rec[rank][b])
            send_base[B] = off
            my_bins = [b for b in range(B) if owner[b] == rank]
            dst, acc = {}, 0
            for b in range(B):
                dst[b] = acc
                acc += int(H_rec[:, b].sum()) if owner[b] == rank else 0
            seg_src, seg_dst, filled = [0], [], {b: 0 for b in my_bins}
            for s_ in range(world):
                for b in my_bins:
                    n = int(H_rec[s_][b])
                    if n:
                        seg_dst.append(dst[b] + filled[b])
                        filled[b] += n
                        seg_src.append(seg_src[-1] + n)
            assert np.array_equal(p["send_base"], send_base)
            assert p["seg_src"].tolist() == seg_src and p["seg_dst"].tolist() == seg_dst
            assert p["send_splits"] == [int(H_rec[rank][owner == g].sum()) for g in range(world)]
            assert p["recv_splits"] == [int(H_rec[s_][my_bins].sum()) for s_ in range(world)]
            assert p["n_recv"] == acc and p["n_send"] == off


def test_fixed_owner_map_is_respected():
    rng = np.random.default_rng(5)
    Hr = rng.integers(0, 40, (3, 50))
    Hk = Hr * 7
    owner = np.arange(50) % 3
    plans = [mg.plan_exchange(Hr, Hk, r, 3, owner=owner) for r in range(3)]
    for r, p in enumerate(plans):
        assert (p["owner"] == owner).all()
        assert p["bin_rec"][owner != r].sum() == 0 and p["bin_rec"][owner == r].sum() == Hr[:, owner == r].sum()
        assert p["recv_splits"] == [plans[s]["send_splits"][r] for s in range(3)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        rec = rng.integers(0, 30, B).astype(np.int64)
        rec[rng.integers(0, B, B // 5)] = 0
        kmer = rec * rng.integers(1, 20, B)
        mine = torch.from_numpy(np.concatenate([rec, kmer]))
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        allh = torch.stack(gathered).numpy().reshape(world, 2, B)
        plan = mg.plan_exchange(allh[:, 0, :], allh[:, 1, :], rank, world)
        # fake "records": rows (source rank, bin, serial), laid out where fkm_mg_scatter would put them
        send = np.full((plan["n_send"], 3), -1, dtype=np.int64)
        for b in range(B):
            o = int(plan["send_base"][b])
            for i in range(int(rec[b])):
                send[o + i] = (rank, b, i)
        assert (send[:, 0] == rank).all()
        send_t = torch.from_numpy(send)
        recv_t = torch.empty((plan["n_recv"], 3), dtype=torch.int64)
        outs = [t.contiguous() for t in recv_t.split(plan["recv_splits"])]
        ins = list(send_t.split(plan["send_splits"]))
        mg.exchange_p2p(dist, outs, ins, rank)                # gloo has no alltoall; NCCL runs use all_to_all_single
        recv = torch.cat(outs).numpy() if outs else np.empty((0, 3), dtype=np.int64)
        # what fkm_mg_regroup does
        binned = np.full((plan["n_recv"], 3), -1, dtype=np.int64)
        for i in range(len(plan["seg_dst"])):
            a, e = int(plan["seg_src"][i]), int(plan["seg_src"][i + 1])
            d = int(plan["seg_dst"][i])
            binned[d:d + (e - a)] = recv[a:e]
        ok = True
        off = 0
        for b in range(B):
            n = int(plan["bin_rec"][b])
            blk = binned[off:off + n]
            if plan["owner"][b] == rank:
                want = sorted((s, b, i) for s in range(world) for i in range(int(allh[s, 0, b])))
                ok &= sorted(map(tuple, blk.tolist())) == want
            else:
                ok &= n == 0
            off += n
        ok &= off == plan["n_recv"]
        out_q.put((rank, bool(ok), int(plan["bin_kmer"].sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_over_gloo(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 97, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)


def test_split_by_sample_uses_the_leading_word_of_the_header():
    fasta = (b"junk\n>SRR1.1 HW len=3\nACG\nTTT\n>SRR2.1 x\nGGG\n>SRR1.2\nCCC\n>  odd_tag-7 z\nAAA\n>\nTT\n")
    parts = mg.split_by_sample(fasta)
    assert list(parts) == ["SRR1", "SRR2", "odd_tag", ""]                      # order of first appearance
    assert parts["SRR1"] == b">SRR1.1 HW len=3\nACG\nTTT\n>SRR1.2\nCCC\n"
    assert parts["SRR2"] == b">SRR2.1 x\nGGG\n"
    assert parts["odd_tag"] == b">  odd_tag-7 z\nAAA\n"
    assert sum(len(v) for v in parts.values()) == len(fasta) - len(b"junk\n")


def test_multisequence_configuration_mirror():
    from fastkmer_b200.multisequence import MultisequenceTestConfiguration
    tc = MultisequenceTestConfiguration("in.fa", "/out/", 28, 10, 3, max_b=2048, sequenceType=1)
    assert tc.b == 2048 and tc.outputDir == "/out/k28_m10_x3_b2048_s1"       # multisequence/package.scala:28-29 (no prefix)
    cc = tc.counting_configuration()
    assert (cc.k, cc.m, cc.x, cc.max_b, cc.useHT, cc.sequenceType) == (28, 10, 3, 2048, False, 1)


def test_bench_workload_shards_cover_the_input_once():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = bench.Workload(3)                                                      # long sequence: position ranges + (k-1) halo
    k = wl.c["k"]
    for world in (1, 2, 4, 8):
        shards = [wl.spec(r, world) for r in range(world)]
        assert shards[0]["first_pos"] == 0
        starts = 0
        for r, sh in enumerate(shards):
            halo = k - 1 if r < world - 1 else 0
            starts += sh["n_bases"] - halo                                      # window starts owned by the shard
            if r + 1 < world:
                assert shards[r + 1]["first_pos"] == sh["first_pos"] + sh["n_bases"] - halo
        assert starts == wl.c["n_bases"]
        assert wl.n_bases_total(world) == wl.c["n_bases"] and wl.scaling == "strong"
    w2 = bench.Workload(2)
    for world in (1, 8):
        sh = [w2.spec(r, world) for r in range(world)]
        assert [s_["first_read"] for s_ in sh] == [r * w2.c["reads"] for r in range(world)]
        assert len({s_["genome_len"] for s_ in sh}) == 1 and w2.scaling == "weak"
        assert w2.n_bases_total(world) == world * w2.c["reads"] * w2.c["L"]
