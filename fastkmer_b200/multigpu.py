"""Multi-GPU k-mer counting: one process per GPU, reads split by range, bins owned per GPU.

This replaces Spark's shuffle between the map and reduce stages of the reference
(`reduceByKey(_ ++ _)`, SparkBinKmerCounter.scala:1035,1042, optionally behind the LPT
`MultiprocessorSchedulingPartitioner`, MultiprocessorSchedulingPartitioner.scala:35-69):

  1. every rank scans its own shard of the reads               (fkm_mg_scan: exact per-bin histogram)
  2. the B-entry histograms are all-gathered                   (torch.distributed, <= 64 KB per rank)
  3. bins are assigned to GPUs by LPT over the exact k-mer counts (plan_exchange; same on every rank)
  4. each rank writes its records owner-major                  (fkm_mg_scatter)
  5. ONE variable-size all-to-all moves the records            (dist.all_to_all_single over NCCL / NVLink)
  6. received records are regrouped bin-major                  (fkm_mg_regroup)
  7. every rank counts the bins it owns                        (fkm_mg_count) — embarrassingly parallel, because a
     canonical k-mer's signature, hence bin, is strand- and position-independent

plan_exchange() is pure host logic (numpy) and is covered by CPU tests with the gloo backend;
emulate_ranks() runs all ranks' stages on ONE GPU (no collective) for the -m gpu tests.
"""
import numpy as np


def assign_owners(bin_kmers, world):
    """LPT (longest processing time first): bins by decreasing total k-mers, each to the least-loaded GPU.
    Deterministic (ties by bin id, then by rank), so every rank computes the same map."""
    import heapq
    bin_kmers = np.asarray(bin_kmers, dtype=np.uint64)
    order = np.lexsort((np.arange(bin_kmers.size), -bin_kmers.astype(np.int64)))
    heap = [(0, r) for r in range(world)]
    owner = np.zeros(bin_kmers.size, dtype=np.int32)
    for b, w in zip(order.tolist(), bin_kmers[order].tolist()):
        load, g = heapq.heappop(heap)
        owner[b] = g
        heapq.heappush(heap, (load + w, g))
    return owner


def plan_exchange(H_rec, H_kmer, rank, world, owner=None, split=1):
    """H_rec, H_kmer: [world, B] records / k-mers every rank puts into every bin.  owner: a fixed bin -> GPU map
    (jobs that must agree on the ownership, e.g. the samples of a distance job); default LPT on this job's histogram.
    split: B counts INTERNAL bins, `split` consecutive ones per bin of the configuration (fkm_job_bins); a bin's
    internal bins always share an owner.
    -> dict with everything rank `rank` needs for steps 4-7.  Array arithmetic only (no loop over bins x ranks)."""
    H_rec = np.asarray(H_rec, dtype=np.uint64)
    H_kmer = np.asarray(H_kmer, dtype=np.uint64)
    B = H_rec.shape[1]
    if owner is None:
        owner = np.repeat(assign_owners(H_kmer.sum(axis=0).reshape(B // split, split).sum(axis=1, dtype=np.uint64), world), split)
    else:
        owner = np.asarray(owner, dtype=np.int32)
        if owner.size * split == B and split > 1:
            owner = np.repeat(owner, split)
    # send buffer: bins ordered by (owner, bin)
    order = np.lexsort((np.arange(B), owner))
    mine = H_rec[rank]
    csum = np.cumsum(mine[order], dtype=np.uint64)
    send_base = np.zeros(B + 1, dtype=np.uint64)
    send_base[order] = csum - mine[order]
    send_base[B] = csum[-1] if B else 0
    send_splits = [int(mine[owner == g].sum(dtype=np.uint64)) for g in range(world)]
    my_bins = np.nonzero(owner == rank)[0]
    M = H_rec[:, my_bins]                                      # [source, my bin] records this rank receives
    recv_splits = [int(v) for v in M.sum(axis=1, dtype=np.uint64)]
    # bin-major layout of the bins this rank owns
    bin_rec = np.zeros(B, dtype=np.uint64)
    bin_kmer = np.zeros(B, dtype=np.uint64)
    bin_rec[my_bins] = M.sum(axis=0, dtype=np.uint64)
    bin_kmer[my_bins] = H_kmer[:, my_bins].sum(axis=0, dtype=np.uint64)
    dst_base = np.zeros(B + 1, dtype=np.uint64)
    dst_base[1:] = np.cumsum(bin_rec, dtype=np.uint64)
    # receive buffer is source-major, each source's block holds this rank's bins in bin order: segment (s, b) moves to
    # the bin's place + what the sources before s put into the bin
    flat = M.ravel()
    keep = flat > 0
    before = np.cumsum(M, axis=0, dtype=np.uint64) - M
    seg_dst = (dst_base[my_bins][None, :] + before).ravel()[keep]
    seg_src = np.concatenate([np.zeros(1, dtype=np.uint64), np.cumsum(flat[keep], dtype=np.uint64)])
    return dict(owner=owner, send_base=send_base, send_splits=send_splits, recv_splits=recv_splits, bin_rec=bin_rec,
                bin_kmer=bin_kmer, seg_src=seg_src.astype(np.uint64), seg_dst=seg_dst.astype(np.uint64),
                n_send=int(send_base[B]), n_recv=int(sum(recv_splits)))


def exchange_p2p(dist, outs, ins, rank):
    """Variable-size all-to-all as point-to-point sends (for backends without alltoall, e.g. gloo on CPU).
    outs[s] receives what rank s sends here; ins[g] goes to rank g."""
    outs[rank].copy_(ins[rank])
    reqs = []
    for peer in range(len(ins)):
        if peer == rank:
            continue
        if ins[peer].numel():
            reqs.append(dist.isend(ins[peer].contiguous(), peer))
        if outs[peer].numel():
            reqs.append(dist.irecv(outs[peer], peer))
    for r in reqs:
        r.wait()


class ShardedJob:
    """One rank of a multi-GPU job (torch.distributed must be initialised with the NCCL backend)."""

    def __init__(self, ctx, configuration, dist, rank, world):
        import torch
        self.torch, self.ctx, self.cfg, self.dist, self.rank, self.world = torch, ctx, configuration, dist, rank, world
        from . import api
        self.rec_bytes = api.record_bytes(configuration)
        self.fixed_owner = None           # set to a bin -> GPU map to override the per-job LPT assignment
        self.last_plan = None
        self.last_exchange_ms = 0.0
        # A rank owns 1/world of the bins but receives them from every rank: the hash path cuts every bin into `world` (rounded up
        # to a power of two) internal bins, so that a bin of this job is as large as a bin of a one-GPU job of the same shard
        # size.  Same value on every rank (the histograms that are exchanged are per internal bin).
        self.split = 1
        if configuration.useHT:
            while self.split < world and self.split < 64:
                self.split *= 2
            ctx.set("bin_split", self.split)

    def count_packed_device(self, d_bases, d_inv, n_positions, want_result=False):
        rec, kmer = self.ctx.mg_scan(self.cfg, d_bases, d_inv, n_positions)
        return self._exchange_and_count(rec, kmer, want_result)

    def count_fasta(self, fasta, want_result=False):
        """This rank's shard as FASTA text in (pinned) host memory."""
        rec, kmer, n_bases = self.ctx.mg_scan_fasta(self.cfg, fasta)
        out = self._exchange_and_count(rec, kmer, want_result)
        st = out[1] if want_result else out
        st["n_bases"] = n_bases
        st["h2d_bytes"] = st["h2d_bytes"] + int(getattr(fasta, "size", len(fasta)))
        return out

    def _exchange_and_count(self, rec, kmer, want_result):
        torch, dist = self.torch, self.dist
        B = rec.size                                                  # internal bins (fkm_job_bins)
        split = B // self.cfg.b
        mine = torch.from_numpy(np.concatenate([rec, kmer]).astype(np.int64)).cuda()
        allh = torch.empty(self.world * 2 * B, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allh, mine)
        allh = allh.cpu().numpy().reshape(self.world, 2, B).astype(np.uint64)
        plan = plan_exchange(allh[:, 0, :], allh[:, 1, :], self.rank, self.world, owner=self.fixed_owner, split=split)
        self.last_plan = plan
        send = torch.empty((max(plan["n_send"], 1), self.rec_bytes), dtype=torch.uint8, device="cuda")
        self.ctx.mg_scatter(plan["send_base"], send.data_ptr())
        recv = torch.empty((max(plan["n_recv"], 1), self.rec_bytes), dtype=torch.uint8, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_to_all_single(recv[:plan["n_recv"]], send[:plan["n_send"]], plan["recv_splits"], plan["send_splits"])
        e1.record()
        torch.cuda.synchronize()
        self.last_exchange_ms = e0.elapsed_time(e1)
        d_records = self.ctx.mg_regroup(self.cfg, recv.data_ptr(), plan["n_recv"], plan["seg_src"], plan["seg_dst"])
        res, st = self.ctx.mg_count(self.cfg, d_records, plan["bin_rec"], plan["bin_kmer"], want_result=want_result)
        # job-wide totals
        tot = torch.tensor([st["n_kmers"], st["n_distinct"], st["total_count"], st["n_superkmers"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        dig = torch.tensor([st["digest_sum"] - (1 << 64) if st["digest_sum"] >= (1 << 63) else st["digest_sum"],
                            st["digest_xor"] - (1 << 64) if st["digest_xor"] >= (1 << 63) else st["digest_xor"]],
                           dtype=torch.int64, device="cuda")
        alld = torch.empty(self.world * 2, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(alld, dig)
        alld = alld.cpu().numpy().reshape(self.world, 2).astype(np.uint64)
        st["n_kmers_global"], st["n_distinct_global"], st["total_count_global"], st["n_superkmers_global"] = [int(v) for v in tot.tolist()]
        st["digest_sum_global"] = int(alld[:, 0].sum(dtype=np.uint64))
        st["digest_xor_global"] = int(np.bitwise_xor.reduce(alld[:, 1]))
        st["exchange_ms"] = self.last_exchange_ms
        st["exchange_bytes_sent"] = (plan["n_send"] - plan["send_splits"][self.rank]) * self.rec_bytes
        if want_result:
            return res, st
        return st


def split_by_sample(fasta: bytes):
    """FASTA text -> ordered {sample tag: text of its records}; the tag is the leading \\w+ of the header LINE
    (SparkMultiSequenceKmerCounter.scala:61-62; the reference's `(\\w+).` full match also keeps
    the one character after the word -- a label difference only).  Same rule as fkm_multiseq_fasta."""
    import re
    out = {}
    for rec in re.split(rb"(?m)^(?=>)", fasta):
        if not rec.startswith(b">"):
            continue
        m_ = re.match(rb">[^\w\n]*(\w+)", rec)
        tag = m_.group(1).decode() if m_ else ""
        out.setdefault(tag, []).append(rec if rec.endswith(b"\n") else rec + b"\n")
    return {t: b"".join(v) for t, v in out.items()}


def multiseq_sharded(ctx, configuration, dist, rank, world, fasta_shard: bytes):
    """Multi-sample squared-euclidean distances over `world` GPUs (BASELINE config 5): every rank holds a shard of
    the reads; per sample the bins are exchanged and counted as in ShardedJob, every rank computes the partial
    sums over the bins it owns (fkm_result_dot), and ONE all-reduce of the S x S matrix finishes the job
    (SparkMultiSequenceKmerCounter.scala:458-520, multiseq/SquaredEuclidean.java:19-32).
    -> (sample names, S x S float64 matrix), identical on every rank."""
    import torch
    from dataclasses import replace
    cfg = replace(configuration, useHT=False, write=False)
    mine = split_by_sample(fasta_shard)
    tags_all = [None] * world
    dist.all_gather_object(tags_all, list(mine))
    names = sorted(set(t for tl in tags_all for t in tl))
    job = ShardedJob(ctx, cfg, dist, rank, world)
    # all samples must agree on who owns a bin (the partial sums pair bin b of sample a with bin b of sample b):
    # bins are hash buckets of the signature, so a static round-robin map is balanced enough
    job.fixed_owner = np.arange(cfg.b, dtype=np.int32) % world
    clones = []
    for t in names:                                               # every rank runs every sample's job (possibly on no reads)
        text = np.frombuffer(mine.get(t, b""), dtype=np.uint8)
        res, _ = job.count_fasta(text, want_result=True)
        clones.append(res.clone())
    S = len(names)
    part = np.zeros((S, S), dtype=np.uint64)
    for a in range(S):
        for b in range(a, S):
            part[a, b] = clones[a].dot(clones[b])
    for c in clones:
        c.free()
    # uint64 sums through two int64 halves (NCCL has no uint64 sum)
    lo = torch.from_numpy((part & np.uint64(0xFFFFFFFF)).astype(np.int64)).cuda()
    hi = torch.from_numpy((part >> np.uint64(32)).astype(np.int64)).cuda()
    dist.all_reduce(lo)
    dist.all_reduce(hi)
    dot = [[(int(hi[a, b]) << 32) + int(lo[a, b]) for b in range(S)] for a in range(S)]
    out = np.zeros((S, S), dtype=np.float64)
    for a in range(S):
        for b in range(a + 1, S):
            out[a, b] = out[b, a] = float(dot[a][a] + dot[b][b] - 2 * dot[a][b])
    return names, out


def emulate_ranks(ctx, configuration, shards, world):
    """All `world` ranks of a job on ONE GPU, stage by stage, without a collective (test helper).
    shards: list of (d_bases, d_inv, n_positions).  -> (merged sorted arrays dict, list of per-rank stats)"""
    import torch
    from . import api
    rb = api.record_bytes(configuration)
    hists = [ctx.mg_scan(configuration, *sh) for sh in shards]       # first pass only for the histograms
    H_rec = np.stack([h[0] for h in hists])
    H_kmer = np.stack([h[1] for h in hists])
    plans = [plan_exchange(H_rec, H_kmer, r, world) for r in range(world)]
    sends = []
    for r, sh in enumerate(shards):
        ctx.mg_scan(configuration, *sh)                                # a job per rank (the scan state lives in the context)
        send = torch.empty((max(plans[r]["n_send"], 1), rb), dtype=torch.uint8, device="cuda")
        ctx.mg_scatter(plans[r]["send_base"], send.data_ptr())
        sends.append(send)
    merged = {"bin": [], "hi": [], "lo": [], "cnt": []}
    stats = []
    for r in range(world):
        parts = []
        for s_ in range(world):                                        # what all_to_all_single would deliver to rank r
            o = sum(plans[s_]["send_splits"][:r])
            parts.append(sends[s_][o:o + plans[s_]["send_splits"][r]])
        recv = torch.cat(parts) if parts else torch.empty((0, rb), dtype=torch.uint8, device="cuda")
        assert recv.shape[0] == plans[r]["n_recv"]
        assert [p.shape[0] for p in parts] == plans[r]["recv_splits"]
        recv = recv.contiguous()
        ctx.mg_scan(configuration, shards[r][0], shards[r][1], 0)      # fresh job on the context (empty scan)
        d_rec = ctx.mg_regroup(configuration, recv.data_ptr(), plans[r]["n_recv"], plans[r]["seg_src"], plans[r]["seg_dst"])
        res, st = ctx.mg_count(configuration, d_rec, plans[r]["bin_rec"], plans[r]["bin_kmer"])
        a = res.arrays()
        for k_ in merged:
            merged[k_].append(a[k_].copy())
        stats.append(st)
    out = {k_: np.concatenate(v) for k_, v in merged.items()}
    order = np.lexsort((out["lo"], out["hi"], out["bin"]))
    return {k_: v[order] for k_, v in out.items()}, stats, plans
