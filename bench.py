#!/usr/bin/env python
"""bench.py — exact k-mer counting throughput of the B200 path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R]

A "step" is one whole counting job over one synthetic read set (SURVEY §8(d) generator):
  value : job throughput with the 2-bit packed reads already resident in HBM
          (fkm_count_packed_device), timed with CUDA events on the launching stream
  e2e   : the same job through the reference-facing call with HOST input
          (fkm_count_fasta: pinned FASTA text -> H2D -> device parse -> count -> stats D2H)
Workload at N=1 is BASELINE.json configs[1]: k=28 m=10 x=3 B=2048 useHT=1 on 50M x 150 bp reads.
With N>1 ranks every rank scans its own shard of the reads (weak scaling: 50M reads per GPU),
bins are owned per GPU and exchanged with one all-to-all (see fastkmer_b200/multigpu.py).

--impl reference times the CPU oracle (a port of the reference's algorithm; Spark cannot run
here) on a bounded sample of the same workload with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs (SURVEY §8(d) generator parameters).  Config 2 is the default: the one the metric is quoted on.
CONFIGS = {
    1: dict(k=28, m=10, x=3, B=2048, ht=0, kind="reads", reads=1_000_000, L=100, cov=20, seeds=(1001, 1002, 1003), seqtype=0,
            name="BASELINE configs[0]: k=28 m=10 x=3 B=2048 useHT=0, 1M synthetic 100 bp reads"),
    2: dict(k=28, m=10, x=3, B=2048, ht=1, kind="reads", reads=50_000_000, L=150, cov=30, seeds=(2001, 2002, 2003), seqtype=0,
            name="BASELINE configs[1]: k=28 m=10 x=3 B=2048 useHT=1, 50M synthetic 150 bp reads"),
    3: dict(k=31, m=11, x=3, B=4096, ht=0, kind="long", n_bases=1_200_000_000, seeds=(3001, 3002, 3003), seqtype=1,
            name="BASELINE configs[2]: long sequence, synthetic 1.2 Gbp genome (5% repeats, 0.5% N runs), k=31 m=11 B=4096 useHT=0"),
    4: dict(k=55, m=13, x=3, B=2048, ht=0, kind="reads", reads=50_000_000, L=150, cov=30, seeds=(4001, 4002, 4003), seqtype=0,
            name="BASELINE configs[3] shape at 50M reads per GPU (the 1B-read set is 125M per GPU on 8): k=55 m=13 x=3 B=2048, 128-bit k-mers, useHT=0"),
}


class Workload:
    def __init__(self, cid, reads_override=None, ht_override=None):
        self.c = dict(CONFIGS[cid])
        if reads_override and self.c["kind"] == "reads":
            self.c["reads"] = reads_override
        if ht_override is not None:
            self.c["ht"] = ht_override
        self.kind = self.c["kind"]

    def tc(self, fk):
        c = self.c
        return fk.TestConfiguration("", "", c["k"], c["m"], c["x"], max_b=c["B"], useHT=bool(c["ht"]), write=False, sequenceType=c["seqtype"])

    @property
    def scaling(self):
        return "weak" if self.kind == "reads" else "strong"

    def n_bases_total(self, world):
        return self.c["reads"] * self.c["L"] * world if self.kind == "reads" else self.c["n_bases"]

    def spec(self, rank, world):
        c = self.c
        if self.kind == "reads":
            total = c["reads"] * world
            return dict(seeds=c["seeds"], genome_len=max(c["L"] * 2, total * c["L"] // c["cov"]), n_reads=c["reads"], read_len=c["L"],
                        first_read=rank * c["reads"])
        per = c["n_bases"] // world                          # byte-range shard with a (k-1)-base halo
        first = rank * per
        n = (per if rank < world - 1 else c["n_bases"] - first) + (c["k"] - 1 if rank < world - 1 else 0)
        return dict(seeds=c["seeds"], first_pos=first, n_bases=n)

    def device_input(self, ctx, rank, world):
        sp = self.spec(rank, world)
        return ctx.synth_packed_device(sp) if self.kind == "reads" else ctx.synth_long_packed_device(sp)

    def host_fasta(self, api, fk, rank, world):
        sp = self.spec(rank, world)
        n = api.C.c_uint64()
        if self.kind == "reads":
            s = api._synth(sp)
            api._check(api.load_library().fkm_synth_fasta_host(api.C.byref(s), None, 0, api.C.byref(n)))
            return fk.synth_fasta(sp, out=api.host_alloc(n.value))
        s = api._synth_long(sp)
        api._check(api.load_library().fkm_synth_long_fasta_host(api.C.byref(s), None, 0, api.C.byref(n)))
        return fk.synth_long_fasta(sp, out=api.host_alloc(n.value))

    def cpu_sample(self, oracle, threads):
        """bounded sample of the same workload for the CPU oracle -> (fasta uint8 array, description).  The text comes from the
        oracle's own restatement of the generator: the CPU arm never loads the product library."""
        c = self.c
        if self.kind == "reads":
            reads = min(c["reads"], 5_000_000)            # ~10 s of work for 16 host threads
            sp = dict(seeds=c["seeds"], genome_len=reads * c["L"] // c["cov"], n_reads=reads, read_len=c["L"])
            return oracle.synth_fasta(sp, threads=threads), "%d reads x %d bp of the same generator" % (reads, c["L"])
        n = min(c["n_bases"], 100_000_000)
        return oracle.synth_long_fasta(dict(seeds=c["seeds"], n_bases=n), threads=threads), "first %d bp of the same synthetic genome" % n


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_cpu_oracle(threads, wl):
    """Times the CPU oracle (port of the reference algorithm) on a bounded sample of the workload.  -> dict"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oracle = oracle_lib.load()
    fasta, desc = wl.cpu_sample(oracle, threads)
    c = wl.c
    t0 = time.perf_counter()
    res = oracle.count(fasta, c["k"], c["m"], c["x"], c["B"], c["ht"], threads=threads, sorted_=False)
    dt = time.perf_counter() - t0
    st = res["stats"]
    return {"seconds": dt, "n_bases": st["n_bases"], "n_kmers": st["n_kmers"], "n_distinct": st["n_distinct"],
            "n_superkmers_ref": st["n_superkmers"], "superkmer_bases_ref": st["superkmer_bases"], "desc": desc}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.config, args.reads, args.ht)
    threads = os.cpu_count() or 1
    run_cpu_oracle(threads, wl)                       # one warm-up pass (page cache, library load)
    t = []
    last = None
    for _ in range(args.steps):
        last = run_cpu_oracle(threads, wl)
        t.append(last["seconds"])
    sec = sum(t) / len(t)
    value = last["n_bases"] / sec
    sample = last["desc"] + ", oracle port of the reference algorithm (Spark cannot run here)"
    line = {"impl": "reference", "metric": "bases_per_sec", "value": value, "unit": "bases/s", "kmers_per_sec": last["n_kmers"] / sec,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": wl.scaling, "vs_baseline": None, "dtype": "u64" if wl.c["k"] <= 32 else "u128", "data": "synthetic",
            "config": {"workload": wl.c["name"] + " — bounded sample: " + last["desc"]},
            "cpu_baseline": {"value": value, "unit": "bases/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (1-based); default 2 = configs[1]")
    ap.add_argument("--reads", type=int, default=None, help="reads per GPU (override, read configs only)")
    ap.add_argument("--ht", type=int, default=None, help="override useHT")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--set", action="append", default=[], help="context knob name=value (tuning experiments)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import fastkmer_b200 as fk
    from fastkmer_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream()                        # a real (non-NULL) stream: kernels and timing events share it
    torch.cuda.set_stream(stream)
    ctx = fk.Context(local_rank, stream.cuda_stream)
    for kv in args.set:
        name, val = kv.split("=")
        ctx.set(name, float(val))
    wl = Workload(args.config, args.reads, args.ht)
    cfg = wl.tc(fk)

    if world > 1:
        from fastkmer_b200 import multigpu
        job = multigpu.ShardedJob(ctx, cfg, dist, rank, world)
    else:
        job = None

    # ---------------- device-resident input ----------------
    d_b, d_i, n_pos = wl.device_input(ctx, rank, world)

    def step_resident():
        if job is None:
            return ctx.count_packed_device(cfg, d_b, d_i, n_pos, want_result=False)[1]
        return job.count_packed_device(d_b, d_i, n_pos)

    launches0 = None
    st = None
    for _ in range(args.warmup):
        st = step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = api.load_library().fkm_total_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = [0.0] * 8
    ms_fold = 0.0
    ms_part = 0.0
    e0.record(stream)
    for _ in range(args.steps):
        st = step_resident()
        for i in range(8):
            stage[i] += st["ms_stage"][i]
        ms_fold += st["ms_fold"]
        ms_part += st.get("ms_partition", 0.0)
    e1.record(stream)
    barrier()
    launches = api.load_library().fkm_total_launches() - launches0
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    stage = [s / args.steps for s in stage]
    ms_fold /= args.steps
    ms_part /= args.steps
    n_bases_total = wl.n_bases_total(world)
    n_kmers_total = st["n_kmers_global"] if "n_kmers_global" in st else st["n_kmers"]
    n_distinct_total = st["n_distinct_global"] if "n_distinct_global" in st else st["n_distinct"]
    value = n_bases_total / (ms_step * 1e-3)

    # ---------------- parity of this very path (same ranks, same exchange) against the CPU oracle on a bounded set ----------------
    parity = None
    if not args.no_cpu:
        small = 100_000 if wl.kind == "reads" else 3_000_000
        c = wl.c
        if wl.kind == "reads":
            sp_all = dict(seeds=c["seeds"], genome_len=max(2 * c["L"], small * world * c["L"] // c["cov"]), n_reads=small * world, read_len=c["L"])
            sp_me = dict(sp_all, n_reads=small, first_read=rank * small)
            sb, si, sn = ctx.synth_packed_device(sp_me)
        else:
            per = small // world
            first = rank * per
            n = (per if rank < world - 1 else small - first) + (c["k"] - 1 if rank < world - 1 else 0)
            sp_all = dict(seeds=c["seeds"], n_bases=small)
            sb, si, sn = ctx.synth_long_packed_device(dict(seeds=c["seeds"], first_pos=first, n_bases=n))
        sst = ctx.count_packed_device(cfg, sb, si, sn, want_result=False)[1] if job is None else job.count_packed_device(sb, si, sn)
        ctx.free_device(sb)
        ctx.free_device(si)
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib
            oracle = oracle_lib.load()
            text = oracle.synth_fasta(sp_all) if wl.kind == "reads" else oracle.synth_long_fasta(sp_all)
            ws = oracle.count(text, c["k"], c["m"], c["x"], c["B"], c["ht"], threads=os.cpu_count() or 1, sorted_=False)["stats"]
            g = (lambda key: sst[key + "_global"] if key + "_global" in sst else sst[key])
            got = (g("n_kmers"), g("n_distinct"), g("digest_sum"), g("digest_xor"))
            want = (ws["n_kmers"], ws["n_distinct"], ws["digest_sum"], ws["digest_xor"])
            assert got == want, "GPU result differs from the CPU oracle on the bounded set: %r vs %r" % (got, want)
            parity = {"oracle_checked": True, "what": "%s, all %d rank(s), digest of every (bin, k-mer, count)" %
                      ("%d reads" % (small * world) if wl.kind == "reads" else "%d bp" % small, world), "n_kmers": int(ws["n_kmers"])}

    # ---------------- end to end from host FASTA ----------------
    e2e = None
    if not args.no_e2e:
        fasta = wl.host_fasta(api, fk, rank, world)

        def step_e2e():
            if job is None:
                return ctx.count_fasta(cfg, fasta, want_result=False)[1]
            return job.count_fasta(fasta)

        for _ in range(max(1, args.warmup - 1)):
            se = step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            se = step_e2e()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert se["digest_sum"] == st["digest_sum"] and se["n_kmers"] == st["n_kmers"], "host path and resident path disagree"
        if dist is not None:
            assert se["digest_sum_global"] == st["digest_sum_global"] and se["total_count_global"] == se["n_kmers_global"]
        e2e = {"value": n_bases_total / (dt / args.steps), "unit": "bases/s", "h2d_bytes_per_step": int(se["h2d_bytes"]),
               "d2h_bytes_per_step": int(se["d2h_bytes"]), "ms_per_step": dt / args.steps * 1e3,
               "ms_h2d_and_parse": se["ms_stage"][0]}

    if rank != 0:
        return

    # ---------------- CPU baseline (oracle port) on a bounded sample ----------------
    cpu = None
    ls_per_kmer = {28: 3.73, 31: 3.77, 55: 3.48}.get(wl.c["k"], 3.7)     # SURVEY §8(d) planning ratios (reference cutting rule)
    if not args.no_cpu and world == 1:
        threads = os.cpu_count() or 1
        r = run_cpu_oracle(threads, wl)
        ls_per_kmer = r["superkmer_bases_ref"] / max(1, r["n_kmers"])
        cpu = {"value": r["n_bases"] / r["seconds"], "unit": "bases/s", "cores": threads, "kind": "port",
               "kmers_per_sec": r["n_kmers"] / r["seconds"], "sample": "%s, %.1f s" % (r["desc"], r["seconds"])}

    # ---------------- roofline ----------------
    peak, peak_src = measured_peaks()
    # algorithmic bytes of the whole pipeline, SURVEY §8(d): N_b/4 + 2*L_s/4 + D*(W+4), L_s under the reference's cutting rule
    L_s = ls_per_kmer * n_kmers_total
    key_bytes = 8 if wl.c["k"] <= 32 else 16
    bytes_alg = n_bases_total / 4 + 2 * L_s / 4 + n_distinct_total * (key_bytes + 4)
    # dominant kernel: the count stage (stage 3; k_count_ht for useHT=1, k_expand + radix passes for useHT=0).  It reads
    # the super-k-mer stream once (L_s/4 algorithmic bytes); the distinct (k-mer,count) pairs leave through stage 4.
    n_count_launches = max(1, int(st["n_batches"]))
    smem_path = bool(st["n_mid_bins"])
    part_path = smem_path and ms_part > 0.0
    # dominant kernel = the count stage's own kernel (stage 3), timed with CUDA events inside the library.  Algorithmic bytes per
    # the count stage reads the super-k-mer stream once (L_s/4); the shared-memory kernels (k_count_keys, k_count_smem)
    # also write the result (D*(W+4)) themselves: they have no compaction kernel.  `traffic` is what DRAM really moved per launch.
    count_bytes = (L_s / 4 + (n_distinct_total * (key_bytes + 4) if smem_path else 0)) / world
    count_ms = stage[3]
    kernel = "k_count_keys" if part_path else "k_count_smem" if smem_path else ("k_count_ht" if wl.c["ht"] else "k_expand+k_radix_*")
    traffic, traffic_src = None, None
    try:                                              # DRAM bytes per launch from an ncu --set full capture of the same kernel (profiles/)
        tname = "r2_count_keys_traffic.json" if part_path else "r2_count_smem_traffic.json" if smem_path else "r1f_count_ht_traffic.json"
        tj = json.load(open(os.path.join(ROOT, "profiles", tname)))
        if smem_path and wl.c["k"] <= 32:
            traffic = tj["dram_bytes_per_kmer"] * (n_kmers_total / world / n_count_launches)
            traffic_src = "static: ncu capture profiles/%s scaled by k-mers per launch" % tname
        elif wl.c["ht"] and wl.c["k"] <= 32 and st["n_folded_records"]:
            traffic = tj["dram_bytes_per_record"] * (st["n_folded_records"] / n_count_launches)
            traffic_src = "static: ncu capture profiles/r1f_count_ht_traffic.json (round 1) scaled by folded records per launch"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": count_bytes / (count_ms * 1e-3) / 1e9 if count_ms else None,
                "peak": peak, "unit": "GB/s", "frac": (count_bytes / (count_ms * 1e-3) / 1e9 / peak) if count_ms else None,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": count_bytes / n_count_launches, "peak_source": peak_src,
                "launches_per_step": n_count_launches, "avg_launch_ms": count_ms / n_count_launches,
                "pipeline": {"bytes_alg": bytes_alg, "achieved": bytes_alg / (ms_step * 1e-3) / 1e9 / world,
                             "frac": bytes_alg / (ms_step * 1e-3) / 1e9 / world / peak}}
    shuffle = None
    if world > 1:
        # the bin exchange: bytes this rank sent to its peers and the time of the all-to-all (rank 0's view)
        sent, xms = st.get("exchange_bytes_sent", 0), st.get("exchange_ms", 0.0)
        gbs = sent / (xms * 1e-3) / 1e9 if xms else None
        shuffle = {"bytes_sent_per_gpu": int(sent), "ms": xms, "GBps_per_gpu": gbs, "nvlink_peak_GBps": 770.0,
                   "frac": gbs / 770.0 if gbs else None, "peak_source": "measured peer copy, B200_PROFILING.md"}
    line = {"metric": "bases_per_sec", "value": value, "unit": "bases/s", "kmers_per_sec": n_kmers_total / (ms_step * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": wl.scaling, "vs_baseline": None, "dtype": "u64" if wl.c["k"] <= 32 else "u128", "data": "synthetic",
            "config": {"workload": wl.c["name"] + (" (per GPU; 1% substitutions, 0.1% N)" if wl.kind == "reads" else ""),
                       "reads_per_gpu": wl.c.get("reads"), "n_bases": n_bases_total,
                       "n_kmers": int(n_kmers_total), "n_distinct": int(n_distinct_total),
                       "n_superkmers": int(st["n_superkmers"]), "n_folded_records": int(st["n_folded_records"]),
                       "count_tables": ("shared memory (%d %s, %d on the slow path)" % (st["n_mid_bins"], "hash-partitioned sub-buckets" if part_path else "mid bins", st["n_slow_bins"])) if smem_path else "global memory",
                       "l2": "inputs (%.1f GB packed) larger than L2, no flush" % (n_pos * 3 / 8 / 1e9)},
            "stage_ms": {"histogram": stage[1], "scatter": stage[2], "fold": ms_fold, "partition": ms_part, "count": stage[3], "compact": stage[4], "digest": stage[5],
                         "device_pipeline": stage[7]},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "shuffle": shuffle, "parity": parity}
    print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
