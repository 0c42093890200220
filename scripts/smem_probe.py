"""One job of configs[1] (or --reads N) through the shared-memory count path with a few knob settings; prints the stage times."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastkmer_b200 as fk

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=50_000_000)
ap.add_argument("--k", type=int, default=28)
ap.add_argument("--m", type=int, default=10)
ap.add_argument("--B", type=int, default=2048)
ap.add_argument("--set", action="append", default=[])
ap.add_argument("--sweep", default="")          # e.g. smem_fill=0.3,0.45,0.6
args = ap.parse_args()
ctx = fk.Context(0)
spec = dict(seeds=(2001, 2002, 2003), genome_len=args.reads * 150 // 30, n_reads=args.reads, read_len=150)
d_b, d_i, n_pos = ctx.synth_packed_device(spec)
cfg = fk.TestConfiguration("", "", args.k, args.m, 3, max_b=args.B, useHT=True, write=False)
for kv in args.set:
    n, v = kv.split("="); ctx.set(n, float(v))
name, vals = (args.sweep.split("=") + [""])[:2] if args.sweep else ("", "")
for v in (vals.split(",") if vals else [None]):
    if v is not None:
        ctx.set(name, float(v))
    for rep in range(2):
        _, st = ctx.count_packed_device(cfg, d_b, d_i, n_pos, want_result=False)
    print(json.dumps({"knob": name, "value": v, "ms": [round(x, 2) for x in st["ms_stage"]], "ms_partition": round(st["ms_partition"], 2), "ms_fold": round(st["ms_fold"], 2), "n_mid": st["n_mid_bins"], "n_slow": st["n_slow_bins"],
                      "fallbacks": st["n_fallbacks"], "records": st["n_superkmers"], "distinct": st["n_distinct"], "kmers": st["n_kmers"]}), flush=True)
ctx.close()
