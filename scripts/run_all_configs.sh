#!/bin/bash
# One B200: every BASELINE config (and the other reduce path of configs 3 and 4) through bench.py, one JSON line per run into
# profiles/<tag>_config*.json.  Usage: scripts/run_all_configs.sh <tag>   (run on the GPU box, e.g. under gpurun)
tag=${1:-r2}
out=${2:-gpurun_out}
mkdir -p "$out"
python bench.py --config 2 --steps 5 --warmup 3 > "$out/${tag}_config2.json" 2> "$out/${tag}_config2.err"
python bench.py --config 1 --steps 5 --warmup 3 > "$out/${tag}_config1.json" 2> "$out/${tag}_config1.err"
python bench.py --config 3 --steps 3 --warmup 3 > "$out/${tag}_config3.json" 2> "$out/${tag}_config3.err"
python bench.py --config 3 --ht 1 --steps 3 --warmup 3 --no-cpu > "$out/${tag}_config3_ht1.json" 2> "$out/${tag}_config3_ht1.err"
python bench.py --config 4 --steps 2 --warmup 3 > "$out/${tag}_config4shape.json" 2> "$out/${tag}_config4shape.err"
python bench.py --config 4 --ht 1 --steps 3 --warmup 3 --no-cpu > "$out/${tag}_config4shape_ht1.json" 2> "$out/${tag}_config4shape_ht1.err"
for f in "$out/${tag}"_config*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "%.1f ms" % d["ms_per_step"], "%.2f Gbases/s" % (d["value"] / 1e9), "e2e %.2f" % (d["e2e"]["value"] / 1e9) if d.get("e2e") else "", d["stage_ms"])
PY
done
