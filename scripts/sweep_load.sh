for lf in 0.6 0.75 0.85 0.95; do
  echo "== load_factor=$lf"; timeout 200 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --set load_factor=$lf 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()})"
done
