for cfg in "async_table_bytes=0" "async_table_bytes=0 table_budget_bytes=1e9" "async_table_bytes=0 table_budget_bytes=16e9" "async_table_bytes=100e6" "async_table_bytes=256e6" "async_table_bytes=1e9" "async_table_bytes=4e9" "async_table_bytes=16e9"; do
  args=""; for kv in $cfg; do args="$args --set $kv"; done
  echo "== $cfg"; timeout 200 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu $args 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['gpu_launches'])"
done
