mkdir -p gpurun_out
# bench.py on the final tree: headline + verify + the cpu_baseline leg through the reference's own API (no e2e, no configs)
timeout -s KILL 110 python bench.py --steps 3 --warmup 3 --no-e2e --no-configs > gpurun_out/r2_bench_final_tree.json 2> gpurun_out/r2_bench_final_tree.err
echo rc=$?
tail -c 1500 gpurun_out/r2_bench_final_tree.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_final_tree.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["verify"], d["cpu_baseline"], d["roofline"]["frac"])
PY
