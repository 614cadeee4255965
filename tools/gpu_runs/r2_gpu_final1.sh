mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; tail -2 gpurun_out/r2_pytest_all.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; cut -c1-400 gpurun_out/r2_bench_reference.json
export REPS=2
bash tools/capture_rep.sh r2_fused_final_s2 corr2d_fused_kernel 0 python tools/sweep_fused.py 2048 2
bash tools/capture_rep.sh r2_fused_final_s4 corr2d_fused_kernel 0 python tools/sweep_fused.py 1024 4
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2_bench_plain_for_ncu.json 2> gpurun_out/r2_bench_plain_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:corr1d|corr2d|uniform1d|corrnd|combine|generic' -c 100 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2_ncu_bench.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'launches',d['gpu_launches'])
r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','moved_gb_s','moved_frac','launch_ms','kernel_share_of_step')})
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['roofline']['frac'],d['e2e']['roofline']['peak_gbs'])
print('cpu',d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('verify',d['verify']['checksum'],d['verify']['subvolume_max_rel_err'], d['verify']['ok'])
print('clocks',d['clocks'])
for c in d['configs']: print(c.get('name'), round(c.get('ms',0),3), round(c.get('frac_of_measured_peak',0),3), c.get('launches'), c.get('kernel'), c.get('checksum'), c.get('error'))
PY
