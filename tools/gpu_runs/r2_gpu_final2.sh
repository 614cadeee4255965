mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log | cut -c1-300
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; tail -2 gpurun_out/r2_pytest_all.log | cut -c1-300
timeout -s KILL 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err | cut -c1-300
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
