mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; tail -6 gpurun_out/r2_pytest_all.log | cut -c1-300
python bench.py --steps 5 --no-cpu --no-e2e > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; tail -3 gpurun_out/r2_bench4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['verify']['checksum'], d['verify']['ok'])
for c in d['configs']: print(c.get('name'), c.get('ms'), c.get('frac_of_measured_peak'), c.get('checksum'), c.get('error'))
PY
cat gpurun_out/parity_errors.json | head -50
