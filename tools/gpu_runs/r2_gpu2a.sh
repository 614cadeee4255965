mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharded.py -q -x > gpurun_out/r2_pytest_sharded_n2.log 2>&1; tail -12 gpurun_out/r2_pytest_sharded_n2.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['verify'], d['e2e'].get('value'), d['e2e'].get('roofline'))
for c in d['configs']: print(c.get('name'), c.get('ms'), c.get('checksum'), c.get('error'))
PY
