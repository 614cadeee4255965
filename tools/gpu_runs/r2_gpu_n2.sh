mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/test_gpu_sharded.py -q -x > gpurun_out/r2_pytest_sharded_n2.log 2>&1; tail -3 gpurun_out/r2_pytest_sharded_n2.log | cut -c1-500
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench_n2_boxpair.json 2> gpurun_out/r2_bench_n2_boxpair.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n2_boxpair.json').read().strip().splitlines()[-1])
print('N=2 value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'checksum',d['verify']['checksum'],d['verify']['ok'])
for c in d['configs']: print('  ',c.get('name'), c.get('ms'), c.get('frac_of_measured_peak'), c.get('launches'), c.get('checksum'), c.get('kernel'), c.get('error'))
PY
