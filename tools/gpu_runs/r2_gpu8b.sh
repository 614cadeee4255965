mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -x > gpurun_out/r2_pytest_sharded_n8.log 2>&1; tail -4 gpurun_out/r2_pytest_sharded_n8.log | cut -c1-400
for n in 8 4 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$n.json').read().strip().splitlines()[-1])
print('N=$n value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'checksum',d['verify']['checksum'],d['verify']['ok'],'e2e',d['e2e'].get('value'),d['e2e'].get('roofline',{}).get('frac'),d['e2e'].get('roofline',{}).get('peak_gbs'))
for c in d['configs']: print('  ',c.get('name'), c.get('ms'), c.get('frac_of_measured_peak'), c.get('checksum'), c.get('kernel'), c.get('error'))
PY
done
