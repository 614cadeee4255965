mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sharded.py -q -x -k "nccl" > gpurun_out/r2_pytest_sharded_n2.log 2>&1; tail -12 gpurun_out/r2_pytest_sharded_n2.log | cut -c1-900
for ph in 1 0; do
B200_PEER_HALO=$ph timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$ph tools/bench_configs.py C5 2>&1 | grep config | cut -c1-220
done
