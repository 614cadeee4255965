mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|Socket|NUMA|Model name"; free -g | head -2; } > gpurun_out/r2_topo8.log 2>&1
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py --gib 4 --reps 2 > gpurun_out/r2_pcie_n$n.json 2> gpurun_out/r2_pcie_n$n.err
  grep n_gpus gpurun_out/r2_pcie_n$n.json | cut -c1-330
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8_pre.json 2> gpurun_out/r2_bench_n8_pre.err
cut -c1-600 gpurun_out/r2_bench_n8_pre.json; grep -o '"e2e": {[^}]*}' gpurun_out/r2_bench_n8_pre.json | cut -c1-400
