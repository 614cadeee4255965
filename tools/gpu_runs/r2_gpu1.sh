# round 2, first GPU call: probes, fused-pair parity + geometry sweep, PCIe ceiling, full GPU suite, bench
mkdir -p gpurun_out
{ python -c "import dask; print('dask', dask.__version__)"; python -c "import dask_image; print('dask_image ok')"; \
  python -c "import cupy; print('cupy ok')"; } > gpurun_out/r2_dask_probe.log 2>&1
{ nvidia-smi topo -m; lscpu | head -40; numactl -H; nproc; python -c "import os; print(sorted(os.sched_getaffinity(0)))"; \
  cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective; free -g; } > gpurun_out/r2_topo.log 2>&1
python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2_pytest_fused.log 2>&1; tail -3 gpurun_out/r2_pytest_fused.log | cut -c1-300
python tools/sweep_fused.py 1024 2 3 4 > gpurun_out/r2_sweep_fused_1024.log 2>&1; cat gpurun_out/r2_sweep_fused_1024.log | cut -c1-200
python tools/sweep_fused.py 2048 2 > gpurun_out/r2_sweep_fused_2048.log 2>&1; cat gpurun_out/r2_sweep_fused_2048.log | cut -c1-200
python tools/pcie_ceiling.py --bind 0 > gpurun_out/r2_pcie_n1_nobind.json 2>&1; cut -c1-400 gpurun_out/r2_pcie_n1_nobind.json
python tools/pcie_ceiling.py --bind 1 > gpurun_out/r2_pcie_n1_bind.json 2>&1; cut -c1-400 gpurun_out/r2_pcie_n1_bind.json
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_all.log 2>&1; tail -3 gpurun_out/r2_pytest_all.log | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; cut -c1-1500 gpurun_out/r2_bench1.json; tail -3 gpurun_out/r2_bench1.err
