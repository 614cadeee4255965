mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2_pytest_fused.log 2>&1; tail -3 gpurun_out/r2_pytest_fused.log | cut -c1-300
B200_FUSED_PERSIST=1 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -1
python tools/sweep_fused.py 2048 1 2 2.5 2>&1 | grep "two launches\|v0"
python tools/sweep_fused.py 1024 2 3 4 2>&1 | grep "two launches\|v0"
B200_FUSED_PERSIST=1 python tools/sweep_fused.py 2048 2 2>&1 | grep "two launches\|v0"
