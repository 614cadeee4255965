mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -q -x > gpurun_out/r2_pytest_fused.log 2>&1; tail -2 gpurun_out/r2_pytest_fused.log | cut -c1-300
timeout 200 python tools/bench_configs.py C4a C4b C2 2>&1 | grep config | cut -c1-200
