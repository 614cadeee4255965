mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_fused.py -q -x > gpurun_out/r2_pytest_fused.log 2>&1; tail -2 gpurun_out/r2_pytest_fused.log | cut -c1-300
timeout -s KILL 300 python tools/bench_configs.py C2 C4a C4b H4 > gpurun_out/r2_configs_minb4.log 2>&1; grep config gpurun_out/r2_configs_minb4.log | cut -c1-220
