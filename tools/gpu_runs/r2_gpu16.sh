mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -q -x -k "out_of_core or host_volume" > gpurun_out/r2_pytest_ooc.log 2>&1; tail -15 gpurun_out/r2_pytest_ooc.log | cut -c1-500
