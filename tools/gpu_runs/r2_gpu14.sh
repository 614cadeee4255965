mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_sharded.py -q -x -k "nd_halo_in_place" > gpurun_out/r2_pytest_ndhalo.log 2>&1; tail -15 gpurun_out/r2_pytest_ndhalo.log | cut -c1-600
