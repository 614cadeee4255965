mkdir -p gpurun_out
python -m pytest tests/test_gpu_tma.py tests/test_gpu_fused.py -q -x > gpurun_out/r2_pytest_tma.log 2>&1; tail -15 gpurun_out/r2_pytest_tma.log | cut -c1-400
python tools/unaligned_scan.py > gpurun_out/r2_unaligned_scan.log 2>&1; cat gpurun_out/r2_unaligned_scan.log | cut -c1-200
