mkdir -p gpurun_out
timeout -s KILL 60 python tools/box_min.py > gpurun_out/r2_box_min.log 2>&1; tail -4 gpurun_out/r2_box_min.log | cut -c1-200
timeout -s KILL 400 python -m pytest tests/test_gpu_fused.py -q -x -k "box" > gpurun_out/r2_pytest_box.log 2>&1
rc=$?
tail -15 gpurun_out/r2_pytest_box.log | cut -c1-600
if [ $rc -eq 0 ]; then
  timeout -s KILL 300 python tools/box_pair_scan.py 512 > gpurun_out/r2_box_pair_scan.log 2>&1; cat gpurun_out/r2_box_pair_scan.log | cut -c1-250
fi
