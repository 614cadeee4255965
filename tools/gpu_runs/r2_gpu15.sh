mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tma.py -q -x -k "uniform" > gpurun_out/r2_pytest_box.log 2>&1; tail -3 gpurun_out/r2_pytest_box.log | cut -c1-300
echo "== direct form (previous build)" > gpurun_out/r2_box_scan.log
B200_NDFILTERS_LIB=/root/repo/dask_image_b200/libb200ndfilters_old.so timeout 120 python tools/box_scan.py 2>&1 | grep "uint16.*axis [01]" >> gpurun_out/r2_box_scan.log
echo "== sliding window" >> gpurun_out/r2_box_scan.log
timeout 120 python tools/box_scan.py 2>&1 | grep "axis [01]" >> gpurun_out/r2_box_scan.log
cat gpurun_out/r2_box_scan.log
