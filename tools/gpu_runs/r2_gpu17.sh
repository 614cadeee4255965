mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tma.py tests/test_gpu_parity.py -q -x -k "uniform or config3 or golden" > gpurun_out/r2_pytest_box2.log 2>&1; tail -3 gpurun_out/r2_pytest_box2.log | cut -c1-300
for sl in 0 1; do echo "== B200_BOX_SLIDE=$sl"; B200_BOX_SLIDE=$sl timeout 120 python tools/box_scan.py 2>&1 | grep "uint16.*axis 2\|uint8 box  *7 axis 2"; done > gpurun_out/r2_box_scan_contig.log 2>&1
cat gpurun_out/r2_box_scan_contig.log
for sl in 0 1; do B200_BOX_SLIDE=$sl timeout 120 python tools/bench_configs.py C3 2>&1 | grep config | cut -c1-200; done
