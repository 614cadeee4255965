mkdir -p gpurun_out
python -m pytest tests/test_gpu_tma.py -q -x -k "long_filter or tma_correlate1d" > gpurun_out/r2_pytest_ss.log 2>&1; tail -3 gpurun_out/r2_pytest_ss.log | cut -c1-300
for st in 0 1; do echo "== B200_STRIDED_STATIC=$st"; B200_STRIDED_STATIC=$st python tools/unaligned_scan.py 2>&1 | grep "1024, 1024" | grep "sigma 4 axis [01]"; done
for st in 0 1; do echo "== B200_STRIDED_STATIC=$st"; B200_STRIDED_STATIC=$st python tools/bench_configs.py C2 C4a H4 2>&1 | cut -c1-200; done
