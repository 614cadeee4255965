mkdir -p gpurun_out
export B200_NDFILTERS_LIB=$PWD/gpurun_dbg/libb200_sweep.so
REPS=10 SWEEP_VARIANTS=0,3 timeout -s KILL 300 python tools/sweep_fused.py 1024 3 4 > gpurun_out/r2_sweep_fused_minb_1024.log 2>&1
cat gpurun_out/r2_sweep_fused_minb_1024.log | cut -c1-220
REPS=10 SWEEP_VARIANTS=0,3 timeout -s KILL 300 python tools/sweep_fused.py 2048 3 4 > gpurun_out/r2_sweep_fused_minb_2048.log 2>&1
cat gpurun_out/r2_sweep_fused_minb_2048.log | cut -c1-220
