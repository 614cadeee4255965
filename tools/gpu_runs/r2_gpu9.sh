mkdir -p gpurun_out
export REPS=2
B200_FUSED_PERSIST=1 bash tools/capture_rep.sh r2_fused_persist_s2 corr2d_fused_persist 0 python tools/sweep_fused.py 2048 2
B200_FUSED_PERSIST=0 bash tools/capture_rep.sh r2_fused_final_s2 corr2d_fused_kernel 0 python tools/sweep_fused.py 2048 2
B200_FUSED_PERSIST=0 bash tools/capture_rep.sh r2_fused_final_s4 corr2d_fused_kernel 0 python tools/sweep_fused.py 1024 4
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
