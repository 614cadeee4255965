mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2_pytest_fused.log 2>&1; tail -3 gpurun_out/r2_pytest_fused.log | cut -c1-300
python tools/sweep_fused.py 2048 2 > gpurun_out/r2_sweep_fused_2048b.log 2>&1; cat gpurun_out/r2_sweep_fused_2048b.log | cut -c1-200
python tools/sweep_fused.py 1024 2 3 4 > gpurun_out/r2_sweep_fused_1024b.log 2>&1; cat gpurun_out/r2_sweep_fused_1024b.log | cut -c1-200
export REPS=2
bash tools/capture_rep.sh r2_fused_b_v0_s2 corr2d_fused 0 python tools/sweep_fused.py 2048 2
bash tools/capture_rep.sh r2_fused_b_v0_s4 corr2d_fused 0 python tools/sweep_fused.py 1024 4
