mkdir -p gpurun_out
# variant 3 (32 rows, 8 warps, alias, 4 CTAs/SM) at sigma 2 on (512, 2048, 2048); launch order in sweep_fused:
# per sigma: 2x(two launches)x(reps+1) then per variant (reps+1) pair launches -> first fused launch of variant 3 = skip 3*(REPS+1)
export REPS=2
bash tools/capture_rep.sh r2_fused_v3_s2 corr2d_fused 9 python tools/sweep_fused.py 2048 2
bash tools/capture_rep.sh r2_fused_v7_s2 corr2d_fused 21 python tools/sweep_fused.py 2048 2
bash tools/capture_rep.sh r2_fused_v2_s4 corr2d_fused 6 python tools/sweep_fused.py 1024 4
ls -la gpurun_out/ | grep r2_fused
