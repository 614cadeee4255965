mkdir -p gpurun_out
B200_FUSED_PERSIST=1 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2_pytest_fused_persist.log 2>&1; tail -3 gpurun_out/r2_pytest_fused_persist.log | cut -c1-300
for pz in 0 1; do
  echo "== B200_FUSED_PERSIST=$pz"
  B200_FUSED_PERSIST=$pz python tools/sweep_fused.py 2048 1 2 2.5 2>&1 | grep "two launches\|v0"
  B200_FUSED_PERSIST=$pz python tools/sweep_fused.py 1024 2 2>&1 | grep "two launches\|v0"
done > gpurun_out/r2_persist_ab.log 2>&1
cat gpurun_out/r2_persist_ab.log | cut -c1-220
