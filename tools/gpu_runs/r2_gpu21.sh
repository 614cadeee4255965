mkdir -p gpurun_out
export B200_NDFILTERS_LIB=$PWD/gpurun_dbg/libb200_boxdbg.so
: > gpurun_out/r2_box_stages.log
for k in 1 2 3 4 5 6 0; do
  echo "== stage $k (running sums)" >> gpurun_out/r2_box_stages.log
  B200_BOX_DBG=$k timeout -s KILL 100 python tools/box_min.py 2>&1 | grep -v "^  \|^Traceback\|^Search\|^CUDA kernel\|^For debugging\|^Compile with" | cut -c1-200 >> gpurun_out/r2_box_stages.log
done
for k in 3 0; do
  echo "== stage $k (direct form)" >> gpurun_out/r2_box_stages.log
  B200_BOX2D_SLIDE=0 B200_BOX_DBG=$k timeout -s KILL 100 python tools/box_min.py 2>&1 | grep -v "^  \|^Traceback\|^Search\|^CUDA kernel\|^For debugging\|^Compile with" | cut -c1-200 >> gpurun_out/r2_box_stages.log
done
cat gpurun_out/r2_box_stages.log
