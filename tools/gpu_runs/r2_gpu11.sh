mkdir -p gpurun_out
for rep in 1 2; do
for lib in old new; do
  if [ $lib = old ]; then export B200_NDFILTERS_LIB=/root/repo/dask_image_b200/libb200ndfilters_old.so; else unset B200_NDFILTERS_LIB; fi
  echo "== $lib (rep $rep)"
  REPS=20 python tools/sweep_fused.py 2048 1 2 2.5 2>&1 | grep "v0"
  REPS=20 python tools/sweep_fused.py 1024 3 4 2>&1 | grep "v0"
done
done > gpurun_out/r2_phaseb_ab.log 2>&1
cat gpurun_out/r2_phaseb_ab.log | cut -c1-150
