# One gpurun call for the exact tiled integer kernels: parity, per-pass timings and
# ncu raw pages (exported as CSV on the box: a .ncu-rep is too large to bring back).
# usage (from the repo root, under gpurun): bash tools/int_gpu.sh <tag>
tag=${1:-r2}
python -m pytest tests/test_gpu_int.py -q > gpurun_out/pytest_int_$tag.log 2>&1; tail -1 gpurun_out/pytest_int_$tag.log | cut -c1-160
(cd tools && python int_scan.py 1 2 4 > ../gpurun_out/int_scan_$tag.log 2>&1; python int_nd_scan.py > ../gpurun_out/int_nd_scan_$tag.log 2>&1)
grep "sigma 2" gpurun_out/int_scan_$tag.log | cut -c1-120
grep -v generic gpurun_out/int_nd_scan_$tag.log | cut -c1-120
bash tools/capture_raw.sh prof_${tag}_int_strided_u16 corr1d_int_strided env N=512 DTYPES=u2 python tools/int_scan.py 2
bash tools/capture_raw.sh prof_${tag}_int_contig_u16 corr1d_int_contig env N=512 DTYPES=u2 python tools/int_scan.py 2
bash tools/capture_raw.sh prof_${tag}_corrnd_int_u16 corrnd_int env DTYPES=u2 python tools/int_nd_scan.py
ls -la gpurun_out/prof_${tag}_*int*_raw.csv
# the eight-outputs-per-lane experiment: rebuild with the flag, rerun the parity tests and the scan with it selected
# (uncomment for an A/B run; the default build does not contain these kernels)
# NVCC_EXTRA="-DB200_INT_NQ8 -DB200_INT_FIXED" python -m dask_image_b200.build --force > /dev/null 2>&1
# B200_INT_NQ=8 python -m pytest tests/test_gpu_int.py -q 2>&1 | tail -1
# (cd tools && B200_INT_NQ=8 DTYPES=u2 python int_scan.py 2 | cut -c1-120; B200_INT_FIXED=1 B200_INT_NQ=8 DTYPES=u2 python int_scan.py 2 | cut -c1-120)
# B200_INT_FIXED=1 python -m pytest tests/test_gpu_int.py -q 2>&1 | tail -1
# (cd tools && B200_INT_FIXED=1 DTYPES=u2 python int_scan.py 2 | cut -c1-120; B200_INT_NQ=8 DTYPES=u2 python int_scan.py 2 | cut -c1-120; B200_INT_FIXED=1 B200_INT_NQ=8 DTYPES=u2 python int_scan.py 2 | cut -c1-120)
