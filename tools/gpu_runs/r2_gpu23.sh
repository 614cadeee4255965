mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_fused.py -q -x > gpurun_out/r2_pytest_fused.log 2>&1; tail -6 gpurun_out/r2_pytest_fused.log | cut -c1-500
timeout -s KILL 120 python tools/box_pair_one.py > gpurun_out/r2_box2d_plain.log 2>&1 && \
timeout -s KILL 400 bash tools/capture_rep.sh r2_box2d_u16 corr2d_fused_kernel 2 python tools/box_pair_one.py
grep -E "Duration|Issue Slots Busy|Registers Per|Achieved Occupancy|L1/TEX Cache Throughput|DRAM Throughput|Mem Busy|bank conflicts|excessive" gpurun_out/r2_box2d_u16_details.txt | cut -c1-200
