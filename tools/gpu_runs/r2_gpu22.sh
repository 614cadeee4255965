mkdir -p gpurun_out
timeout -s KILL 120 python tools/box_pair_one.py > gpurun_out/r2_box2d_plain.log 2>&1 && \
timeout -s KILL 400 bash tools/capture_rep.sh r2_box2d_u16 corr2d_fused_kernel 2 python tools/box_pair_one.py
tail -3 gpurun_out/r2_box2d_u16_ncu.log | cut -c1-200
grep -E "Duration|Issue Slots Busy|Executed Ipc Active|Registers Per|Achieved Occupancy|Warp Cycles Per Issued|Theoretical Occupancy|Block Limit" gpurun_out/r2_box2d_u16_details.txt | cut -c1-160
