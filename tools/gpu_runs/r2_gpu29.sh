mkdir -p gpurun_out
# random parity hunt on volumes that cross tile boundaries (tools/gpu_hunt.py), time-boxed
timeout -s KILL 60 python tools/gpu_hunt.py 24 1 > gpurun_out/r2_gpu_hunt_seed1.log 2>&1
tail -12 gpurun_out/r2_gpu_hunt_seed1.log | cut -c1-600
