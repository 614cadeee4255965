mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; tail -4 gpurun_out/r2_pytest_all.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2_pytest_all.log | head
