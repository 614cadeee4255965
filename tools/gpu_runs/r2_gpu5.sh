mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; tail -8 gpurun_out/r2_pytest_all.log | cut -c1-300
python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; cut -c1-3000 gpurun_out/r2_bench3.json; tail -5 gpurun_out/r2_bench3.err
