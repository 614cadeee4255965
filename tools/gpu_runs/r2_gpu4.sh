mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_all.log 2>&1; tail -3 gpurun_out/r2_pytest_all.log | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; cut -c1-1200 gpurun_out/r2_bench2.json; tail -3 gpurun_out/r2_bench2.err
python tools/bench_configs.py > gpurun_out/r2_configs_n1.log 2>&1; cat gpurun_out/r2_configs_n1.log | cut -c1-250
