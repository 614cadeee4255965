mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2_bench_plain_for_ncu.json 2> gpurun_out/r2_bench_plain_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:corr1d|corr2d|uniform1d|corrnd|combine|generic' -c 100 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2_ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches_bench.csv
