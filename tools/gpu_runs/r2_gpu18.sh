mkdir -p gpurun_out
B200_CFG_SCALE=2 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:corr1d|corr2d' -c 40 --csv --log-file gpurun_out/r2_launches_c4.csv python tools/bench_configs.py C4a C4b > gpurun_out/r2_ncu_c4.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_c4.csv')) if r and not r[0].startswith('==')]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
for r in rows[1:]:
    try: print(r[ix['ID']], r[ix['Kernel Name']][:60], r[ix['Grid Size']], r[ix['Metric Value']], r[ix['Metric Unit']])
    except Exception: pass
PY
