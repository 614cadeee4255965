python -m pytest tests -m gpu -q > gpurun_out/pytest25.log 2>&1; tail -4 gpurun_out/pytest25.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -2 gpurun_out/bench_r1_n1.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_r1_n1.json') if l.startswith('{')][-1]);print({k:d[k] for k in ('value','ms_per_step','op_level_roofline_frac','gpu_launches')}, d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['traffic'])"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-250
python tools/prof_kernels.py f32 2 1 > gpurun_out/p1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:contig_static -s 1 -c 1 -o gpurun_out/prof_r1_contig_static_s2 python tools/prof_kernels.py f32 2 1 > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:strided -s 1 -c 1 -o gpurun_out/prof_r1_strided_s2 python tools/prof_kernels.py f32 2 1 > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:contig_static -s 1 -c 1 -o gpurun_out/prof_r1_contig_static_s4 python tools/prof_kernels.py f32 4 1 > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:contig_static -s 1 -c 1 -o gpurun_out/prof_r1_box_contig_u16 python tools/prof_kernels.py u16 2 1 > gpurun_out/ncu_h.log 2>&1
ls -la gpurun_out/*.ncu-rep | wc -l
