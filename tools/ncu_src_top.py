"""Summarise an `ncu --page source --csv` export: instruction mix of the hot instructions, shared-memory
excess wavefronts per instruction, top stall sites.   usage: ncu_src_top.py file.csv [N]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]


def num(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot_samples = sum(num(r, '# Samples') for r in body)
tot_inst = sum(num(r, 'Instructions Executed') for r in body)
print("instructions executed %.3g, samples %d" % (tot_inst, tot_samples))
mix = Counter()
smp = Counter()
for r in body:
    op = r[ix['Source']].split()[0] if r[ix['Source']].split() else '?'
    if op.startswith('@'):
        op = r[ix['Source']].split()[1]
    op = op.split('.')[0] if not op.startswith(('LDS', 'STS', 'STG', 'LDG')) else '.'.join(op.split('.')[:3])
    mix[op] += num(r, 'Instructions Executed')
    smp[op] += num(r, '# Samples')
print("-- mix (executed warp instructions, share; stall-sample share)")
for op, c in mix.most_common(18):
    print("  %-14s %6.2f%%   samples %6.2f%%" % (op, 100 * c / tot_inst, 100 * smp[op] / max(1, tot_samples)))
print("-- shared-memory wavefronts: excess by instruction")
ex = sorted(body, key=lambda r: -num(r, 'L1 Wavefronts Shared Excessive'))[:8]
for r in ex:
    if num(r, 'L1 Wavefronts Shared Excessive') > 0:
        print("  %-60s wavefronts %.3g ideal %.3g excess %.3g" % (r[ix['Source']].strip()[:60], num(r, 'L1 Wavefronts Shared'),
              num(r, 'L1 Wavefronts Shared Ideal'), num(r, 'L1 Wavefronts Shared Excessive')))
print("   total shared wavefronts %.4g (ideal %.4g)" % (sum(num(r, 'L1 Wavefronts Shared') for r in body),
                                                         sum(num(r, 'L1 Wavefronts Shared Ideal') for r in body)))
print("-- top stall sites")
stall_cols = [h for h in hdr if h.startswith('stall_')]
for r in sorted(body, key=lambda r: -num(r, '# Samples'))[:n]:
    st = sorted(((num(r, c), c) for c in stall_cols), reverse=True)[:2]
    print("  %5.2f%%  %-58s %s" % (100 * num(r, '# Samples') / tot_samples, r[ix['Source']].strip()[:58],
                                   ", ".join("%s=%d" % (c[6:], v) for v, c in st if v > 0)))
agg = Counter()
for r in body:
    for c in stall_cols:
        agg[c] += num(r, c)
tot = sum(agg.values())
print("-- stall reasons overall: " + ", ".join("%s %.1f%%" % (c[6:], 100 * v / tot) for c, v in agg.most_common(8)))
