"""Summarise `ncu --page source --csv` output: instruction mix by opcode
(executed warp-instructions) and the hottest stall sites."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, ist, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ops = collections.Counter()
samples = collections.Counter()
total_ex = 0
body = []
for r in rows[2:]:
    try:
        ex = int(r[iex]); sm = int(r[ist])
    except Exception:
        continue
    src = r[isrc].strip()
    toks = src.split()
    op = toks[0]
    if op.startswith("@"):
        op = toks[1]
    op = op.split(".")[0] if not op.startswith(("FFMA2", "LDS", "STG", "LDG", "LDC")) else op
    ops[op] += ex
    samples[op] += sm
    total_ex += ex
    body.append((ex, sm, r[ia][-5:], src))
print("total warp-instructions executed: %d" % total_ex)
for op, c in ops.most_common(22):
    print("  %-18s %12d  %5.1f%%   samples %d" % (op, c, 100.0 * c / total_ex, samples[op]))
if len(sys.argv) > 2:
    print("-- hottest stall sites")
    for ex, sm, a, src in sorted(body, key=lambda t: -t[1])[:int(sys.argv[2])]:
        print("  %6d samples  %10d exec  %s  %s" % (sm, ex, a, src[:90]))
