"""Turn the ncu reports fetched into gpurun_out/ into the tracked summaries under
profiles/: a markdown table per capture (duration, DRAM traffic, issue / pipe
utilisation, occupancy limiters, top stall reasons), the per-voxel DRAM traffic
table bench.py reads (profiles/traffic.json), and SASS listings of the hot
kernels.  Usage: make_profiles.py <round tag> <name>=<report>:<voxels per launch> ..."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe busy %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe busy %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTA/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTA/SM limit (shared memory)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def raw_page(rep):
    if rep.endswith(".csv"):        # `ncu -i X.ncu-rep --page raw --csv` exported on the GPU box
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    if tag == "--sass-only":        # regenerate profiles/sass/ from dask_image_b200/build/*.o, nothing else
        return dump_sass()
    summaries(tag)
    dump_sass()


def summaries(tag):
    md = ["# ncu --set full summaries, round %s" % tag, "",
          "Captured with `ncu --set full --clock-control none --import-source on` under gpurun on a B200 "
          "(one launch per kernel, after the same command had run clean without ncu).  ncu serialises and "
          "replays launches with cold caches: durations are for shares, not bench values.", ""]
    traffic_path = os.path.join(PROF, "traffic.json")
    traffic = {"dram_bytes_per_voxel": {}, "source": {}}
    if os.path.exists(traffic_path):
        traffic = json.load(open(traffic_path))
    for spec in sys.argv[2:]:
        name, rest = spec.split("=", 1)
        rep, vox = rest.rsplit(":", 1)
        hdr, units, rows = raw_page(rep)
        for r in rows:
            get = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
            kname = get["Kernel Name"][0]
            md += ["## %s" % name, "", "`%s`" % kname, "", "| metric | value |", "|---|---|"]
            for k, label in KEYS:
                if k in get:
                    md.append("| %s | %s %s |" % (label, get[k][0], get[k][1]))
            rd = to_bytes(*get["dram__bytes_read.sum"])
            wr = to_bytes(*get["dram__bytes_write.sum"])
            md.append("| DRAM bytes / voxel (read + write) | %.3f (%d voxels per launch) |" % ((rd + wr) / float(vox), int(vox)))
            stalls = []
            for h, (v, u) in get.items():
                if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                    try:
                        stalls.append((float(v.replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                    except ValueError:
                        pass
            tot = sum(s for s, _ in stalls) or 1.0
            top = sorted(stalls, reverse=True)[:6]
            md.append("| warp-state samples (top) | %s |" % ", ".join("%s %.0f%%" % (n, 100 * s / tot) for s, n in top))
            md.append("")
            traffic["dram_bytes_per_voxel"][name] = (rd + wr) / float(vox)
            traffic["source"][name] = "profiles/%s_ncu_summary.md (%s)" % (tag, os.path.basename(rep))
    with open(os.path.join(PROF, "%s_ncu_summary.md" % tag), "w") as f:
        f.write("\n".join(md) + "\n")
    with open(traffic_path, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)


def dump_sass():
    """SASS listings of the hot kernels (cuobjdump of the object files of the in-tree build)."""
    sass_dir = os.path.join(PROF, "sass")
    os.makedirs(sass_dir, exist_ok=True)
    wanted = {
        "corr1d_strided.o": [r"corr1d_strided_tma_kernelIffLi128ELi8ELi4ELi3ELi1E",
                             r"corr1d_strided_tma_kernelIftLi256ELi8ELi4ELi3ELi3E",
                             r"corr1d_strided_peer_kernelIffLi128ELi8ELi4ELi3ELi1E",
                             r"corr1d_strided_ldg_kernelIffLi128ELi8ELi4ELi3ELi1E"],
        "corr1d_contig.o": [r"corr1d_contig_static_kernelIfLi17ELb0E", r"corr1d_contig_static_kernelItLi16ELb0E",
                            r"corr1d_contig_tma_kernelIffLi0ELi1E"],
        "corr2d_fused.o": [r"corr2d_fused_kernelILi17ELi17ELi32ELi8ELi8ELb1ELi4ELb1EfLb0E",
                           r"corr2d_fused_kernelILi33ELi33ELi32ELi8ELi8ELb1ELi4ELb1EfLb0E",
                           r"corr2d_fused_kernelILi17ELi17ELi32ELi8ELi8ELb1ELi4ELb0EfLb0E",
                           r"corr2d_fused_kernelILi7ELi12ELi32ELi8ELi8ELb1ELi4ELb1EtLb1E"],
        "corrnd_tiled.o": [r"corrnd_static_kernelILi9ELb0E"],
        "corr1d_int.o": [r"corr1d_int_strided_kernelItLi1E"],
        "corrnd_int.o": [r"corrnd_int_kernelItE"],
    }
    if os.environ.get("SASS_ONLY"):
        wanted = {k: v for k, v in wanted.items() if k in os.environ["SASS_ONLY"].split(",")}
    for obj, pats in wanted.items():
        path = os.path.join(ROOT, "dask_image_b200", "build", obj)
        if not os.path.exists(path):
            continue
        text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        blocks = re.split(r"(?=\n\s*Function : )", text)
        for pat in pats:
            for b in blocks:
                m = re.search(r"Function : (\S*%s\S*)" % pat, b)
                if m:
                    lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln) for ln in b.splitlines()
                             if re.search(r"/\*[0-9a-f]{4,}\*/|Function :", ln)]
                    with open(os.path.join(sass_dir, "%s.sass" % pat), "w") as f:
                        f.write("\n".join(lines) + "\n")
                    break


if __name__ == "__main__":
    main()
