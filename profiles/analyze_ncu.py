#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall reasons, and stall samples per source line.
usage: python profiles/analyze_ncu.py gpurun_out/prof.ncu-rep [source-file] [top]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
srcfile = sys.argv[2] if len(sys.argv) > 2 else "esdg_cns_b200/csrc/kernels_hex.cuh"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg"]
print("kernel:", data[0][hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for name in want:
    if name in hdr:
        i = hdr.index(name)
        print(f"  {name:72s} {units[i]:16s} {[r[i] for r in data]}")
print("stall cycles per issued instruction:")
for i, h in enumerate(hdr):
    if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        v = float(data[0][i])
        if v > 0.05:
            print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.3f}")
cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(cs)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
if hi:
    hdr = rows[hi[0]]
    ci = {h: i for i, h in enumerate(hdr)}
    nh = len(hdr)
    end = hi[1] if len(hi) > 1 else len(rows)
    samples, inst, stall = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    cur, tot = None, 0.0
    for r in rows[hi[0] + 1:end]:
        if not r:
            continue
        if r[0].strip().isdigit():
            cur = int(r[0])
            continue
        if len(r) >= 4 and r[2].startswith("0x"):
            off = len(r) - nh

            def g(nm):
                try:
                    return float(r[ci[nm] + off])
                except ValueError:
                    return 0.0
            s = g("# Samples")
            samples[cur] += s
            tot += s
            inst[cur] += g("Instructions Executed")
            for sn in names:
                stall[cur][sn] += g(sn)
    try:
        src = open(srcfile).read().split("\n")
    except OSError:
        src = []
    print(f"stall samples per source line of {srcfile} (total {int(tot)}):")
    for ln, s in samples.most_common(top):
        t3 = ", ".join(f"{k.replace('stall_', '')}:{int(v)}" for k, v in stall[ln].most_common(3))
        text = src[ln - 1].strip()[:90] if 0 < ln <= len(src) else ""
        print(f"  {ln:4d} {100 * s / tot:5.1f}% inst={inst[ln] / 1e6:7.1f}M [{t3}] {text}")
