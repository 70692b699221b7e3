#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall reasons, opcode mix and stall samples per source
line (all source files of the kernel).
usage: python profiles/analyze_ncu.py gpurun_out/prof.ncu-rep [top]"""
import collections, csv, io, os, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
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
samples, inst, stall, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
fpath, hdrs, cur = None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
        continue
    if r[0] == "Line No":
        hdrs = r
        ci = {h: i for i, h in enumerate(hdrs)}
        names = [h for h in hdrs if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdrs is None:
        continue
    if r[0].strip().isdigit():
        cur = (fpath, int(r[0]))
        continue
    if len(r) >= 4 and r[2].startswith("0x"):
        off = len(r) - len(hdrs)
        try:
            n = float(r[ci["# Samples"] + off])
            ni = float(r[ci["Instructions Executed"] + off])
        except ValueError:
            continue
        samples[cur] += n
        inst[cur] += ni
        t = r[3].strip().split()
        op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "")).split(".")[0]
        ops[op] += ni
        for h in names:
            try:
                stall[cur][h[6:]] += float(r[ci[h] + off])
            except ValueError:
                pass
tot, toti = sum(samples.values()), sum(inst.values())
print(f"total warp instructions {toti/1e6:.1f} M; opcode mix (%):", [(k, round(100 * v / toti, 1)) for k, v in ops.most_common(20)])
byfile = collections.Counter()
for (f, l), n in samples.items():
    byfile[f] += n
print("stall samples by file:", [(os.path.basename(f or "?"), int(n)) for f, n in byfile.most_common()])
srcs = {}
print(f"stall samples per source line (total {int(tot)}):")
for (f, ln), n in samples.most_common(top):
    if f not in srcs:
        try:
            srcs[f] = open(f).read().split("\n")
        except OSError:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:84] if 0 < ln <= len(srcs[f]) else ""
    st = ", ".join(f"{k}:{int(v)}" for k, v in stall[(f, ln)].most_common(3))
    print(f"  {os.path.basename(f or '?')[:18]:18s}{ln:4d} {100*n/tot:5.1f}% inst={inst[(f, ln)]/1e6:7.1f}M [{st}] {text}")
