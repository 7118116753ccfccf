#!/usr/bin/env python
"""Turns one round of GPU captures (gpurun_out/) into the tracked summaries under profiles/:
    python tools/profile_report.py v8 "chained loop, split spectrum layout, graph replay"
reads gpurun_out/prof_<tag>.ncu-rep (ncu --set full), gpurun_out/r01_launches_<tag>.csv (launch list) and
gpurun_out/bench_r01_<tag>.json, writes profiles/r01_ncu_fast_<tag>.md, profiles/roofline_traffic.json and copies
the launch list and the bench line."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_%s.ncu-rep" % tag), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
S, C = 268.435456, 270.532608


def key(name):
    for pat, k, alg in (("k_strided<512, 0>", "fast_y_fwd", 2 * C), ("k_strided<512, 1>", "fast_y_inv", 2 * C),
                        ("k_strided<512, 2>", "fast_z_mul", 3 * C), ("k_rows_inv_fwd<128, 1>", "fast_rows_inv_quotient_fwd", 2 * C + S),
                        ("k_rows_inv_fwd<128, 2>", "fast_rows_inv_update_fwd", 2 * C + 3 * S), ("k_rows_fwd2", "fast_rows_fwd", S + C),
                        ("k_rows_inv2<128, 1>", "fast_rows_inv_quotient", C + 2 * S), ("k_rows_inv2<128, 2>", "fast_rows_inv_update", C + 3 * S)):
        if pat in name:
            return k, alg
    return name, 0.0


def mb(v, u):
    return float(v) * (1000.0 if u.startswith("G") else 1.0)


stall = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
out = ["# ncu --set full, fast path %s (%s), B200, config 3\n" % (tag, what),
       "Command (after the same command exited 0 without ncu): `ncu --set full --clock-control none --import-source on -k "
       "\"regex:k_rows|k_strided\" -s 8 -c 8 python tools/kbench.py 512,512,256 1`. Launch list of the bench command: "
       "`profiles/r01_ncu_launches_fast_%s.csv` (`ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 100 "
       "python bench.py --steps 1 --warmup 3 --iterations 4 --no-e2e --no-cpu-baseline`).\n" % tag,
       "Per-launch values (cold-cache, serialised under the profiler: compare shares, not absolutes). Algorithmic bytes: "
       "S = 268.4 MB, C = 270.5 MB.\n",
       "| kernel | time us | dram read MB | dram write MB | alg bytes MB | dram % of ncu peak | lts % | l1tex % | issue % | warps active % | regs | top stalls |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = collections.defaultdict(list)
for d in data:
    name = d[idx["Kernel Name"]].replace("void ", "").split("(")[0]
    k, alg = key(name)
    f = lambda kk: float(d[idx[kk]])
    tot = sum(float(d[idx[h]] or 0) for h in stall) or 1.0
    top = sorted(stall, key=lambda h: -float(d[idx[h]] or 0))[:3]
    tops = ", ".join("%s %.0f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * float(d[idx[h]]) / tot) for h in top)
    rd = mb(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = mb(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    traffic[k].append((rd + wr) * 1e6)
    out.append("| %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s | %s |" % (
        name, f("gpu__time_duration.sum"), rd, wr, alg, f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        f("lts__throughput.avg.pct_of_peak_sustained_elapsed"), f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"), f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        d[idx["launch__registers_per_thread"]], tops))
ll = os.path.join(G, "r01_launches_%s.csv" % tag)
if os.path.exists(ll):
    agg = collections.OrderedDict()
    for r in csv.reader(open(ll)):
        if len(r) > 10 and r[0].isdigit():
            a = agg.setdefault(r[4], [0.0, 0])
            a[0] += float(r[-1]); a[1] += 1
    tot = sum(a[0] for a in agg.values())
    out += ["\nLaunch list of the bench command (100 consecutive launches of the steady state):\n",
            "| kernel | launches | avg us | share of captured time |", "|---|---|---|---|"]
    for k, a in agg.items():
        out.append("| %s | %d | %.1f | %.1f %% |" % (k.replace("void ", "")[:60], a[1], a[0] / a[1] / 1e3, 100 * a[0] / tot))
    shutil.copy(ll, os.path.join(P, "r01_ncu_launches_fast_%s.csv" % tag))
bj = os.path.join(G, "bench_r01_%s.json" % tag)
if os.path.exists(bj):
    b = json.load(open(bj))
    out.append("\n`bench.py` of the same build (no profiler): value %.2f G voxel*view*iter/s (%.1f ms per 50-iteration step), whole-step "
               "%.1f %% of the 7S+10C roofline at %.0f GB/s, dominant kernel %s at %.1f %% (share of step %.1f %%), e2e %.2f (%.1f ms), "
               "SM clock %.0f MHz, reasons %s.\n" % (
                   b["value"], b["ms_per_step"], 100 * b["roofline"]["whole_step"]["frac"], b["roofline"]["peak"], b["roofline"]["kernel"],
                   100 * b["roofline"]["frac"], 100 * b["roofline"]["share_of_step"], b["e2e"]["value"], b["e2e"]["ms_per_step"],
                   b["clocks"]["sm_mhz"], b["clocks"]["reasons"]))
    shutil.copy(bj, os.path.join(P, "r01_bench_fast_%s.json" % tag))
open(os.path.join(P, "r01_ncu_fast_%s.md" % tag), "w").write("\n".join(out) + "\n")
tj = {"_source": "profiles/r01_ncu_fast_%s.md: dram__bytes_read.sum + dram__bytes_write.sum per launch (mean of the captured launches), "
                 "ncu --set full, B200, config 3 (bytes)" % tag}
for k, v in traffic.items():
    tj[k] = int(sum(v) / len(v))
json.dump(tj, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=2)
print("\n".join(out[-14:]))
