#!/usr/bin/env python
"""Turn the raw ncu artefacts in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py            # needs ncu on PATH (reads .ncu-rep files), no GPU
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def ncu_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    return d


def summarize_rep(name, rep, out_name):
    d = ncu_raw(rep)
    lines = [f"# {name}: {d.get('Kernel Name', ('?', ''))[0][:150]}", "metric,value,unit"]
    for k in KEYS:
        if k in d:
            lines.append(f"{k},{d[k][0]},{d[k][1]}")
    for k, (v, u) in sorted(d.items()):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            lines.append(f"{k},{v},{u}")
    open(os.path.join(OUT, out_name), "w").write("\n".join(lines) + "\n")
    return d


def launch_table(csv_path, out_name):
    with open(csv_path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        short = re.sub(r"\(.*", "", row["Kernel Name"].replace("(anonymous namespace)::", ""))[:100]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    out = [f"# ncu launch list: {n} launches, {tot:.0f} us total (cold-cache, serialised: compare SHARES)",
           "share_pct,total_us,launches,avg_us,kernel"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{t / tot * 100:.1f},{t:.1f},{c},{t / c:.1f},{k}")
    open(os.path.join(OUT, out_name), "w").write("\n".join(out) + "\n")
    return agg, tot


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    res = {}
    for name, rep, outn in [("GRU forward recurrence", "rec_fwd_final.ncu-rep", "r01_ncu_rec_fwd.csv"),
                            ("tcgen05 3xTF32 input-projection GEMM", "gemm_tc_final.ncu-rep", "r01_ncu_gemm_tc.csv"),
                            ("GRU backward recurrence", "rec_bwd_final.ncu-rep", "r01_ncu_rec_bwd.csv")]:
        p = os.path.join(SRC, rep)
        if os.path.exists(p):
            d = summarize_rep(name, p, outn)
            rd = float(d["dram__bytes_read.sum"][0]) * (1e6 if d["dram__bytes_read.sum"][1] == "Mbyte" else 1e3 if d["dram__bytes_read.sum"][1] == "Kbyte" else 1)
            wr = float(d["dram__bytes_write.sum"][0]) * (1e6 if d["dram__bytes_write.sum"][1] == "Mbyte" else 1e3 if d["dram__bytes_write.sum"][1] == "Kbyte" else 1)
            res[outn] = {"dram_bytes_per_launch": rd + wr, "duration_us": float(d["gpu__time_duration.sum"][0])}
    lp = os.path.join(SRC, "launches_r01_final.csv")
    if os.path.exists(lp):
        launch_table(lp, "r01_ncu_launch_list.csv")
    json.dump(res, open(os.path.join(OUT, "r01_ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))
