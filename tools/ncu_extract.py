#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/ncu_extract.py metrics  gpurun_out/prof.ncu-rep  profiles/<name>_ncu_metrics.json
    python tools/ncu_extract.py source   gpurun_out/prof.ncu-rep  profiles/<name>_ncu_source_top.txt
    python tools/ncu_extract.py launches gpurun_out/launches.csv  profiles/<name>_launch_shares.json

Runs where ncu is installed (no GPU needed: it only reads reports).
"""
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.max",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def metrics(rep, dst):
    hdr, units, launches = raw(rep)
    res = []
    for vals in launches:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name":
                d["kernel"] = v
            if h in KEEP or ("smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                try:
                    d[h] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    d[h] = {"value": v, "unit": u}
        res.append(d)
    json.dump({"report": rep, "launches": res}, open(dst, "w"), indent=1)
    print("wrote", dst)


def source(rep, dst, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    cur, hdr, agg = None, None, []
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 4 and r[2] == "-":
            d = dict(zip(hdr[4:], r[4:]))
            agg.append((cur, int(r[0]), r[1].strip()[:80], int(d["# Samples"]), int(d["Instructions Executed"]),
                        int(d["L1 Wavefronts Shared"]), int(d["L1 Wavefronts Shared Excessive"]),
                        int(d["L2 Theoretical Sectors Global"])))
    ts, ti, tw = (sum(a[i] for a in agg) or 1 for i in (3, 4, 5))
    with open(dst, "w") as f:
        f.write(f"# {rep}: per source line, sorted by warp-state samples (total {ts}), instructions {ti}, "
                f"shared wavefronts {tw}\n# file:line  samples%  inst%  smem-wavefront%  excessive-wavefronts  global-sectors | source\n")
        for a in sorted(agg, key=lambda a: -a[3])[:top]:
            f.write(f"{a[0]}:{a[1]:<4d} {100 * a[3] / ts:5.1f} {100 * a[4] / ti:5.1f} {100 * a[5] / tw:5.1f} "
                    f"{a[6]:>11d} {a[7]:>12d} | {a[2]}\n")
    print("wrote", dst)


def launches(path, dst):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = {}, {}
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("b200::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
        tot[name] = tot.get(name, 0.0) + v
        cnt[name] = cnt.get(name, 0) + 1
    total = sum(tot.values())
    out = [{"kernel": k, "launches": cnt[k], "total_us": round(v, 1), "share": round(v / total, 4)}
           for k, v in sorted(tot.items(), key=lambda kv: -kv[1])]
    json.dump({"source": path, "total_us": round(total, 1), "kernels": out}, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    {"metrics": metrics, "source": source, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
