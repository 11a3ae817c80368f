#!/usr/bin/env python
"""Summarise ncu output into the small text/JSON files kept under profiles/.

    tools/ncu_summary.py full  <report.ncu-rep> <out.json>     selected counters of every captured launch
    tools/ncu_summary.py list  <launches.csv>   <out.json>     per-kernel totals and shares of a launch list
                                                               (ncu --metrics gpu__time_duration.sum --csv)
    tools/ncu_summary.py traffic <report.ncu-rep> <fingerprint file> <out.json>
                                                               DRAM bytes per launch of the config-2 contractions,
                                                               stamped with the kernel fingerprint (tools/refresh_traffic.sh)
Runs where ncu is installed (no GPU needed: it only reads the report).
"""
import csv
import io
import json
import re
import subprocess
import sys

FULL_KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_active.avg",
    "smsp__cycles_active.avg",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum",
    "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def short_name(full):
    m = re.match(r"(?:void\s+)?([A-Za-z0-9_:]+(?:<[^>]*>)?)", full)
    return m.group(1) if m else full


def full(report, out):
    txt = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        d = {"kernel": short_name(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for k in FULL_KEYS:
            if k in col and r[col[k]] != "":
                d[k] = {"value": float(r[col[k]].replace(",", "")), "unit": units[col[k]]}
        if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = d["dram__bytes_read.sum"]["value"] * scale[d["dram__bytes_read.sum"]["unit"]]
            wr = d["dram__bytes_write.sum"]["value"] * scale[d["dram__bytes_write.sum"]["unit"]]
            d["dram_bytes_per_launch"] = rd + wr
        launches.append(d)
    json.dump({"source": report, "note": "ncu --set full --clock-control none; clocks under ncu are lower than in the "
               "timed runs, so compare cycle counts and percentages, not durations", "launches": launches},
              open(out, "w"), indent=1)
    print("wrote", out, len(launches), "launches")


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path)) if r and r[0].strip('"').isdigit()]
    # columns: ID, PID, process, host, kernel, ctx, stream, block, grid, device, cc, section, metric, unit, value
    agg, order, total = {}, [], 0.0
    for r in rows:
        name, val, unit = short_name(r[4]), float(r[-1].replace(",", "")), r[-2]
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        a = agg.setdefault(name, {"launches": 0, "ns": 0.0, "min_ns": 1e30, "max_ns": 0.0, "all": []})
        a["all"].append(ns)
        if a["launches"] == 0:
            order.append(name)
        a["launches"] += 1
        a["ns"] += ns
        a["min_ns"] = min(a["min_ns"], ns)
        a["max_ns"] = max(a["max_ns"], ns)
        total += ns
    def busy(a):  # launches that did work: the host runs a few passes ahead of the `done` flag, and the kernels of a
        v = [x for x in a["all"] if x >= 0.2 * a["max_ns"]]  # pass that finds the queue drained exit at once
        return {"busy_launches": len(v), "busy_avg_us": sum(v) / len(v) / 1e3}
    kernels = [dict({"kernel": n, "launches": agg[n]["launches"], "total_ms": agg[n]["ns"] / 1e6,
                     "avg_us": agg[n]["ns"] / agg[n]["launches"] / 1e3, "min_us": agg[n]["min_ns"] / 1e3,
                     "max_us": agg[n]["max_ns"] / 1e3, "share": agg[n]["ns"] / total}, **busy(agg[n])) for n in order]
    kernels.sort(key=lambda k: -k["total_ms"])
    json.dump({"source": path, "note": "ncu --metrics gpu__time_duration.sum --clock-control none: per-launch times "
               "are cold-cache and serialised -- read the SHARES; avg_us includes the passes the host launches after the "
               "queue has drained (kernels exit at once), busy_avg_us does not", "launches": len(rows), "total_ms": total / 1e6,
               "kernels": kernels}, open(out, "w"), indent=1)
    print("wrote", out)
    for k in kernels:
        print("%-40s %6d launches %10.3f ms  %5.1f%%  busy avg %8.1f us x %d" % (k["kernel"], k["launches"], k["total_ms"],
                                                                             100 * k["share"], k["busy_avg_us"], k["busy_launches"]))


def traffic(report, fingerprint_file, out):
    """profiles/traffic_r02.json: what bench.py's roofline.traffic reads (only while the fingerprint matches)."""
    import datetime
    import os
    tmp = out + ".full.tmp"
    full(report, tmp)
    launches = json.load(open(tmp))["launches"]
    os.remove(tmp)
    fp = open(fingerprint_file).read().strip()
    kernels, seen = [], set()
    for l in launches:
        name = l["kernel"]
        if name in seen or "dram_bytes_per_launch" not in l:
            continue
        seen.add(name)
        kernels.append({"kernel": name, "workload": "config2", "dram_bytes_per_launch": l["dram_bytes_per_launch"],
                        "dram_read": l["dram__bytes_read.sum"], "dram_write": l["dram__bytes_write.sum"],
                        "duration": l.get("gpu__time_duration.sum"),
                        "dmma_pipe_pct": l.get("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
                        "grid": l["grid"], "block": l["block"]})
    json.dump({"kernel_fingerprint": fp, "captured": datetime.date.today().isoformat(),
               "how": "tools/refresh_traffic.sh: ncu --set full --clock-control none of tools/ncu_target_cfg.py 2 1 3 "
                      "(BASELINE config 2, 200 models, C = 2100), one launch per kernel; dram__bytes_read.sum + "
                      "dram__bytes_write.sum", "kernels": kernels}, open(out, "w"), indent=1)
    print("wrote", out, fp, [(k["kernel"], k["dram_bytes_per_launch"]) for k in kernels])


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"full": full, "list": launch_list}[cmd](sys.argv[2], sys.argv[3])
