#!/usr/bin/env python
"""Warp-stall summary of an ncu report with source-level sampling (ncu --set full --import-source on):
per kernel the total samples per stall reason and the instructions with the most samples.

    tools/ncu_stalls.py <report.ncu-rep> <out.json> [top_n]
"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    res, name = [], None
    i = 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "Kernel Name":
            name = r[1]
        if r and r[0] == "Address":
            h = r
            cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
            idx = {c: h.index(c) for c in cols}
            i_s, i_src, i_ex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
            tot = {c: 0 for c in cols}
            instrs = []
            j = i + 1
            while j < len(rows) and rows[j] and rows[j][0] not in ("Address", "Kernel Name"):
                q = rows[j]
                if len(q) >= len(h):
                    try:
                        smp = int(q[i_s] or 0)
                    except ValueError:
                        smp = 0
                    for c in cols:
                        try:
                            tot[c] += int(q[idx[c]] or 0)
                        except ValueError:
                            pass
                    if smp:
                        top = max(cols, key=lambda c: int(q[idx[c]] or 0))
                        instrs.append({"n": j - i - 1, "samples": smp, "sass": q[i_src].strip()[:80], "top_stall": top,
                                       "executed": q[i_ex]})
                j += 1
            instrs.sort(key=lambda d: -d["samples"])
            res.append({"kernel": name, "samples": sum(tot.values()),
                        "stalls": dict(sorted(((k, v) for k, v in tot.items() if v), key=lambda kv: -kv[1])),
                        "top_instructions": instrs[:top_n]})
            i = j
            continue
        i += 1
    json.dump(res, open(out, "w"), indent=1)
    for k in res:
        print(k["kernel"][:70], k["samples"], list(k["stalls"].items())[:6])


if __name__ == "__main__":
    main()
