#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: counts of the instructions that prove what the product library runs on --
DMMA (FP64 tensor cores), UTMALDG (TMA loads), SYNCS (mbarrier), LDS/STS, DFMA/DMUL, SHFL, MUFU -- from
`cuobjdump -sass cp-cals_b200/libcals_b200.so`.  No GPU needed.

    python tools/sass_excerpt.py [out.json]        (default profiles/sass_r02.json)
"""
import collections
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cp-cals_b200", "libcals_b200.so")
WATCH = ["DMMA", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "LDL", "STL", "DFMA", "DMUL", "DADD", "SHFL",
         "MUFU", "BAR", "ATOM", "RED", "USETMAXREG", "HMMA", "IMMA", "UTCHMMA", "UTCIMMA"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_r02.json")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("calsb200::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["_instructions"] += 1
            op = m.group(1)
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    with open(LIB, "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()[:16]
    sys.path.insert(0, ROOT)
    import bench
    res = {"library": "cp-cals_b200/libcals_b200.so", "library_sha256_16": sha, "arch": arch,
           "kernel_fingerprint": bench.kernel_fingerprint(),
           "how": "cuobjdump -sass | per-function opcode counts (tools/sass_excerpt.py)",
           "totals": {w: sum(k[w] for k in kernels.values()) for w in WATCH if sum(k[w] for k in kernels.values())},
           "kernels": {n: {"instructions": c["_instructions"], **{w: c[w] for w in WATCH if c[w]}}
                       for n, c in kernels.items()}}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", out, "arch", arch, "totals", res["totals"])


if __name__ == "__main__":
    main()
