"""MTTKRP time against the column count through the C ABI test hook (contraction + reduce pass, mode 0 of 200^3):
how much a short last octet costs, with and without the column tail (CALS_B200_NO_TAIL=1)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package  # noqa: E402

pkg = load_package()


def main():
    modes = (200, 200, 200)
    cols = [int(a) for a in sys.argv[1:]] or [128, 192, 248, 255, 256, 257, 263, 264, 280, 288, 296, 297, 320, 384, 512,
                                              520, 525, 528, 576]
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, size=modes)
    with pkg.Engine(0) as eng:
        eng.set_tensor(X)
        for C in cols:
            fs = [rng.uniform(-1, 1, size=(i, C)) for i in modes]
            for n in (0,):
                G, ms = eng.mttkrp(fs, n, repeats=8)
                flops = 2.0 * X.size * C
                print(json.dumps({"C": C, "mode": n, "ms": round(ms, 4), "us_per_col": round(1e3 * ms / C, 4),
                                  "tflops": round(flops / (ms * 1e-3) / 1e12, 2),
                                  "tail": os.environ.get("CALS_B200_NO_TAIL") is None}), flush=True)


if __name__ == "__main__":
    main()
