"""Quick MTTKRP timing probe through the C ABI test hook: prints achieved FP64 TFLOP/s per mode."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package  # noqa: E402

pkg = load_package()


def main():
    shapes = [((100, 100, 100), 220), ((200, 200, 200), 2100)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        shapes += [((80, 80, 80, 80), 2325), ((299, 301, 41), 2691)]
    rng = np.random.default_rng(0)
    with pkg.Engine(0) as eng:
        for modes, C in shapes:
            X = rng.uniform(-1, 1, size=modes)
            fs = [rng.uniform(-1, 1, size=(i, C)) for i in modes]
            eng.set_tensor(X)
            for n in range(len(modes)):
                G, ms = eng.mttkrp(fs, n, repeats=6)
                flops = 2.0 * X.size * C
                print(json.dumps({"modes": modes, "C": C, "mode": n, "ms": round(ms, 4),
                                  "tflops": round(flops / (ms * 1e-3) / 1e12, 3) if ms > 0 else None,
                                  "checksum": float(np.abs(G).sum())}), flush=True)


if __name__ == "__main__":
    main()
