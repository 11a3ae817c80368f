"""ncu target: MTTKRP launches on the per-GPU shards of BASELINE configs 4 and 3 (narrow column blocks)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package
pkg = load_package()
rng = np.random.default_rng(0)
for modes, C, todo in [((80, 80, 80, 80), 285, (1,)), ((299, 301, 41), 898, (0, 2))]:
    X = rng.uniform(-1, 1, size=modes)
    fs = [rng.uniform(-1, 1, size=(i, C)) for i in modes]
    with pkg.Engine(0) as eng:
        eng.set_tensor(X)
        for n in todo:
            G, ms = eng.mttkrp(fs, n, repeats=3)
            print(modes, C, "mode", n, "ms", ms, "TF", 2 * X.size * C / ms / 1e9)
