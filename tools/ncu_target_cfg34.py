"""ncu target: one MTTKRP launch per mode at the full column counts of BASELINE configs 4 (80^4, C = 2325) and 3
(299 x 301 x 41, C = 7176)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package
pkg = load_package()
rng = np.random.default_rng(0)
for modes, C, todo in [((80, 80, 80, 80), 2325, (0, 2)), ((299, 301, 41), 7176, (0, 2))]:
    X = rng.uniform(-1, 1, size=modes)
    fs = [rng.uniform(-1, 1, size=(i, C)) for i in modes]
    with pkg.Engine(0) as eng:
        eng.set_tensor(X)
        for n in todo:
            G, ms = eng.mttkrp(fs, n, repeats=2)
            print(modes, C, "mode", n, "ms", ms, "TF", 2 * X.size * C / ms / 1e9)
