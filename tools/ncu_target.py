"""Short ncu target: a few launches of the MTTKRP kernel on BASELINE config 2 (200^3, C = 2100) through the C-ABI hook."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package
pkg = load_package()
rng = np.random.default_rng(0)
modes, C = (200, 200, 200), 2100
X = rng.uniform(-1, 1, size=modes)
fs = [rng.uniform(-1, 1, size=(i, C)) for i in modes]
with pkg.Engine(0) as eng:
    eng.set_tensor(X)
    for n in (0, 1):
        G, ms = eng.mttkrp(fs, n, repeats=2)
        print("mode", n, "ms", ms, "checksum", float(np.abs(G).sum()))
