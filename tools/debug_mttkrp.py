import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from conftest import load_package
import oracle
pkg = load_package()
np.set_printoptions(linewidth=200, precision=2)
rng = np.random.default_rng(0)
cases = [((64, 24, 50), 300), ((100, 100, 100), 220), ((96, 40, 8), 256), ((48, 40, 8), 512), ((96, 80, 8), 64)]
with pkg.Engine(0) as eng:
    for modes, C in cases:
        X = rng.uniform(-1, 1, size=modes)
        fs = [rng.uniform(-1, 1, size=(i, C)) for i in modes]
        eng.set_tensor(X)
        for n in range(len(modes)):
            want = oracle.mttkrp(X, fs, n)
            got, _ = eng.mttkrp(fs, n)
            err = np.abs(got - want)
            In = modes[n]
            # block error map: rows in blocks of 8, cols in blocks of 32
            rb = [err[r:r+8].max() for r in range(0, In, 8)]
            cb = [err[:, c:c+32].max() for c in range(0, C, 32)]
            ratio = np.where(np.abs(want) > 1e-9, got / np.where(want == 0, 1, want), 0)
            print("case", modes, C, "mode", n, "max err %.2e" % err.max(), "median ratio got/want %.3f" % np.median(ratio))
            if err.max() > 1e-9:
                print("  row-block(8) max err:", np.array(rb))
                print("  col-block(32) max err:", np.array(cb))
