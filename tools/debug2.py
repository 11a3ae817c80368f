import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from conftest import load_package
import oracle, caseio
from helpers import rel_err, to_ktensors
pkg = load_package()
np.set_printoptions(linewidth=220, precision=2)
rng = np.random.default_rng(0)
print("== A: determinism + naive vs dmma of the MTTKRP hook")
cases = [((48, 40, 8), 512), ((48, 40, 8), 256), ((30, 20, 25), 110), ((40, 30, 20), 19), ((64, 24, 50), 300)]
with pkg.Engine(0) as eng:
    for modes, C in cases:
        X = rng.uniform(-1, 1, size=modes)
        fs = [rng.uniform(-1, 1, size=(i, C)) for i in modes]
        eng.set_tensor(X)
        for n in range(len(modes)):
            want = oracle.mttkrp(X, fs, n)
            g1, _ = eng.mttkrp(fs, n)
            g2, _ = eng.mttkrp(fs, n)
            g3, _ = eng.mttkrp(fs, n, repeats=3)
            gn, _ = eng.mttkrp(fs, n, variant=pkg.MTTKRP_NAIVE)
            print(modes, C, "mode", n, "err1 %.1e err2 %.1e err3(rep) %.1e naive %.1e | g1==g2 %s | nan %d of %d; nan rows %s nan cols %s" % (
                np.nanmax(np.abs(g1 - want)), np.nanmax(np.abs(g2 - want)), np.nanmax(np.abs(g3 - want)), np.abs(gn - want).max(),
                np.array_equal(g1, g2), np.isnan(g1).sum(), g1.size, np.unique(np.where(np.isnan(g1))[0] // 8)[:12], np.unique(np.where(np.isnan(g1))[1] // 32)[:20]))
print("== B: cp_cals naive vs dmma vs oracle, k=1")
for modes, ranks, bs in [((30, 20, 25), list(range(1, 11)) * 2, None), ((40, 30, 20), [3, 5, 2, 8, 1], None)]:
    X = rng.uniform(-1, 1, size=modes)
    ms = caseio.random_models(rng, modes, ranks)
    bs = bs or sum(ranks)
    want = oracle.cp_cals(X, ms, max_iter=1, force_max_iter=True, buffer_size=bs)
    for vname, v in (("naive", pkg.MTTKRP_NAIVE), ("dmma", pkg.MTTKRP_DMMA)):
        kts = to_ktensors(pkg, ms)
        pkg.cp_cals(X, kts, pkg.CalsParams(max_iterations=1, buffer_size=bs, force_max_iter=True), mttkrp_variant=v)
        errs = [[rel_err(g.factors[n], w.factors[n]) for n in range(3)] for g, w in zip(kts, want.models)]
        print(modes, vname, "max factor err per mode:", np.array(errs).max(axis=0), " worst model per mode:", np.array(errs).argmax(axis=0))
