"""ncu target: one cp_cals pass of BASELINE config 2 (2 forced iterations) so that every kernel of the path launches:
prep (swap01_copy, rowsumsq_*), init_grams, sched, move, mttkrp_dmma, mttkrp_reduce, model_update."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import load_package
pkg = load_package()
cfg = bench.CONFIGS[2]
X, models, jk = bench.workload(cfg, 0)
C = sum(fs[0].shape[1] for fs in models)
with pkg.Engine(0) as eng:
    eng.set_tensor(X)
    eng.configure(C, 2, 1e-7, force_max_iter=True)
    eng.clear_models()
    for fs in models:
        eng.enqueue(fs)
    rep = eng.run()
    print("iter", rep.iter, "launches", rep.kernel_launches, "fit0", eng.fetch(0)[2].fit)
