"""ncu target: BASELINE config 1 (100^3, 40 models, C = 220), 3 forced iterations -- the small, latency-bound case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import load_package
pkg = load_package()
cfg = bench.CONFIGS[1]
X, models, jk = bench.workload(cfg, 0)
C = sum(fs[0].shape[1] for fs in models)
with pkg.Engine(0) as eng:
    eng.set_tensor(X)
    eng.configure(C, 3, 1e-7, force_max_iter=True)
    eng.clear_models()
    eng.enqueue_many(models)
    rep = eng.run()
    print("iter", rep.iter, "launches", rep.kernel_launches)
