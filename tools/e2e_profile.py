"""Where does the end-to-end time of one cp_cals call go?  (config 2 by default; run on the GPU box)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import load_package
pkg = load_package()
cfg = bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
X, models, jk = bench.workload(cfg, 0)
C = sum(fs[0].shape[1] for fs in models)
eng = pkg.Engine(0)
def T():
    return time.perf_counter()
for rep in range(4):
    t0 = T(); eng.set_tensor(X); t1 = T()
    eng.configure(C, cfg["als_iters"], 1e-7, force_max_iter=True); eng.clear_models()
    for fs, j in zip(models, jk):
        eng.enqueue(fs, j[0], j[1])
    t2 = T(); r = eng.run(); t3 = T()
    out = [eng.fetch(i) for i in range(len(models))]
    t4 = T()
    print("pass %d: set_tensor %.2f ms | enqueue %.2f ms | run %.2f ms (device loop %.2f ms) | fetch %.2f ms | total %.2f ms"
          % (rep, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, r.device_ms, (t4-t3)*1e3, (t4-t0)*1e3))
