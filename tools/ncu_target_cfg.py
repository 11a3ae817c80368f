"""ncu target: any BASELINE configuration (optionally only the first shard of an N-way model split) through the engine,
a few forced iterations.    python tools/ncu_target_cfg.py <config> [shard_of] [iterations]"""
import importlib
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import load_package
pkg = load_package()
dmod = importlib.import_module("cp_cals_b200.distributed")
cid = int(sys.argv[1])
shard_of = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = bench.CONFIGS[cid]
X, models, jk = bench.workload(cfg, 0)
if shard_of > 1:
    mine = dmod.shard_models([fs[0].shape[1] for fs in models], shard_of)[0]
    models, jk = [models[i] for i in mine], [jk[i] for i in mine]
C = sum(fs[0].shape[1] for fs in models)
with pkg.Engine(0) as eng:
    eng.set_tensor(X)
    eng.configure(C, iters, 1e-7, force_max_iter=True)
    eng.clear_models()
    eng.enqueue_many(models, jk)
    rep = eng.run()
    print("config", cid, "shard_of", shard_of, "models", len(models), "C", C, "iter", rep.iter, "launches",
          rep.kernel_launches, "device_ms", rep.device_ms)
