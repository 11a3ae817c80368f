"""Write the case file of a BASELINE configuration (what bench.py hands to examples/bench_e2e.cpp and to the CPU
reference):   python tools/make_case.py <config> <out.case> [shard_of]"""
import importlib
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench
import caseio
cid = int(sys.argv[1])
cfg = bench.CONFIGS[cid]
X, models, jk = bench.workload(cfg, 0)
if len(sys.argv) > 3 and int(sys.argv[3]) > 1:
    from conftest import load_package
    load_package()
    dmod = importlib.import_module("cp_cals_b200.distributed")
    mine = dmod.shard_models([fs[0].shape[1] for fs in models], int(sys.argv[3]))[0]
    models, jk = [models[i] for i in mine], [jk[i] for i in mine]
ms = [caseio.Model(factors=fs, jk_mode=j[0], jk_fiber=j[1]) for fs, j in zip(models, jk)]
caseio.write_case(sys.argv[2], X, ms, max_iter=cfg["als_iters"], force_max_iter=True,
                  buffer_size=sum(m.rank for m in ms))
print("wrote", sys.argv[2], len(ms), "models")
