"""Short ncu target: two ALS iterations of BASELINE config 2 (200^3, 200 models, C = 2100) through cp_cals, so that the
pair-node kernels (pair_gemm_kernel, pair_leaf_*_kernel) are captured next to the MTTKRP and update kernels.
CONFIG=4 selects the 4-mode config (80^4, C = 2325): there the pair contractions are mttkrp_dmma_kernel launches."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package
pkg = load_package()
rng = np.random.default_rng(0)
cfg = int(os.environ.get("CONFIG", "2"))
modes, ranks = ((200, 200, 200), [r for r in range(1, 21) for _ in range(10)]) if cfg == 2 else \
               ((80, 80, 80, 80), [r for r in range(1, 31) for _ in range(5)])
X = np.asfortranarray(rng.uniform(-1, 1, size=modes))
kts = [pkg.Ktensor([np.asfortranarray(rng.uniform(-1, 1, size=(i, r))) for i in modes]) for r in ranks]
rep = pkg.cp_cals(X, kts, pkg.CalsParams(max_iterations=2, buffer_size=sum(ranks), force_max_iter=True))
print("iterations", rep.iter, "pair_node", rep.pair_node, "mean fit", float(np.mean([k.fit for k in kts])))
