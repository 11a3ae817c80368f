"""CPU tests: pin the oracle (oracle/cals_oracle.c) against the reference's outputs.

* tests/golden/*.npz were produced by the UNMODIFIED reference library (oracle/make_golden.py).
* where oracle/_ref/cals_ref exists (it is built from /root/reference by oracle/build_ref.sh and travels to the GPU
  box) the oracle is also compared live on fresh seeded cases, and the reference's own relational pins are re-checked
  (CALS == ALS, tests/cals/test_cals.cpp:13-86; fast error == explicit error, tests/als/test_als.cpp:125-145).
"""
import numpy as np
import pytest

import caseio
import oracle
from helpers import RTOL, assert_models_close, load_golden

GOLDEN_CASES = ["forced_3d_k1", "forced_3d_k2", "forced_3d_k5", "queue_tol_3d", "forced_4d_queue", "jackknife_3d",
                "evict_first_3d", "nnls_3d_queue", "nnls_4d_tol", "ls_noerr_3d_queue", "ls_errcheck_3d"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_golden(name):
    X, ins, refs, params, report = load_golden(name)
    res = oracle.cp_cals(X, ins, **params)
    assert res.iters == report["iter"]
    assert res.n_ktensors == report["n_ktensors"]
    assert res.comp_sum == report["comp_sum"]
    assert abs(res.x_norm - report["x_norm"]) <= 1e-13 * report["x_norm"]
    assert_models_close(res.models, refs, report["x_norm"], what=name)
    for g, r in zip(res.models, refs):
        assert abs(g.fit_diff - r.fit_diff) <= RTOL


def test_oracle_mttkrp_vs_einsum():
    rng = np.random.default_rng(3)
    for modes in [(6, 5, 4), (3, 4, 5, 2), (2, 3, 2, 3, 2)]:
        X = rng.uniform(-1, 1, size=modes)
        fs = [rng.uniform(-1, 1, size=(i, 7)) for i in modes]
        letters = "abcde"[:len(modes)]
        for n in range(len(modes)):
            expr = letters + "," + ",".join(l + "r" for k, l in enumerate(letters) if k != n) + "->" + letters[n] + "r"
            want = np.einsum(expr, X, *[F for k, F in enumerate(fs) if k != n])
            got = oracle.mttkrp(X, fs, n)
            assert np.abs(got - want).max() <= 1e-12


def test_oracle_norms():
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, size=(7, 5, 6))
    assert abs(oracle.norm(X) - np.linalg.norm(X)) < 1e-12
    want = np.array([np.sqrt((X ** 2).sum() - (X[i] ** 2).sum()) for i in range(7)])
    assert np.abs(oracle.jk_norms(X) - want).max() < 1e-12


def test_fast_error_equals_explicit_error():
    """reference tests/als/test_als.cpp:125-145 (Als.ComputeCorrectError): after 3 iterations the fast error equals
    ||X - M|| within 1e-10."""
    rng = np.random.default_rng(5)
    modes = (9, 3, 2)
    fs = [rng.uniform(-1, 1, size=(i, 5)) for i in modes]
    X = caseio.ktensor_to_tensor(fs, np.ones(5))
    ms = caseio.random_models(rng, modes, [5])
    res = oracle.cp_cals(X, ms, max_iter=3, force_max_iter=True)
    m = res.models[0]
    explicit = np.linalg.norm(X - caseio.ktensor_to_tensor(m.factors, m.lam))
    assert abs(m.error - explicit) < 1e-10


needs_ref = pytest.mark.skipif(not caseio.ref_available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("modes,ranks,K", [((10, 9, 8), [1, 2, 3, 4, 5, 6], 6), ((4, 5, 3, 6), [2, 1, 4, 3], 5),
                                            ((21, 4, 17), [7, 1, 9], 8)])
def test_oracle_vs_live_reference(modes, ranks, K):
    rng = np.random.default_rng(hash((modes, K)) % 2 ** 32)
    X = rng.uniform(-1, 1, size=modes)
    ms = caseio.random_models(rng, modes, ranks)
    for k in range(1, K + 1):  # per-iteration parity (north_star): rerun with a growing forced iteration count
        ref = caseio.run_reference(X, ms, max_iter=k, force_max_iter=True)
        res = oracle.cp_cals(X, ms, max_iter=k, force_max_iter=True)
        assert res.iters == ref.iters
        assert_models_close(res.models, ref.models, ref.x_norm, what="k=%d" % k)


@needs_ref
def test_reference_cals_equals_als():
    """reference tests/cals/test_cals.cpp:13-86 (SimpleCorrectness), reduced: CALS with a small buffer (queueing,
    eviction, compaction) equals one-model-at-a-time ALS on the reconstructed tensors (< 1e-11)."""
    rng = np.random.default_rng(7)
    modes = (13, 12, 11)
    fs = [rng.uniform(-1, 1, size=(i, 10)) for i in modes]
    X = caseio.ktensor_to_tensor(fs, np.ones(10))
    ranks = list(rng.permutation(np.repeat(np.arange(1, 9), 3)))
    ms = caseio.random_models(rng, modes, ranks)
    kw = dict(max_iter=200, tol=1e-5, buffer_size=20)
    cals = caseio.run_reference(X, ms, algo=caseio.ALGO_CALS, **kw)
    als = caseio.run_reference(X, ms, algo=caseio.ALGO_ALS, **kw)
    orc = oracle.cp_cals(X, ms, **kw)
    for a, b, o in zip(cals.models, als.models, orc.models):
        ta = caseio.ktensor_to_tensor(a.factors, a.lam)
        tb = caseio.ktensor_to_tensor(b.factors, b.lam)
        assert np.linalg.norm(ta - tb) < 1e-11
        assert a.iters == b.iters
        # the oracle follows the same trajectory (tol-based stopping may flip by one iteration only through rounding)
        if o.iters == a.iters:
            to = caseio.ktensor_to_tensor(o.factors, o.lam)
            assert np.linalg.norm(ta - to) < 1e-9 * max(1.0, np.linalg.norm(ta))
    assert sum(o.iters == a.iters for o, a in zip(orc.models, cals.models)) >= len(ms) - 2
