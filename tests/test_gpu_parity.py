"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden fixtures.

Tolerance (north_star): per-iteration fit and factor matrices within 1e-9 relative error in FP64.
"""
import numpy as np
import pytest

import caseio
import oracle
from helpers import RTOL, assert_models_close, load_golden, rel_err, to_ktensors

pytestmark = pytest.mark.gpu

MTTKRP_SHAPES = [
    ((13, 12, 11), 30),      # tiny, ragged everything
    ((40, 40, 40), 64),      # exact tiles in M/K
    ((100, 100, 100), 220),  # BASELINE config 1
    ((41, 30, 29), 7),       # odd extents (pitch padding), C < 8
    ((9, 8, 7, 6), 20),      # 4 modes
    ((20, 6, 5, 4, 3), 33),  # 5 modes
    ((64, 24, 50), 300),     # two n-tiles
    ((130, 17, 90), 513),    # several m-tiles, 3 n-tiles, K tail
    # ragged last octet: warps whose n8 group of the last octet starts beyond C run one slot less (csrc/mttkrp.cuh)
    ((200, 30, 20), 263),    # an 8-way shard of config 2: 5 octets in tiles of 2 and 3, one valid n8 group in the last
    ((56, 24, 10), 71),      # second octet of the only n-tile holds 7 columns; 7 m8 groups
    ((24, 9, 5, 4), 137),    # 4 modes (slow outer weights), third octet holds 9 columns
    ((48, 20, 12), 289),     # 4 octets + 33 columns: five valid groups in the fifth octet
    ((48, 20, 12), 296),     # ... exactly five groups
    ((48, 20, 12), 297),     # ... and one column into the sixth group
    ((33, 14, 9), 327),      # 6 octets in tiles of 3 and 3, the last holds 7 columns
    ((30, 10, 6, 5), 537),   # 4 modes, 9 octets in three n-tiles, 25 columns in the last octet
    ((72, 16, 11), 257),     # one column in the fifth octet; m-tiles of unequal height
    ((17, 16, 15), 1),       # a single column
]


@pytest.mark.parametrize("modes,C", MTTKRP_SHAPES)
@pytest.mark.parametrize("variant", ["naive", "dmma"])
def test_mttkrp_matches_oracle(pkg, engine, modes, C, variant):
    rng = np.random.default_rng(hash((modes, C)) % 2 ** 32)
    X = rng.uniform(-1, 1, size=modes)
    fs = [rng.uniform(-1, 1, size=(i, C)) for i in modes]
    engine.set_tensor(X)
    v = pkg.MTTKRP_NAIVE if variant == "naive" else pkg.MTTKRP_DMMA
    for n in range(len(modes)):
        want = oracle.mttkrp(X, fs, n)
        got, _ = engine.mttkrp(fs, n, variant=v)
        e = rel_err(got, want)
        assert e <= 1e-12, "%s mode %d rel err %.3e" % (variant, n, e)
        # column-wise as well: no column may be off even if the matrix norm hides it
        ce = np.linalg.norm(got - want, axis=0) / np.linalg.norm(want, axis=0)
        assert ce.max() <= 1e-11, "%s mode %d worst column %.3e at %d" % (variant, n, ce.max(), ce.argmax())


def test_norms(engine):
    rng = np.random.default_rng(11)
    for modes in [(13, 12, 11), (41, 30, 29), (100, 64, 50), (7, 6, 5, 4)]:
        X = rng.uniform(-1, 1, size=modes)
        engine.set_tensor(X)
        assert abs(engine.tensor_norm() - oracle.norm(X)) <= 1e-12 * oracle.norm(X)
        assert rel_err(engine.jk_norms(), oracle.jk_norms(X)) <= 1e-12


def test_khatri_rao_hook(engine):
    """mttkrp::khatri_rao(A, B) (reference src/utils/mttkrp.cpp:78-103): K[ib + IB*ia, c] = A[ia, c] * B[ib, c]."""
    rng = np.random.default_rng(12)
    for IA, IB, C in [(1, 1, 1), (7, 5, 3), (33, 64, 17), (200, 200, 21), (3, 1031, 2)]:
        A, B = rng.uniform(-1, 1, size=(IA, C)), rng.uniform(-1, 1, size=(IB, C))
        want = np.einsum("ac,bc->abc", A, B).reshape(IA * IB, C)  # row index = ia * IB + ib
        got = engine.khatri_rao(A, B)
        assert got.shape == want.shape and np.array_equal(got, want)  # one multiplication per entry: exact


def test_jackknife_flag_on_another_mode_is_refused(pkg):
    """Leave-one-out norms exist for mode 0 only (reference src/utils/utils.cpp:103-152, :40-51); a flag on another mode
    would silently use the norm of the wrong slice -- it is rejected at enqueue time."""
    rng = np.random.default_rng(13)
    X = rng.uniform(-1, 1, size=(6, 5, 4))
    ms = caseio.random_models(rng, X.shape, [2])
    with pkg.Engine(0) as eng:
        eng.set_tensor(X)
        eng.configure(2, 3, 1e-7)
        eng.clear_models()
        with pytest.raises(pkg.CalsB200Error, match="mode 0"):
            eng.enqueue(ms[0].factors, 1, 2)
        assert eng.enqueue(ms[0].factors, 0, 5) == 0


GOLDEN_CASES = ["forced_3d_k1", "forced_3d_k2", "forced_3d_k5", "forced_4d_queue", "jackknife_3d", "evict_first_3d"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("variant", ["naive", "dmma"])
def test_cp_cals_matches_golden(pkg, name, variant):
    X, ins, refs, params, report = load_golden(name)
    kts = to_ktensors(pkg, ins)
    p = pkg.CalsParams(max_iterations=params["max_iter"], tol=params["tol"], buffer_size=params["buffer_size"],
                       force_max_iter=bool(params.get("force_max_iter", False)),
                       always_evict_first=bool(params.get("always_evict_first", False)))
    rep = pkg.cp_cals(X, kts, p, mttkrp_variant=pkg.MTTKRP_NAIVE if variant == "naive" else pkg.MTTKRP_DMMA)
    assert rep.iter == report["iter"]
    assert rep.n_ktensors == report["n_ktensors"]
    assert rep.ktensor_comp_sum == report["comp_sum"]
    assert abs(rep.X_norm - report["x_norm"]) <= 1e-12 * report["x_norm"]
    assert_models_close(kts, refs, report["x_norm"], what=name)
    for g, r in zip(kts, refs):
        assert abs(g.fit_diff - r.fit_diff) <= RTOL


@pytest.mark.parametrize("name", ["nnls_3d_queue", "nnls_4d_tol"])
def test_cp_cals_nnls_matches_golden(pkg, name):
    """update_method == NNLS (reference src/utils/update.cpp:61-176): row-wise active-set solves on the device against
    the reference's outputs; every factor entry non-negative, active sets consistent with the zeros."""
    X, ins, refs, params, report = load_golden(name)
    kts = to_ktensors(pkg, ins)
    p = pkg.CalsParams(max_iterations=params["max_iter"], tol=params["tol"], buffer_size=params["buffer_size"],
                       force_max_iter=bool(params.get("force_max_iter", False)), update_method="nnls")
    rep = pkg.cp_cals(X, kts, p)
    assert rep.n_ktensors == report["n_ktensors"] and rep.ktensor_comp_sum == report["comp_sum"]
    if p.force_max_iter:
        assert rep.iter == report["iter"]
    assert_models_close(kts, refs, report["x_norm"], rtol=1e-8, check_iters=p.force_max_iter, what=name)
    for kt in kts:
        assert kt.chol_info == 0
        # lambda carries the sign/scale of the last mode; the other factors are normalised non-negative solutions
        for n, (F, A) in enumerate(zip(kt.factors, kt.active_set)):
            assert A.shape == F.shape
            assert not np.any(F[A] != 0.0)  # constrained entries are exactly zero


def test_nnls_active_sets_persist_across_calls(pkg):
    """Ktensor::active_set lives in the model (reference include/ktensor.h:36): two calls of k iterations each equal...
    not one call of 2k iterations in general (normalisation restarts at iters == 1), but the second call must start
    from the active sets the first one left -- checked against the oracle driven the same way."""
    rng = np.random.default_rng(12)
    modes = (10, 9, 8)
    gen = [rng.uniform(0, 1, size=(i, 3)) for i in modes]
    X = caseio.ktensor_to_tensor(gen, np.ones(3)) + 0.02 * rng.standard_normal(modes)
    ms = caseio.random_models(rng, modes, [4, 2])
    want = oracle.cp_cals(X, ms, max_iter=5, force_max_iter=True, nnls=True)
    kts = to_ktensors(pkg, ms)
    p = pkg.CalsParams(max_iterations=5, buffer_size=6, force_max_iter=True, update_method="nnls")
    pkg.cp_cals(X, kts, p)
    assert_models_close(kts, want.models, want.x_norm, rtol=1e-8, what="nnls first call")
    assert all(a is not None for kt in kts for a in kt.active_set)
    # second call: must run and stay feasible with the warm-started sets
    pkg.cp_cals(X, kts, p)
    for kt in kts:
        assert kt.chol_info == 0
        for F in kt.factors[:-1]:
            assert F.min() >= 0.0


@pytest.mark.parametrize("name", ["ls_noerr_3d_queue", "ls_errcheck_3d"])
def test_cp_cals_line_search_matches_golden(pkg, name):
    """Line search on the device (csrc/ls.cuh) against the reference's outputs, both methods."""
    X, ins, refs, params, report = load_golden(name)
    kts = to_ktensors(pkg, ins)
    method = {0: "no-error-checking", 1: "error-checking-serial"}[params["ls_method"]]
    p = pkg.CalsParams(max_iterations=params["max_iter"], tol=params["tol"], buffer_size=params["buffer_size"],
                       force_max_iter=True, line_search=True, line_search_method=method,
                       line_search_interval=params["ls_interval"], line_search_step=params["ls_step"])
    rep = pkg.cp_cals(X, kts, p)
    want = oracle.cp_cals(X, ins, **params)
    assert rep.iter == report["iter"] == want.iters
    assert (rep.ls_performed, rep.ls_failed) == (want.ls_performed, want.ls_failed)
    assert rep.ls_performed > 0
    assert_models_close(kts, refs, report["x_norm"], rtol=1e-8, what=name)


def test_cp_cals_queue_tol_golden(pkg):
    """tol-based stopping with queueing/eviction/compaction (buffer 14 < sum of ranks).  Stopping can flip by one
    iteration through rounding (the fast error is a cancellation), so iteration counts are compared leniently and
    factors only where the counts agree."""
    X, ins, refs, params, report = load_golden("queue_tol_3d")
    kts = to_ktensors(pkg, ins)
    p = pkg.CalsParams(max_iterations=params["max_iter"], tol=params["tol"], buffer_size=params["buffer_size"])
    rep = pkg.cp_cals(X, kts, p)
    assert rep.n_ktensors == report["n_ktensors"] and rep.ktensor_comp_sum == report["comp_sum"]
    same = [i for i, (g, r) in enumerate(zip(kts, refs)) if g.iters == r.iters]
    assert len(same) >= len(refs) - 2
    assert_models_close([kts[i] for i in same], [refs[i] for i in same], report["x_norm"], rtol=1e-7, what="queue_tol")
    for g, r in zip(kts, refs):
        assert abs(g.fit - (1 - abs(r.error) / report["x_norm"])) < 1e-4


@pytest.mark.parametrize("modes,ranks,K,buffer", [
    ((30, 20, 25), list(range(1, 11)) * 2, 6, None),        # everything resident
    ((30, 20, 25), list(range(1, 11)) * 2, 4, 25),           # queueing
    ((12, 9, 8, 7), [3, 1, 4, 1, 5, 9, 2, 6], 5, 12),        # 4 modes + queueing
    ((50, 41, 33), [20, 17, 1, 12], 5, None),                # larger ranks, odd extents
])
def test_cp_cals_per_iteration_vs_oracle(pkg, modes, ranks, K, buffer):
    """Per-iteration parity: forced iteration counts k = 1..K, each compared with the oracle."""
    rng = np.random.default_rng(hash((modes, K)) % 2 ** 32)
    X = rng.uniform(-1, 1, size=modes)
    ms = caseio.random_models(rng, modes, ranks)
    bs = buffer or sum(ranks)
    with pkg.Engine(0) as eng:
        for k in range(1, K + 1):
            want = oracle.cp_cals(X, ms, max_iter=k, force_max_iter=True, buffer_size=bs)
            kts = to_ktensors(pkg, ms)
            rep = pkg.cp_cals(X, kts, pkg.CalsParams(max_iterations=k, buffer_size=bs, force_max_iter=True), engine=eng)
            assert rep.iter == want.iters
            assert_models_close(kts, want.models, want.x_norm, what="k=%d" % k)
            for g, r in zip(kts, want.models):
                assert abs(g.fit - r.fit) <= RTOL and abs(g.old_fit - r.old_fit) <= RTOL


@pytest.mark.parametrize("modes,ranks,K,buffer", [
    ((23, 17, 29), [1, 2, 3, 4, 5, 6, 7], 6, None),             # ragged everything, one n-tile
    ((40, 41, 9), [3, 8, 2, 5], 5, 10),                          # queueing: live columns change between iterations
    ((64, 100, 48), list(range(1, 13)) * 4, 4, None),            # 312 columns: two n-tiles, several m-tiles of T
    ((299, 31, 41), [5, 9, 2], 3, None),                         # K tail of the contracted mode (299 = 7 * 40 + 19)
    ((9, 8, 7, 6), [1, 2, 3, 4, 5], 5, None),                    # 4 modes: two pair nodes, (0,1) and (2,3)
    ((12, 9, 8, 7), [3, 1, 4, 1, 5, 9, 2, 6], 5, 12),            # 4 modes + queueing
    ((20, 44, 9, 31), list(range(1, 10)) * 8, 3, None),          # 4 modes, 360 columns, odd pitches in both layouts
    ((40, 56, 24), [20] * 13 + [3], 3, None),                    # 263 columns: ragged last octet in MTTKRP and pair GEMM
    ((30, 12, 9, 16), [7] * 10 + [1], 3, None),                  # 4 modes, 71 columns: ragged octet with slow outer modes
    ((40, 52, 20), [20] * 14 + [9], 2, None),                    # 289 columns: five groups in the fifth octet, last m-tile of T ragged
    ((24, 20, 12), [16] * 20 + [7], 2, None),                    # 327 columns: two n-tiles of 3 octets, 7 columns in the last
    ((40, 41, 9), [20, 20, 20, 11, 20, 3, 20], 4, 71),           # queueing at 71 buffer columns: the ragged octet comes and goes
])
def test_pair_node_equals_per_mode_mttkrp(pkg, modes, ranks, K, buffer):
    """Pair nodes (csrc/pairnode.cuh): 3-mode tensors take the MTTKRPs of modes 1 and 2 from T = X_(0)^T A_0, 4-mode
    tensors those of modes (0,1) and (2,3) from one contraction each.  Same sums as one full MTTKRP per mode (reference
    src/cals.cpp:214-222), only associated differently: both paths against each other and against the oracle."""
    rng = np.random.default_rng(hash((modes, K, 7)) % 2 ** 32)
    X = rng.uniform(-1, 1, size=modes)
    ms = caseio.random_models(rng, modes, ranks)
    bs = buffer or sum(ranks)
    want = oracle.cp_cals(X, ms, max_iter=K, force_max_iter=True, buffer_size=bs)
    out = {}
    for method in ("auto", "mttkrp"):
        kts = to_ktensors(pkg, ms)
        rep = pkg.cp_cals(X, kts, pkg.CalsParams(max_iterations=K, buffer_size=bs, force_max_iter=True,
                                                 mttkrp_method=method))
        assert rep.pair_node == (method == "auto")
        assert rep.iter == want.iters
        assert_models_close(kts, want.models, want.x_norm, what=method)
        out[method] = kts
    for a, b in zip(out["auto"], out["mttkrp"]):
        for Fa, Fb in zip(a.factors, b.factors):
            assert rel_err(Fa, Fb) <= 1e-11
        assert abs(a.fit - b.fit) <= 1e-12


@pytest.mark.parametrize("fused", ["0", "1"])
@pytest.mark.parametrize("modes,ranks,K", [
    ((37, 45, 23), [4, 9, 1, 12, 7], 4),       # i1 not a multiple of 8, i2 not a multiple of any tile height
    ((40, 56, 24), [20] * 13 + [3], 3),        # 263 columns: two n-tiles
    ((16, 8, 100), [5, 6], 3),                 # one i1 block, many i2 blocks: many partial results per element
    ((16, 203, 3), [5, 6, 2], 3),              # many i1 blocks, a single short i2 block
])
def test_first_leaf_fused_and_separate(pkg, monkeypatch, fused, modes, ranks, K):
    """The first leaf of a 3-mode pair node either rides in the epilogue of the contraction (per-tile partial results,
    summed in fixed order by pair_partial_reduce_kernel) or runs as its own pass over T (csrc/pairnode.cuh;
    CALS_B200_FUSED_LEAF=0/1 forces either, the default depends on the buffer shape).  Both against the oracle."""
    monkeypatch.setenv("CALS_B200_FUSED_LEAF", fused)
    rng = np.random.default_rng(hash((modes, K, 11)) % 2 ** 32)
    X = rng.uniform(-1, 1, size=modes)
    ms = caseio.random_models(rng, modes, ranks)
    want = oracle.cp_cals(X, ms, max_iter=K, force_max_iter=True, buffer_size=sum(ranks))
    kts = to_ktensors(pkg, ms)
    rep = pkg.cp_cals(X, kts, pkg.CalsParams(max_iterations=K, buffer_size=sum(ranks), force_max_iter=True))
    assert rep.pair_node
    assert_models_close(kts, want.models, want.x_norm, what="fused leaf " + fused)


def test_pair_node_with_jackknife_nnls_and_line_search(pkg):
    """The pair node only replaces where G comes from: flagged (jackknife) models, the NNLS update and line search
    give the same results with and without it."""
    rng = np.random.default_rng(77)
    modes = (21, 19, 16)
    gen = [rng.uniform(0, 1, size=(i, 3)) for i in modes]
    X = caseio.ktensor_to_tensor(gen, np.ones(3)) + 0.05 * rng.standard_normal(modes)
    ms = caseio.random_models(rng, modes, [2, 3, 4, 3])
    ms[1].jk_mode, ms[1].jk_fiber = 0, 5
    ms[3].jk_mode, ms[3].jk_fiber = 0, 20
    cases = [dict(), dict(update_method="nnls"), dict(line_search=True, line_search_interval=3)]
    for extra in cases:
        res = {}
        for method in ("auto", "mttkrp"):
            kts = to_ktensors(pkg, ms)
            rep = pkg.cp_cals(X, kts, pkg.CalsParams(max_iterations=8, buffer_size=12, force_max_iter=True,
                                                     mttkrp_method=method, **extra))
            assert rep.pair_node == (method == "auto")
            res[method] = kts
        for a, b in zip(res["auto"], res["mttkrp"]):
            assert a.iters == b.iters
            for Fa, Fb in zip(a.factors, b.factors):
                assert rel_err(Fa, Fb) <= 1e-9, extra
            assert abs(a.fit - b.fit) <= 1e-10, extra


def test_cp_cals_large_ranks_vs_oracle(pkg):
    """Ranks as in BASELINE config 5 (up to 50) and beyond: the update kernel keeps two R x R matrices in shared
    memory and stages the factor rows in chunks."""
    rng = np.random.default_rng(50)
    modes = (70, 66, 61)
    gen = [rng.uniform(-1, 1, size=(i, 8)) for i in modes]
    X = caseio.ktensor_to_tensor(gen, np.ones(8)) + 0.1 * rng.standard_normal(modes)
    ranks = [50, 33, 64, 1]
    ms = caseio.random_models(rng, modes, ranks)
    K = 3
    want = oracle.cp_cals(X, ms, max_iter=K, force_max_iter=True)
    kts = to_ktensors(pkg, ms)
    rep = pkg.cp_cals(X, kts, pkg.CalsParams(max_iterations=K, buffer_size=sum(ranks), force_max_iter=True))
    assert rep.iter == K
    # rank 64 > every extent but one: H is still positive definite here (Hadamard of three Gramians), so everything compares
    assert all(k.chol_info == 0 for k in kts)
    assert_models_close(kts, want.models, want.x_norm, rtol=1e-8, what="large ranks")


def test_cp_cals_config1_shape_vs_oracle(pkg):
    """BASELINE config 1: 100x100x100, 40 models of ranks 1..10 x4, forced iterations."""
    rng = np.random.default_rng(1)
    modes = (100, 100, 100)
    X = rng.uniform(-1, 1, size=modes)
    ranks = [r for r in range(1, 11) for _ in range(4)]
    ms = caseio.random_models(rng, modes, ranks)
    K = 5
    want = oracle.cp_cals(X, ms, max_iter=K, force_max_iter=True)
    kts = to_ktensors(pkg, ms)
    rep = pkg.cp_cals(X, kts, pkg.CalsParams(max_iterations=K, buffer_size=sum(ranks), force_max_iter=True))
    assert rep.iter == K
    assert_models_close(kts, want.models, want.x_norm, what="cfg1")


def test_rerun_is_deterministic(pkg):
    rng = np.random.default_rng(2)
    modes = (40, 30, 20)
    X = rng.uniform(-1, 1, size=modes)
    ms = caseio.random_models(rng, modes, [3, 5, 2, 8, 1])
    with pkg.Engine(0) as eng:
        eng.set_tensor(X)
        eng.configure(19, 7, 1e-7, force_max_iter=True)
        eng.clear_models()
        for m in ms:
            eng.enqueue(m.factors)
        eng.run()
        first = [eng.fetch(i) for i in range(len(ms))]
        eng.rerun()
        second = [eng.fetch(i) for i in range(len(ms))]
        for (f1, l1, s1), (f2, l2, s2) in zip(first, second):
            assert all(np.array_equal(a, b) for a, b in zip(f1, f2))  # bit-exact: no atomics, fixed reduction order
            assert np.array_equal(l1, l2) and s1.error == s2.error and s1.iters == s2.iters


def test_jk_cp_cals_vs_oracle(pkg):
    """jk_cp_cals (reference src/cals.cpp:397-446, without the LSAP permutation) against the oracle's cp_cals on the
    same flagged sub-models."""
    rng = np.random.default_rng(6)
    modes = (8, 7, 6)
    fs = [rng.uniform(-1, 1, size=(i, 3)) for i in modes]
    X = caseio.ktensor_to_tensor(fs, np.ones(3)) + 0.01 * rng.standard_normal(modes)
    bases = [pkg.Ktensor(m.factors, m.lam) for m in caseio.random_models(rng, modes, [3, 2])]
    p = pkg.CalsParams(max_iterations=10, buffer_size=12, force_max_iter=True)
    rep, results = pkg.jk_cp_cals(X, bases, p)
    # the same through the oracle
    flat_in = []
    for b in bases:
        bb = b.copy().denormalize().normalize()
        for i in range(modes[0]):
            flat_in.append(caseio.Model(factors=[F.copy() for F in bb.factors], lam=bb.lam.copy(), jk_mode=0, jk_fiber=i))
    want = oracle.cp_cals(X, flat_in, max_iter=10, force_max_iter=True, buffer_size=12)
    assert rep.iter == want.iters
    flat_out = [m for g in results for m in g]
    for got, w in zip(flat_out, want.models):
        wk = pkg.Ktensor([F.copy() for F in w.factors], w.lam.copy(), 0, w.jk_fiber)
        wk.set_jk_fiber(0.0)
        wk.denormalize().normalize()
        for n in range(3):
            a, b = got.factors[n].copy(), wk.factors[n].copy()
            if n == 0:
                assert np.isnan(a[got.jk_fiber]).all()
                a[got.jk_fiber] = 0.0
            assert rel_err(a, b) <= RTOL
        assert rel_err(got.lam, wk.lam) <= RTOL


@pytest.mark.parametrize("seed", range(40))
def test_cp_cals_random_configurations_vs_oracle(pkg, seed):
    """Randomised sweep: number of modes, ragged extents, ranks, buffer sizes (queueing / eviction / compaction),
    jackknife flags and the update method are drawn at random; forced iteration counts so that the comparison is exact
    in the discrete outcomes (iterations, admissions) and within 1e-8 in the values."""
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.integers(3, 6))
    modes = tuple(int(x) for x in rng.integers(2, [40, 30, 20, 8, 6][:N], endpoint=True))
    n_models = int(rng.integers(1, 12))
    ranks = [int(r) for r in rng.integers(1, 9, size=n_models)]
    nnls = bool(rng.integers(0, 3) == 0)
    K = int(rng.integers(1, 7))
    buffer = int(rng.integers(max(ranks), sum(ranks) + 1))
    R0 = int(rng.integers(2, 5))
    gen = [rng.uniform(0 if nnls else -1, 1, size=(i, R0)) for i in modes]
    X = caseio.ktensor_to_tensor(gen, np.ones(R0)) + 0.05 * rng.standard_normal(modes)
    ms = caseio.random_models(rng, modes, ranks)
    if not nnls:
        for m in ms:
            if rng.integers(0, 4) == 0:  # leave-one-out flag on some models (any mode, as Ktensor::to_jk allows)
                m.jk_mode = int(rng.integers(0, N))
                m.jk_fiber = int(rng.integers(0, modes[m.jk_mode]))
                m.factors[m.jk_mode][m.jk_fiber, :] = 0.0
    # the reference keeps jackknife norms for mode 0 only (src/utils/utils.cpp:103): restrict flags to mode 0 fibres
    for m in ms:
        if m.jk_mode > 0:
            m.factors[m.jk_mode][m.jk_fiber, :] = rng.uniform(-1, 1, size=m.rank)
            m.jk_mode, m.jk_fiber = 0, int(rng.integers(0, modes[0]))
            m.factors[0][m.jk_fiber, :] = 0.0
    want = oracle.cp_cals(X, ms, max_iter=K, force_max_iter=True, buffer_size=buffer, nnls=nnls)
    kts = to_ktensors(pkg, ms)
    rep = pkg.cp_cals(X, kts, pkg.CalsParams(max_iterations=K, buffer_size=buffer, force_max_iter=True,
                                             update_method="nnls" if nnls else "unconstrained"))
    tag = "seed %d: modes %s ranks %s K %d buffer %d nnls %s" % (seed, modes, ranks, K, buffer, nnls)
    assert (rep.iter, rep.n_ktensors, rep.ktensor_comp_sum) == (want.iters, want.n_ktensors, want.comp_sum), tag
    clean = [i for i, w in enumerate(want.models) if not w.chol_fail and not kts[i].chol_info]
    assert len(clean) >= len(ms) - 2, tag  # singular systems (rank > extents) have nothing to compare
    assert_models_close([kts[i] for i in clean], [want.models[i] for i in clean], want.x_norm, rtol=1e-7, what=tag)
