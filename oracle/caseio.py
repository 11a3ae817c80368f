"""TEST INFRASTRUCTURE -- not part of the product path.

Case-file I/O for ``oracle/_ref/cals_ref`` (the unmodified reference library behind
``oracle/ref_tool.cpp``) and a helper that runs it.  Only ``tests/``, ``bench.py``'s
cpu_baseline / ``--impl reference`` leg and ``__graft_entry__.smoke()`` may import this.

Input file, little endian::

    "CALSIN01"
    i64 n_modes, i64 modes[n_modes]
    i64 n_models, i64 max_iterations, f64 tol, i64 buffer_size,
    i64 flags (bit0 force_max_iter, bit1 always_evict_first), i64 threads,
    i64 algo (0 cp_cals, 1 cp_als loop, 2 jk_cp_cals, 3 jk_cp_als), i64 mttkrp_method (0..3, 3 = AUTO)
    per model: i64 rank, i64 jk_mode (-1 = regular), i64 jk_fiber
    f64 X[prod(modes)]                        column-major, mode 0 fastest
    per model: per mode f64 factor[I_n*rank]  column-major ; f64 lambda[rank]

Output file::

    "CALSOUT1"
    i64 n_models_out, f64 seconds (reference's own timer), f64 wall, i64 rep.iter, i64 rep.n_ktensors,
    i64 rep.ktensor_comp_sum, f64 X_norm
    per model: i64 rank, i64 iters, f64 approx_error, f64 fit_diff, i64 rows[n_modes],
               f64 lambda[rank], per mode f64 factor[rows*rank]
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "cals_ref")

ALGO_CALS, ALGO_ALS, ALGO_JK_CALS, ALGO_JK_ALS = 0, 1, 2, 3
METHOD_MTTKRP, METHOD_TWOSTEP0, METHOD_TWOSTEP1, METHOD_AUTO = 0, 1, 2, 3


@dataclass
class Model:
    """One CP model: factors[n] is (I_n, R) float64 (any memory order), lam is (R,)."""

    factors: list
    lam: np.ndarray | None = None
    jk_mode: int = -1
    jk_fiber: int = 0
    # filled on output
    iters: int = 0
    error: float = 0.0
    fit_diff: float = 0.0

    @property
    def rank(self) -> int:
        return int(self.factors[0].shape[1])


@dataclass
class RefResult:
    models: list
    seconds: float
    wall: float
    iters: int
    n_ktensors: int
    comp_sum: int
    x_norm: float
    stdout: str = ""


def ref_available() -> bool:
    return os.path.exists(REF_BIN)


def ref_binary(release: bool = False) -> str:
    """Path of the reference harness.  release=True: the build with the reference's Release flags
    (-O3 -ffast-math, oracle/build_ref.sh) at the highest x86-64 level the running CPU supports -- for TIMING only
    (bench.py's CPU arm); parity work uses the plain build."""
    if release:
        try:
            with open("/proc/cpuinfo") as f:
                flags = f.read()
        except OSError:
            flags = ""
        for lvl, need in ((4, "avx512f"), (3, "avx2")):
            path = os.path.join(HERE, "_ref", "cals_ref_rel%d" % lvl)
            if need in flags and os.path.exists(path):
                return path
    return REF_BIN


def write_case(path, X, models, *, max_iter, tol=1e-7, buffer_size=None, force_max_iter=False,
               always_evict_first=False, threads=1, algo=ALGO_CALS, mttkrp_method=METHOD_AUTO, nnls=False,
               line_search=False, ls_method=0, ls_interval=5, ls_step=0.0):
    modes = list(X.shape)
    if buffer_size is None:
        buffer_size = sum(m.rank for m in models)
    flags = ((1 if force_max_iter else 0) | (2 if always_evict_first else 0) | (4 if nnls else 0) |
             (8 if line_search else 0))
    with open(path, "wb") as f:
        f.write(b"CALSIN01")
        f.write(struct.pack("<q", len(modes)))
        f.write(struct.pack("<%dq" % len(modes), *modes))
        f.write(struct.pack("<qqdqqqqq", len(models), max_iter, tol, buffer_size, flags, threads, algo, mttkrp_method))
        f.write(struct.pack("<qqd", ls_method, ls_interval, ls_step))
        for m in models:
            f.write(struct.pack("<qqq", m.rank, m.jk_mode, m.jk_fiber))
        f.write(np.asfortranarray(X, dtype=np.float64).tobytes(order="F"))
        for m in models:
            for F in m.factors:
                f.write(np.asfortranarray(F, dtype=np.float64).tobytes(order="F"))
            lam = m.lam if m.lam is not None else np.ones(m.rank)
            f.write(np.asarray(lam, dtype=np.float64).tobytes())


def read_result(path, n_modes) -> RefResult:
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"CALSOUT1"
    off = 8
    n_out, seconds, wall, it, nk, cs, xn = struct.unpack_from("<qddqqqd", buf, off)
    off += struct.calcsize("<qddqqqd")
    models = []
    for _ in range(n_out):
        rank, iters, err, fit_diff = struct.unpack_from("<qqdd", buf, off)
        off += 32
        rows = struct.unpack_from("<%dq" % n_modes, buf, off)
        off += 8 * n_modes
        lam = np.frombuffer(buf, dtype=np.float64, count=rank, offset=off).copy()
        off += 8 * rank
        factors = []
        for r in rows:
            F = np.frombuffer(buf, dtype=np.float64, count=r * rank, offset=off).reshape((r, rank), order="F").copy()
            off += 8 * r * rank
            factors.append(F)
        models.append(Model(factors=factors, lam=lam, iters=iters, error=err, fit_diff=fit_diff))
    assert off == len(buf)
    return RefResult(models, seconds, wall, it, nk, cs, xn)


def run_reference(X, models, *, env_extra=None, timeout=1800, release=False, **kw) -> RefResult:
    """Run the unmodified reference (oracle/_ref/cals_ref) on a case and return its results."""
    if not ref_available():
        raise FileNotFoundError("oracle/_ref/cals_ref not built; run oracle/build_ref.sh")
    binary = ref_binary(release)
    env = dict(os.environ)
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference sets its own thread counts from the case file
    # (set_threads), but the BLAS pool is sized from the environment when the library loads: make both say `threads`.
    th = str(int(kw.get("threads", 1)))
    env["OMP_NUM_THREADS"] = th
    env["OPENBLAS_NUM_THREADS"] = th
    # The wheel OpenBLAS is a pthreads build nested under the reference's OpenMP loops: keep the waiters passive
    # (SURVEY.md section 8c "known threading hazard"); results are unaffected.
    env.setdefault("OMP_WAIT_POLICY", "passive")
    if env_extra:
        env.update(env_extra)
    with tempfile.TemporaryDirectory() as td:
        pin, pout = os.path.join(td, "case.in"), os.path.join(td, "case.out")
        write_case(pin, X, models, **kw)
        p = subprocess.run([binary, pin, pout], env=env, capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("cals_ref failed: %s\n%s" % (p.returncode, p.stderr[-2000:]))
        res = read_result(pout, X.ndim)
        res.stdout = p.stdout
    for src, dst in zip(models, res.models):
        dst.jk_mode, dst.jk_fiber = src.jk_mode, src.jk_fiber
    return res


def random_models(rng, modes, ranks, normalize=True):
    """Initial models as the reference's Ktensor::fill leaves them (src/ktensor.cpp:19-28, 85-99):
    uniform(-1,1) factors, then every column scaled to unit 2-norm and lambda = product of the norms."""
    out = []
    for r in ranks:
        fs = [rng.uniform(-1.0, 1.0, size=(i, r)) for i in modes]
        lam = np.ones(r)
        if normalize:
            for k, F in enumerate(fs):
                nrm = np.linalg.norm(F, axis=0)
                fs[k] = F / nrm
                lam = lam * nrm
        out.append(Model(factors=[np.asfortranarray(F) for F in fs], lam=lam))
    return out


def ktensor_to_tensor(factors, lam):
    """Dense reconstruction sum_r lam_r a_r o b_r o c_r ... (reference src/ktensor.cpp:30-64)."""
    R = factors[0].shape[1]
    N = len(factors)
    letters = "abcdefgh"[:N]
    expr = ",".join(l + "r" for l in letters) + ",r->" + letters
    return np.einsum(expr, *factors, np.asarray(lam).reshape(R), optimize=True)
