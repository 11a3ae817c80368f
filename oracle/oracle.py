"""TEST INFRASTRUCTURE -- ctypes binding of oracle/libcals_oracle.so (oracle/cals_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcals_oracle.so")


class _Model(C.Structure):
    _fields_ = [("rank", C.c_int64), ("jk_mode", C.c_int64), ("jk_fiber", C.c_int64),
                ("factors", C.POINTER(C.c_double)), ("lam", C.POINTER(C.c_double)),
                ("iters", C.c_int64), ("error", C.c_double), ("fit", C.c_double), ("old_fit", C.c_double),
                ("chol_fail", C.c_int64), ("active", C.POINTER(C.c_ubyte))]


class _Report(C.Structure):
    _fields_ = [("iter", C.c_int64), ("n_ktensors", C.c_int64), ("ktensor_comp_sum", C.c_int64),
                ("x_norm", C.c_double), ("ls_performed", C.c_int64), ("ls_failed", C.c_int64)]


class _Ls(C.Structure):
    _fields_ = [("enabled", C.c_int), ("method", C.c_int), ("interval", C.c_int), ("step", C.c_double)]


_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.cals_oracle_norm.restype = C.c_double
        _lib.cals_oracle_norm.argtypes = [C.c_int64, C.c_void_p]
        _lib.cals_oracle_jk_norms.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.cals_oracle_mttkrp.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int64]
        _lib.cals_oracle_cp_cals.restype = C.c_int
        _lib.cals_oracle_cp_cals.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(_Model), C.c_int64,
                                             C.c_double, C.c_int64, C.c_int, C.POINTER(_Report)]
        _lib.cals_oracle_cp_cals_ls.restype = C.c_int
        _lib.cals_oracle_cp_cals_ls.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(_Model), C.c_int64,
                                                C.c_double, C.c_int64, C.c_int, C.POINTER(_Ls), C.POINTER(_Report)]
        _lib.cals_oracle_denormalize_normalize.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    return _lib


def _modes_arr(modes):
    return np.asarray(list(modes), dtype=np.int64)


def _xf(X):
    return np.asfortranarray(X, dtype=np.float64)


def norm(X) -> float:
    Xf = _xf(X)
    return lib().cals_oracle_norm(Xf.size, Xf.ctypes.data)


def jk_norms(X):
    Xf = _xf(X)
    m = _modes_arr(Xf.shape)
    out = np.zeros(Xf.shape[0])
    lib().cals_oracle_jk_norms(Xf.ndim, m.ctypes.data, Xf.ctypes.data, out.ctypes.data)
    return out


def mttkrp(X, factors, mode):
    """factors[k]: (I_k, C) arrays (entry `mode` ignored).  Returns G (I_mode, C), Fortran order."""
    Xf = _xf(X)
    m = _modes_arr(Xf.shape)
    fs = [np.asfortranarray(F, dtype=np.float64) for F in factors]
    Cc = fs[(mode + 1) % len(fs)].shape[1]
    ptrs = (C.c_void_p * len(fs))(*[F.ctypes.data for F in fs])
    ld = np.asarray([F.shape[0] for F in fs], dtype=np.int64)
    G = np.zeros((Xf.shape[mode], Cc), order="F")
    lib().cals_oracle_mttkrp(Xf.ndim, m.ctypes.data, Xf.ctypes.data, mode, Cc, ptrs, ld.ctypes.data, G.ctypes.data,
                             G.shape[0])
    return G


class OracleResult:
    def __init__(self):
        self.models = []
        self.iters = 0
        self.n_ktensors = 0
        self.comp_sum = 0
        self.x_norm = 0.0


def cp_cals(X, models, *, max_iter, tol=1e-7, buffer_size=None, force_max_iter=False, always_evict_first=False,
            nnls=False, line_search=False, ls_method=0, ls_interval=5, ls_step=0.0):
    """Run the oracle's cp_cals.  `models` is a list of caseio.Model (inputs untouched); returns OracleResult whose
    .models are new caseio.Model objects with factors/lam/iters/error/fit/old_fit filled."""
    from caseio import Model  # same directory

    Xf = _xf(X)
    m = _modes_arr(Xf.shape)
    if buffer_size is None:
        buffer_size = sum(mm.rank for mm in models)
    arr = (_Model * len(models))()
    keep = []
    for i, mm in enumerate(models):
        flat = np.concatenate([np.asfortranarray(F, dtype=np.float64).ravel(order="F") for F in mm.factors])
        lam = np.zeros(mm.rank)
        keep.append((flat, lam))
        arr[i].rank = mm.rank
        arr[i].jk_mode = mm.jk_mode
        arr[i].jk_fiber = mm.jk_fiber
        arr[i].factors = flat.ctypes.data_as(C.POINTER(C.c_double))
        arr[i].lam = lam.ctypes.data_as(C.POINTER(C.c_double))
    rep = _Report()
    flags = (1 if force_max_iter else 0) | (2 if always_evict_first else 0) | (4 if nnls else 0)
    ls = _Ls(1 if line_search else 0, ls_method, ls_interval, ls_step)
    rc = lib().cals_oracle_cp_cals_ls(Xf.ndim, m.ctypes.data, Xf.ctypes.data, len(models), arr, max_iter, tol,
                                      buffer_size, flags, C.byref(ls), C.byref(rep))
    if rc != 0:
        raise RuntimeError("oracle cp_cals failed rc=%d" % rc)
    res = OracleResult()
    res.iters, res.n_ktensors, res.comp_sum, res.x_norm = rep.iter, rep.n_ktensors, rep.ktensor_comp_sum, rep.x_norm
    res.ls_performed, res.ls_failed = rep.ls_performed, rep.ls_failed
    for i, mm in enumerate(models):
        flat, lam = keep[i]
        fs, off = [], 0
        for I in Xf.shape:
            fs.append(flat[off:off + I * mm.rank].reshape((I, mm.rank), order="F").copy(order="F"))
            off += I * mm.rank
        out = Model(factors=fs, lam=lam.copy(), jk_mode=mm.jk_mode, jk_fiber=mm.jk_fiber, iters=arr[i].iters,
                    error=arr[i].error)
        out.fit, out.old_fit, out.chol_fail = arr[i].fit, arr[i].old_fit, arr[i].chol_fail
        out.fit_diff = abs(out.old_fit - out.fit)
        res.models.append(out)
    return res
