"""Multi-GPU host logic of the B200 CP-CALS path: one process per GPU (torch.distributed), the tensor replicated, the
model set -- or, for the jackknife, the leave-one-out sub-models -- sharded across ranks.

Models are independent given X, so there is NO collective on the data path: every rank runs the single-GPU loop
(cp_cals over its own Engine) on its shard; one all-gather of the packed results at the end gives every rank all
fitted models (NCCL over NVLink on a GPU box; gloo in the CPU tests, which exercise exactly this file with a stand-in
fit function).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


def shard_models(ranks: Sequence[int], n_parts: int) -> List[List[int]]:
    """Deterministic split of a FIFO model list over n_parts shards: largest rank first (ties: queue order), every model
    goes to the shard with the smallest sum of ranks so far (ties: lowest shard index); inside a shard the queue order is
    kept.  Largest-first keeps the shards within one small rank of each other -- the per-GPU column counts set the MTTKRP
    tile fill, and the slowest shard sets the time of the job.  Same rule as the C++ host layer
    (cals::detail::shard_models, cp-cals_b200/host/cals.cpp)."""
    parts: List[List[int]] = [[] for _ in range(max(1, n_parts))]
    load = [0] * len(parts)
    for i in sorted(range(len(ranks)), key=lambda i: (-int(ranks[i]), i)):
        best = min(range(len(parts)), key=lambda p: (load[p], p))
        parts[best].append(i)
        load[best] += int(ranks[i])
    for p in parts:
        p.sort()
    return parts


def shard_slabs(extent: int, n_parts: int, align: int = 2) -> List[tuple]:
    """Cut [0, extent) into n_parts contiguous slabs whose inner boundaries are multiples of `align` (the engine reads
    factor rows through TMA descriptors whose base must be 16-byte aligned => even row offsets)."""
    cuts = [0]
    for p in range(1, n_parts):
        c = int(p * extent / n_parts / align + 0.5) * align
        cuts.append(min(max(c, cuts[-1]), extent))
    cuts.append(extent)
    return [(cuts[p], cuts[p + 1]) for p in range(n_parts)]


# ---------------------------------------------------------------------------------------------------------------------
_STATS = 5  # iters, error, fit, old_fit, chol_info


def pack_models(models) -> np.ndarray:
    """Flatten fitted models (objects with .factors .lam .iters .error .fit .old_fit .chol_info) into one float64 vector."""
    chunks = []
    for m in models:
        chunks.append(np.array([m.iters, m.error, m.fit, m.old_fit, getattr(m, "chol_info", 0)], dtype=np.float64))
        chunks.append(np.asarray(m.lam, dtype=np.float64).ravel())
        for F in m.factors:
            chunks.append(np.asarray(F, dtype=np.float64).ravel(order="F"))
    return np.concatenate(chunks) if chunks else np.zeros(0)


def unpack_models(buf: np.ndarray, models) -> None:
    """Inverse of pack_models: overwrite `models` (whose factor shapes are known) in place."""
    off = 0
    for m in models:
        st = buf[off:off + _STATS]
        off += _STATS
        m.iters, m.error, m.fit, m.old_fit = int(st[0]), float(st[1]), float(st[2]), float(st[3])
        m.chol_info = int(st[4])
        R = m.factors[0].shape[1]
        m.lam = buf[off:off + R].copy()
        off += R
        for n, F in enumerate(m.factors):
            cnt = F.shape[0] * R
            m.factors[n] = np.asfortranarray(buf[off:off + cnt].reshape(F.shape, order="F").copy())
            off += cnt
    assert off == buf.size, (off, buf.size)


def packed_size(models) -> int:
    return sum(_STATS + m.factors[0].shape[1] * (1 + sum(F.shape[0] for F in m.factors)) for m in models)


def all_gather_packed(local: np.ndarray, sizes: Sequence[int], group=None, device=None) -> List[np.ndarray]:
    """All-gather of per-rank float64 vectors of known (different) sizes.  Buffers are padded to the largest size so one
    collective suffices; on NCCL the staging tensors live on `device`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    cap = max(max(sizes), 1)
    send = torch.zeros(cap, dtype=torch.float64, device=device)
    if local.size:
        send[:local.size] = torch.from_numpy(local).to(send.device)
    recv = [torch.empty(cap, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return [r[:n].cpu().numpy() for r, n in zip(recv, sizes)]


def cp_cals_sharded(X: np.ndarray, ktensors, params, *, fit_fn: Optional[Callable] = None, group=None,
                    gather_device=None, gather: bool = True, **kw):
    """cals::cp_cals with the model set sharded over the ranks of `group` (default: the world).

    Every rank passes the SAME `ktensors` list (same initial values).  Rank r fits shard r of shard_models() on its own
    GPU; with gather=True every rank ends up with all fitted models, as if it had fitted them itself.  Returns
    (report_of_this_rank, indices_of_this_rank's_shard).  `fit_fn(X, kts, params, **kw)` defaults to this package's
    cp_cals (the CUDA path, no fallback; pass device=<CUDA ordinal> through **kw); the CPU tests pass a stand-in to
    exercise the host logic over gloo.  `gather_device`: torch device of the all-gather staging buffers (a CUDA device
    for NCCL, None for gloo).
    """
    import torch.distributed as dist
    if fit_fn is None:
        from . import cp_cals as fit_fn  # noqa: N813
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    parts = shard_models([k.rank for k in ktensors], world)
    mine = [ktensors[i] for i in parts[rank]]
    rep = fit_fn(X, mine, params, **kw) if mine else None
    if gather and world > 1:
        sizes = [packed_size([ktensors[i] for i in p]) for p in parts]
        bufs = all_gather_packed(pack_models(mine), sizes, group=group, device=gather_device)
        for r, p in enumerate(parts):
            if r != rank:
                unpack_models(bufs[r], [ktensors[i] for i in p])
    return rep, parts[rank]


def jk_cp_cals_sharded(X: np.ndarray, ktensors, params, *, fit_fn: Optional[Callable] = None, group=None,
                       gather_device=None, **kw):
    """cals::jk_cp_cals (reference src/cals.cpp:397-446, without the LSAP column matching) with the leave-one-out
    sub-models -- not the base models -- sharded across ranks.  Returns (report_of_this_rank, results[b][i])."""
    from . import generate_jk_ktensors
    bases = [kt.copy().denormalize().normalize() for kt in ktensors]
    groups = [generate_jk_ktensors(b) for b in bases]
    flat = [m for g in groups for m in g]
    rep, _ = cp_cals_sharded(X, flat, params, fit_fn=fit_fn, group=group, gather_device=gather_device, gather=True,
                             **kw)
    for m in flat:
        m.set_jk_fiber(0.0)
        m.denormalize()
        m.normalize()
        m.set_jk_fiber(float("nan"))
    return rep, groups


# ---------------------------------------------------------------------------------------------------------------------
def exchange_capacity(modes: Sequence[int], buffer_cols: int) -> int:
    """Doubles per exchange buffer: the largest G matrix (pitch = extent rounded up to even) x buffer columns."""
    return max((m + 1) // 2 * 2 for m in modes) * int(buffer_cols)


def cp_cals_sliced(slab: np.ndarray, modes: Sequence[int], slice_mode: int, ktensors, params, *, engine=None,
                   device: int = 0, group=None, timing: int = 0):
    """cals::cp_cals for a tensor too large for one GPU (BASELINE config 5): X is cut into slabs along `slice_mode`
    (shard_slabs), rank r holds `slab` = X[.., cuts[r]:cuts[r+1], ..] and ALL models; every MTTKRP is combined across
    GPUs by the engine's own exchange kernel over NVLink peer memory (csrc/comm.cuh), the per-model updates run
    replicated.  Every rank passes the same `ktensors` and ends up with the same fitted models (bit-identical).
    torch.distributed is used only to pass the CUDA IPC handles and the slab norms around."""
    import torch
    import torch.distributed as dist
    from . import LS_METHODS, CalsReport, Engine, _check_params
    _check_params(params)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    cuts = [0] + [hi for _, hi in shard_slabs(modes[slice_mode], world)]
    own = engine is None
    eng = engine or Engine(device)
    try:
        buffer_cols = min(params.buffer_size, sum(k.rank for k in ktensors))
        cap = exchange_capacity(modes, buffer_cols)
        have = getattr(eng, "_comm", None)
        if not (have and have[0] == rank and have[1] == world and have[2] >= cap):  # else: reuse the mapped blocks
            if have and world > 1:
                eng.comm_disconnect()
                dist.barrier(group=group)
            handle = eng.comm_alloc(rank, world, cap)
            if world > 1:
                handles = [None] * world
                dist.all_gather_object(handles, handle, group=group)
                eng.comm_connect_ipc(handles)
            eng._comm = (rank, world, cap)
        eng.set_tensor_slab(modes, slice_mode, cuts, slab)
        sq = torch.tensor([eng.tensor_norm() ** 2], dtype=torch.float64,
                          device=torch.device("cuda", device) if world > 1 and dist.get_backend(group) == "nccl"
                          else None)
        if world > 1:
            parts = [torch.zeros_like(sq) for _ in range(world)]
            dist.all_gather(parts, sq, group=group)
            total = sum(float(p.item()) for p in parts)  # rank order: the same value on every rank
        else:
            total = float(sq.item())
        eng.set_tensor_norm(total ** 0.5)
        # same option forwarding as cp_cals and as the C++ run_sliced (host/cals.cpp); combinations the sliced engine
        # does not implement (error-checking line search, line search with NNLS) are refused by the engine itself
        nnls = params.update_method == "nnls"
        eng.configure(buffer_cols, params.max_iterations, params.tol, params.force_max_iter, params.always_evict_first,
                      nnls)
        eng.set_line_search(params.line_search, LS_METHODS.get(params.line_search_method, 0),
                            params.line_search_interval, params.line_search_step)
        eng.set_timing(timing)
        eng.set_pair_node(str(params.mttkrp_method).lower() != "mttkrp")
        eng.clear_models()
        for kt in ktensors:
            eng.enqueue(kt.factors, kt.jk_mode, kt.jk_fiber)
        if nnls:  # warm-start active sets travel with the Ktensor (reference include/ktensor.h:36)
            for i, kt in enumerate(ktensors):
                if kt.active_set is not None:
                    eng.set_active_set(i, kt.active_set)
        if world > 1:
            dist.barrier(group=group)  # every peer's exchange block is mapped before anyone starts to signal
        rep = eng.run()
        for i, (kt, (fs, lam, st)) in enumerate(zip(ktensors, eng.fetch_all())):
            kt.factors, kt.lam = fs, lam
            kt.iters, kt.error, kt.fit, kt.old_fit, kt.chol_info = st.iters, st.error, st.fit, st.old_fit, st.chol_info
            if nnls:
                kt.active_set = eng.fetch_active_set(i)
        if world > 1:
            dist.barrier(group=group)  # nobody frees its exchange block while a peer may still read it
        return CalsReport(n_modes=len(modes), modes=tuple(modes), X_norm=rep.x_norm, iter=rep.iter,
                          max_iter=params.max_iterations, buffer_size=params.buffer_size, n_ktensors=rep.n_ktensors,
                          ktensor_comp_sum=rep.ktensor_comp_sum, tol=params.tol, total_time=rep.total_time,
                          device_ms=rep.device_ms, mttkrp_ms=rep.mttkrp_ms, update_ms=rep.update_ms,
                          mttkrp_launches=rep.mttkrp_launches, kernel_launches=rep.kernel_launches)
    finally:
        if own:
            release_sliced_engine(eng, group)


def release_sliced_engine(eng, group=None):
    """Tear-down order CUDA IPC asks for: every rank unmaps its peers' exchange blocks, all ranks synchronise, and
    only then does anybody free its own block."""
    import torch.distributed as dist
    if getattr(eng, "_comm", None) and dist.is_initialized() and dist.get_world_size(group) > 1:
        try:
            eng.comm_disconnect()
            dist.barrier(group=group)
        except Exception:
            pass
    eng.close()
