#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 CP-CALS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5] [--strong]

Metric (BASELINE.json): concurrent-ALS iterations per second summed over all models ("model-iterations/s").  The
default workload is BASELINE config 2 (the largest configuration the metric is quoted on that fits one GPU): synthetic
200x200x200 FP64 tensor, 200 concurrent models (ranks 1..20 x 10), buffer = sum of ranks, forced iteration count.
One "step" = one complete cals::cp_cals pass: all 200 models x 50 forced ALS iterations (the iteration count of the
reference's own experiments, src/experiments/experiments.cpp:62-63).

  value    : whole-job throughput with the tensor and the initial models already resident in HBM (cals_b200_rerun),
             timed with CUDA events on the engine's stream, max over ranks
  e2e      : same metric through the public API (cp_cals over HOST buffers in pinned memory): H2D of X and of the
             initial models, the loop, and D2H of every fitted model inside the timed region
  roofline : the MTTKRP kernel (dominant), algorithmic 2*nX*C flop per launch / CUDA-event time per launch, against the
             measured FP64 DMMA peak of this pool's B200 (profiles/fp64_peak_r01.json; MEASURED_PEAKS.json has no FP64
             entry)
  cpu_baseline : the UNMODIFIED reference (oracle/_ref) timed on the box's host cores on a bounded sample

N > 1 (torchrun): the model set is the unit of sharding -- X is replicated, no collective on the data path.  Default is
weak scaling (every rank fits its own full model set); --strong shards ONE model set over the ranks.
Other configs (--config 1, 3, 4) are measurement aids for DESIGN.md, not the headline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "concurrent-ALS iters/sec (all models)"
UNIT = "model-iterations/s"

# BASELINE.json configs.  als_iters: forced ALS iterations per model per step (config 2: 50, the protocol of the
# reference's own experiments, src/experiments/experiments.cpp:62-63); ref_iters: per reference-arm step (bounded sample).
CONFIGS = {
    1: dict(name="config 1: 100x100x100 tensor, 40 models (ranks 1..10 x4), buffer 220", modes=(100, 100, 100),
            ranks=[r for r in range(1, 11) for _ in range(4)], als_iters=50, ref_iters=20, jk=False),
    2: dict(name="config 2: 200x200x200 tensor, 200 models (ranks 1..20 x10), buffer 2100", modes=(200, 200, 200),
            ranks=[r for r in range(1, 21) for _ in range(10)], als_iters=50, ref_iters=2, jk=False),
    3: dict(name="config 3: jackknife on a 299x301x41 tensor (fluorescence_cancer_UD shape, synthetic), all 299 "
                 "leave-one-out sub-models of base models of ranks 3,5,7,9", modes=(299, 301, 41),
            ranks=[r for r in (3, 5, 7, 9) for _ in range(299)], als_iters=10, ref_iters=1, jk=True),
    4: dict(name="config 4: 80x80x80x80 tensor, 150 models (ranks 1..30 x5), buffer 2325", modes=(80, 80, 80, 80),
            ranks=[r for r in range(1, 31) for _ in range(5)], als_iters=4, ref_iters=1, jk=False),
    5: dict(name="config 5: 1000x1000x1000 tensor sliced along mode 2 over the GPUs, 50 models (ranks 1..50), buffer "
                 "1275, partial MTTKRPs combined over NVLink peer memory", modes=(1000, 1000, 1000),
            ranks=list(range(1, 51)), als_iters=2, ref_iters=1, jk=False, sliced=2),
}


def init_nccl(local_rank):
    """init_process_group + the first collective (which creates the communicator) with file descriptor 1 pointed at
    stderr: NCCL prints its version banner on STDOUT at that moment, and stdout is reserved for the one JSON line."""
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return dist


def load_pkg():
    from conftest import load_package
    return load_package()


def workload(cfg, seed):
    """Synthetic inputs as the reference driver makes them (src/examples/driver.cpp:133-153): X uniform(-1,1), models
    uniform(-1,1) then Ktensor::normalize().  X is the same on every rank (replicated); models differ with `seed`.
    Jackknife config: sub-model i of a base model is that model flagged "leave mode-0 sample i out" (the reference
    derives them from fitted base models; for throughput the starting values do not matter)."""
    modes = cfg["modes"]
    rng = np.random.default_rng(1234)
    X = np.asfortranarray(rng.uniform(-1.0, 1.0, size=modes))
    mrng = np.random.default_rng(1000 + seed)
    models, jk = [], []
    base = {}
    for idx, r in enumerate(cfg["ranks"]):
        if cfg["jk"] and r in base:
            fs = base[r]
        else:
            fs = []
            for i in modes:
                F = mrng.uniform(-1.0, 1.0, size=(i, r))
                fs.append(np.asfortranarray(F / np.linalg.norm(F, axis=0)))
            base[r] = fs
        models.append(fs)
        jk.append((0, idx % modes[0]) if cfg["jk"] else (-1, 0))
    return X, models, jk


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        busy = [s for s in sm if mx and s > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def fp64_peak():
    path = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["dmma_tflops_sustained"]), "measured (tools/fp64_peak.cu, profiles/fp64_peak_r01.json)"
    except Exception:
        return 37.0, "fallback (no profiles/fp64_peak_r01.json)"


def measured_hbm_peak():
    """HBM copy bandwidth of this pool's B200 in GB/s: MEASURED_PEAKS.json (driver-written), else the profiling
    recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6534.8


def run_reference_sample(X, models, jk, iters, threads):
    import caseio  # oracle/ (test infrastructure): allowed here as the CPU baseline / reference arm only
    ms = [caseio.Model(factors=fs, jk_mode=j[0], jk_fiber=j[1]) for fs, j in zip(models, jk)]
    res = caseio.run_reference(X, ms, max_iter=iters, force_max_iter=True, threads=threads,
                               buffer_size=sum(m.rank for m in ms))
    return res.seconds


def cpu_baseline(cfg, X, models, jk):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import caseio
    if not caseio.ref_available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    iters = cfg["ref_iters"]
    ncpu = os.cpu_count() or 1
    best = None
    for th in sorted({1, ncpu}):
        sec = run_reference_sample(X, models, jk, iters, th)
        v = len(models) * iters / sec
        if best is None or v > best[0]:
            best = (v, th, sec)
    return {"value": best[0], "unit": UNIT, "cores": best[1], "kind": "reference",
            "sample": "unmodified reference cp_cals (OpenBLAS), full model set of the workload, %d forced ALS "
                      "iterations, %.2f s of cp_cals time; better of 1 and %d threads (OMP_WAIT_POLICY=passive)"
                      % (iters, best[2], ncpu)}


def reference_arm(args, cfg, rank):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    X, models, jk = workload(cfg, 0)
    ncpu = os.cpu_count() or 1
    warm = args.warmup if args.warmup is not None else 1
    for _ in range(warm):
        run_reference_sample(X, models, jk, 1, ncpu)
    steps = args.steps if args.steps is not None else 3
    it = cfg["ref_iters"]
    secs = [run_reference_sample(X, models, jk, it, ncpu) for _ in range(steps)]
    t = float(np.mean(secs))
    v = len(models) * it / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 0, "steps": steps,
            "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "als_iters_per_step": it, "l2": "inputs larger than L2"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncpu, "kind": "reference",
                             "sample": "each step = %d forced ALS iterations of the full model set through the "
                                       "unmodified reference's cp_cals, %d threads" % (it, ncpu)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_sliced(args, cfg, rank, world, local_rank):
    """BASELINE config 5: ONE tensor sliced over the ranks (strong scaling), all models replicated."""
    import importlib

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        init_nccl(local_rank)
    pkg = load_pkg()
    dmod = importlib.import_module("cp_cals_b200.distributed")
    modes, s_mode, als_iters = cfg["modes"], cfg["sliced"], cfg["als_iters"]
    steps = args.steps if args.steps is not None else 5
    warmup = max(3, args.warmup if args.warmup is not None else 3)
    lo, hi = dmod.shard_slabs(modes[s_mode], world)[rank]
    shape = tuple((hi - lo) if n == s_mode else m for n, m in enumerate(modes))
    # the slab of a seeded uniform(-1,1) tensor: slab r is drawn from its own stream so ranks generate in parallel
    t = torch.empty(int(np.prod(shape)), dtype=torch.float64, pin_memory=True)
    slab = t.numpy().reshape(shape, order="F")
    rng = np.random.default_rng(5000 + rank)
    chunk = 1 << 24
    flat = t.numpy()
    for o in range(0, flat.size, chunk):
        flat[o:o + chunk] = rng.uniform(-1.0, 1.0, size=min(chunk, flat.size - o))
    mrng = np.random.default_rng(77)  # the same models on every rank
    models = []
    for r in cfg["ranks"]:
        fs = []
        for i in modes:
            F = mrng.uniform(-1.0, 1.0, size=(i, r))
            fs.append(np.asfortranarray(F / np.linalg.norm(F, axis=0)))
        models.append(fs)
    C, n_models = sum(cfg["ranks"]), len(models)
    params = pkg.CalsParams(max_iterations=als_iters, buffer_size=C, force_max_iter=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        v = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    eng = pkg.Engine(local_rank)
    es = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device("cuda", local_rank))

    def timed(fn, n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        ev0.record(es)
        out = [fn() for _ in range(n)]
        ev1.record(es)
        barrier()
        return ev0.elapsed_time(ev1), time.perf_counter() - t0, out

    def e2e_step():
        kts = [pkg.Ktensor(list(fs)) for fs in models]
        return dmod.cp_cals_sliced(slab, modes, s_mode, kts, params, engine=eng, device=local_rank), kts

    e2e_step()  # allocations, IPC mapping, upload (untimed)
    for _ in range(warmup):
        eng.rerun()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev_ms, wall, reps = timed(eng.rerun, steps)
    clocks = sampler.stop()
    ev_ms, wall = max_over_ranks(ev_ms), max_over_ranks(wall)
    value = n_models * als_iters * steps / (ev_ms * 1e-3)
    launches = sum(r.kernel_launches for r in reps)

    eng.set_timing(1)
    rep = eng.rerun()
    eng.set_timing(0)
    # tensor-sized contractions of the slab: one per mode, or -- slab cut along mode 1 or 2, pair node -- two per iteration
    flops_per_launch = 2.0 * float(np.prod(shape)) * C
    tensor_launches = rep.tensor_flops / flops_per_launch
    mt_ms = max_over_ranks(rep.mttkrp_ms - rep.pair_leaf_ms) / tensor_launches
    peak, peak_src = fp64_peak()
    ach = flops_per_launch / (mt_ms * 1e-3) / 1e12
    # NVLink volume of the exchange: every GPU pulls the partials of the other W-1 GPUs (full-size for the modes that
    # are not sliced, the foreign row blocks for the sliced one)
    ld = [(m + 1) // 2 * 2 for m in modes]
    pulled = [8.0 * C * ((modes[n] - (hi - lo)) if n == s_mode else (world - 1) * ld[n]) for n in range(len(modes))]
    xch_ms = max_over_ranks(rep.exchange_ms) / rep.mttkrp_launches
    exchange = {"kernel": "exchange_sum_kernel (barrier wait + peer-memory pulls + rank-ordered sum)",
                "ms_per_launch": xch_ms, "nvlink_bytes_pulled_per_gpu_per_launch": float(np.mean(pulled)),
                "achieved_gbs_incl_barrier_wait": float(np.mean(pulled)) / (xch_ms * 1e-3) / 1e9 if xch_ms > 0 else None,
                "peer_copy_peak_gbs": 770.0, "share_of_mttkrp_window": rep.exchange_ms / rep.mttkrp_ms
                if rep.mttkrp_ms > 0 else None} if world > 1 else None
    roofline = {"bound": "tensor", "kernel": "mttkrp_dmma_kernel (+ reduce + NVLink exchange_sum_kernel)" +
                          (" and pair_gemm_kernel" if rep.tree else ""),
                "pair_node": {"contractions_per_als_iteration": tensor_launches / rep.iter,
                              "leaf_ms_per_launch_incl_exchange": max_over_ranks(rep.pair_leaf_ms) / (2 * rep.iter),
                              "pair_gemm_ms_per_launch": max_over_ranks(rep.pair_gemm_ms) / rep.iter} if rep.tree else None,
                "exchange": exchange,
                "achieved": ach, "peak": peak, "unit": "TFLOP/s per GPU", "frac": ach / peak, "traffic": None,
                "peak_source": peak_src, "ms_per_launch": mt_ms, "flops_per_launch_per_gpu": flops_per_launch,
                "mttkrp_share_of_step": rep.mttkrp_ms / (rep.mttkrp_ms + rep.update_ms)}

    e2e_steps = 2
    e2e_ms, e2e_wall, outs = timed(e2e_step, e2e_steps)
    e2e_t = max(max_over_ranks(e2e_ms) * 1e-3, max_over_ranks(e2e_wall))
    kts = outs[-1][1]
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ev_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s, %d forced ALS iterations per model per step" % (cfg["name"], als_iters),
                       "models": n_models, "sum_ranks": C, "slab_per_gpu": list(shape),
                       "parallelism": "tensor sliced along mode %d over %d GPU(s), models replicated, one peer-memory "
                                      "exchange per mode" % (s_mode, world),
                       "l2": "inputs larger than L2 (%.1f GB of tensor per GPU)" % (np.prod(shape) * 16 / 1e9)},
            "wall_ms_per_step": wall / steps * 1e3,
            "timing": "CUDA events on the engine's stream around the K steps, max over ranks",
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "e2e": {"value": n_models * als_iters * e2e_steps / e2e_t, "unit": UNIT,
                    "h2d_bytes_per_step": int(slab.nbytes + sum(F.nbytes for fs in models for F in fs)),
                    "d2h_bytes_per_step": int(sum(F.nbytes for fs in models for F in fs) + 8 * C),
                    "steps": e2e_steps, "ms_per_step": e2e_t / e2e_steps * 1e3,
                    "mean_fit": float(np.mean([k.fit for k in kts]))},
            "cpu_baseline": {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                             "sample": "not run: the reference materialises a 10 GB Khatri-Rao workspace for this "
                                       "tensor (SURVEY 8d); config 2 carries the CPU baseline"},
        }
        print(json.dumps(line), flush=True)
    dmod.release_sliced_engine(eng)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--strong", action="store_true", help="N > 1: shard ONE model set over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard-of", type=int, default=0, help="tuning aid on ONE GPU: run only the first shard of an "
                    "N-way --strong split (what each GPU of an N-GPU job gets)")
    ap.add_argument("--no-pair-node", action="store_true", help="measurement aid: one full MTTKRP per mode (3-mode "
                    "tensors otherwise share one contraction between modes 1 and 2, csrc/pairnode.cuh)")
    ap.add_argument("--nnls", action="store_true", help="measurement aid: update_method = NNLS instead of the Cholesky solve")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if cfg.get("sliced") is not None:
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "config 5 is not run through the CPU reference "
                                  "(10 GB Khatri-Rao workspace); use --config 2"}), flush=True)
            return
        reference_arm(args, cfg, rank)
        return
    if cfg.get("sliced") is not None:
        run_sliced(args, cfg, rank, world, local_rank)
        return

    steps = args.steps if args.steps is not None else 10
    warmup = max(3, args.warmup if args.warmup is not None else 3)
    als_iters = cfg["als_iters"]

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        dist = init_nccl(local_rank)

    pkg = load_pkg()
    import importlib
    dmod = importlib.import_module("cp_cals_b200.distributed")
    X, models, jk = workload(cfg, 0 if args.strong else rank)
    total_models = len(models) if args.strong else world * len(models)
    if args.strong and world > 1:
        mine = dmod.shard_models([fs[0].shape[1] for fs in models], world)[rank]
        models, jk = [models[i] for i in mine], [jk[i] for i in mine]
    if args.shard_of > 1 and world == 1:
        mine = dmod.shard_models([fs[0].shape[1] for fs in models], args.shard_of)[0]
        models, jk = [models[i] for i in mine], [jk[i] for i in mine]
        total_models = len(models)
    nX = X.size
    modes = X.shape
    C = sum(fs[0].shape[1] for fs in models)
    n_models = len(models)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = pkg.Engine(local_rank)
    eng.set_tensor(X)
    eng.configure(C, als_iters, 1e-7, force_max_iter=True, nnls=args.nnls)
    eng.set_pair_node(not args.no_pair_node)
    eng.clear_models()
    for fs, j in zip(models, jk):
        eng.enqueue(fs, j[0], j[1])
    eng.run()  # uploads + first pass (untimed)
    # CUDA events are recorded on the stream the engine launches on (torch's current stream would see nothing)
    es = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device("cuda", local_rank))

    def timed(fn, n):
        """n calls of fn bracketed by barrier + synchronize; returns (device ms between events, wall seconds)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        ev0.record(es)
        out = [fn() for _ in range(n)]
        ev1.record(es)
        barrier()
        t1 = time.perf_counter()
        return ev0.elapsed_time(ev1), t1 - t0, out

    # ---------------- value: inputs resident ----------------
    for _ in range(warmup):
        eng.rerun()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev_ms, wall, reps = timed(eng.rerun, steps)
    clocks = sampler.stop()
    launches = sum(r.kernel_launches for r in reps)
    assert all(r.iter == als_iters and r.n_ktensors == n_models for r in reps)
    ev_ms = max_over_ranks(ev_ms)       # CUDA events on the engine stream around the K steps, max over ranks
    wall = max_over_ranks(wall)         # host clock around the same region (reported beside it)
    value = total_models * als_iters * steps / (ev_ms * 1e-3)

    # ---------------- roofline of the dominant kernel (extra passes with per-kernel CUDA events) ----------------
    eng.set_timing(1)
    mt_ms, mt_launches, up_ms, gemm_ms, leaf_ms, tensor_flops, tree = 0.0, 0, 0.0, 0.0, 0.0, 0.0, False
    for _ in range(2):
        rep = eng.rerun()
        mt_ms += rep.mttkrp_ms
        up_ms += rep.update_ms
        mt_launches += rep.mttkrp_launches
        gemm_ms += rep.pair_gemm_ms
        leaf_ms += rep.pair_leaf_ms
        tensor_flops += rep.tensor_flops
        tree = tree or bool(rep.tree)
    eng.set_timing(0)
    # The tensor-core kernels: one full MTTKRP per mode, or -- 3-mode tensors, pair node (csrc/pairnode.cuh) -- the MTTKRP
    # of mode 0 and the shared contraction T = X_(0)^T A_0; each launch is one tensor-sized contraction of 2*nX*C flop.
    flops_per_launch = 2.0 * nX * C
    tensor_ms = mt_ms - leaf_ms
    tensor_launches = tensor_flops / flops_per_launch
    ach = tensor_flops / (tensor_ms * 1e-3) / 1e12
    peak, peak_src = fp64_peak()
    traffic = None
    if args.config == 2 and world == 1:
        try:
            with open(os.path.join(ROOT, "profiles", "mttkrp_traffic_r01.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "tensor",
                "kernel": "mttkrp_dmma_kernel (timed together with its mttkrp_reduce_kernel)" +
                          (" and the pair-node contraction (pair_gemm_kernel)" if tree and len(modes) == 3 else ""),
                "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "ms_per_launch": tensor_ms / tensor_launches, "flops_per_launch": flops_per_launch,
                "launches_per_als_iteration": tensor_launches / (mt_launches / len(modes)),
                "mttkrp_ms_per_mode": mt_ms / mt_launches,
                "mttkrp_share_of_step": mt_ms / (mt_ms + up_ms) if mt_ms + up_ms > 0 else None}
    if tree:
        iters = mt_launches / len(modes)
        # 3 modes: one pair node (modes 1, 2) next to the full MTTKRP of mode 0; 4 modes: two pair nodes (0, 1), (2, 3)
        node_rows = [modes[1] * modes[2]] if len(modes) == 3 else [modes[0] * modes[1], modes[2] * modes[3]]
        full_modes = len(modes) - 2 * len(node_rows)
        leaf_launches = 2 * len(node_rows) * iters
        leaf_bytes = sum(2 * 8.0 * r * C for r in node_rows) * iters  # every leaf streams its node's T once
        roofline["pair_node"] = {
            "what": "two modes take their MTTKRP from one shared contraction T (csrc/pairnode.cuh): %d tensor-sized "
                    "contractions per ALS iteration instead of %d; the leaf kernels stream T from HBM"
                    % (full_modes + len(node_rows), len(modes)),
            "mttkrp_dmma_ms_per_launch": (tensor_ms - gemm_ms) / (iters * full_modes) if full_modes else None,
            "pair_contraction_ms_per_launch": gemm_ms / (iters * len(node_rows)),
            "pair_contraction_tflops": flops_per_launch / (gemm_ms / (iters * len(node_rows)) * 1e-3) / 1e12,
            "leaf_ms_per_launch": leaf_ms / leaf_launches,
            "leaf_bytes_per_launch": leaf_bytes / leaf_launches,
            "leaf_achieved_gbs": leaf_bytes / (leaf_ms * 1e-3) / 1e9,
            "leaf_peak_gbs": measured_hbm_peak(),
            "algorithmic_mttkrp_tflops": len(modes) * flops_per_launch * iters / (mt_ms * 1e-3) / 1e12,
        }
        if args.config == 2 and world == 1:  # DRAM traffic of the pair contraction from its ncu --set full capture
            try:
                with open(os.path.join(ROOT, "profiles", "pair_node_full_r01.json")) as f:
                    for l in json.load(f)["launches"]:
                        if l["kernel"].startswith("pair_gemm_kernel"):
                            unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
                            roofline["pair_node"]["pair_contraction_traffic"] = sum(
                                l[k]["value"] * unit.get(l[k]["unit"], 1.0)
                                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                            break
            except Exception:
                pass

    # ---------------- e2e: public API, host buffers in pinned memory ----------------
    def pinned_copy(a):
        t = torch.empty(a.size, dtype=torch.float64, pin_memory=True)
        v = t.numpy().reshape(a.shape, order="F")
        v[...] = a
        return t, v

    keep = []
    tX, Xp = pinned_copy(X)
    keep.append(tX)
    pinned_models = []
    for fs in models:
        row = []
        for F in fs:
            t, v = pinned_copy(F)
            keep.append(t)
            row.append(v)
        pinned_models.append(row)
    params = pkg.CalsParams(max_iterations=als_iters, buffer_size=C, force_max_iter=True,
                            update_method="nnls" if args.nnls else "unconstrained",
                            mttkrp_method="mttkrp" if args.no_pair_node else "auto")
    h2d = X.nbytes + sum(F.nbytes for fs in models for F in fs)
    d2h = sum(F.nbytes for fs in models for F in fs) + 8 * C + n_models * 40

    def e2e_step():
        kts = [pkg.Ktensor(list(fs), None, j[0], j[1]) for fs, j in zip(pinned_models, jk)]
        rep = pkg.cp_cals(Xp, kts, params, engine=eng)
        return rep, kts

    e2e_steps = max(2, min(steps, 5))
    e2e_step()
    e2e_ms, e2e_wall, outs = timed(e2e_step, e2e_steps)
    rep, kts = outs[-1]
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_wall = max_over_ranks(e2e_wall)
    # the end-to-end step includes host work (queue packing, result unpacking) that no CUDA event sees once the
    # stream is idle, so the slower of the two clocks is the honest one
    e2e_t = max(e2e_ms * 1e-3, e2e_wall)
    e2e_value = total_models * als_iters * e2e_steps / e2e_t
    fit_checksum = float(np.mean([k.fit for k in kts]))

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cb = cpu_baseline(cfg, X, models, jk)
            except Exception as e:  # the baseline is reported, never allowed to break the bench line
                cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %s" % e}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ev_ms / steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s, %d forced ALS iterations per model per step%s"
                                   % (cfg["name"], als_iters, ", NNLS update" if args.nnls else ""),
                       "models_per_gpu": n_models, "sum_ranks_per_gpu": C, "als_iters_per_step": als_iters,
                       "parallelism": "model set sharded over %d GPU(s) (%s), tensor replicated, no data-path "
                                      "collective" % (world, "one set split" if args.strong else "one full set per GPU"),
                       "l2": ("inputs larger than L2: every ALS iteration streams both device copies of the tensor "
                              "(2 x %.0f MB) plus the partial-tile workspace; no explicit flush" % (nX * 8 / 1e6))
                       if 2 * nX * 8 > 126e6 else
                       ("working set (2 x %.0f MB tensor copies) fits the 126 MB L2 -- that is this workload's "
                        "steady state inside cp_cals, not a cache artefact of the timing loop" % (nX * 8 / 1e6))},
            "wall_ms_per_step": wall / steps * 1e3,
            "timing": "CUDA events on the engine's stream around the K steps (barrier + synchronize on both sides), "
                      "max over ranks",
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": e2e_t / e2e_steps * 1e3, "mean_fit": fit_checksum},
        }
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
