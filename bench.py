#!/usr/bin/env python
"""Benchmark of the local-BA hot path (BASELINE.json metric: LM iterations/s and observations/s through
residual + Jacobian + Schur at 1/2/4/8 B200, against the HBM roofline, with the CPU path timed beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--impl b200|reference]

Default workload = BASELINE config 4, the one the metric is quoted on: ONE window of 1 000 cameras /
1 M points / 8 M observations.  With N > 1 (torchrun, one rank per GPU) the SAME window is sharded by map
point over the ranks (cameras replicated) and the reduced camera system + costs are NCCL all-reduced inside
the CUDA library every LM iteration: strong scaling.  N = 1 runs the same window on one GPU.

A *step* is one complete solve of the window: `max_iter` (10) LM iterations, each = linearise + Schur ->
(all-reduce) -> reduced-camera solve -> back-substitution + candidate cost -> accept/reject, plus state
init and outlier flags.  Every step starts from the same initial state.

  value    = LM iterations/s, inputs resident in HBM (CUDA events around each step on the launching
             stream, L2 flushed between steps, max over ranks).
  e2e      = the same through the host-buffer C ABI (`ba_solve_local_host`): pinned host inputs -> H2D ->
             solve -> D2H of poses, points, flags and the LM report inside the timed region.
  roofline = the linearise (residual + Jacobian + Schur) kernel: algorithmic bytes per launch / its live
             CUDA-event duration, against the measured HBM peak; `roofline.fp64` the same in flops.
  parity   = N > 1: the sharded solve against the single-GPU solve of the same window run in the same job
             (cost history, accept sequence, poses, points, flags; poses identical across ranks);
             N = 1: full-size properties against the oracle's objective / flag restatement.
  cpu_baseline = oracle/lm_oracle.py (same LM, numpy float64) on a bounded sample of the window.
  configs  = (N = 1 only) BASELINE configs 1-3 as sub-objects, each with value / e2e / roofline / parity.
             For C1/C2 `e2e` is what a user of the reference calls, `Map.optimize_local(B200Optimizer, keyframes)`
             on the window as an object graph (collection + packing + H2D + solve + D2H + write-back, wall
             clock); the host-buffer C-ABI figure is kept as `e2e_abi`.

`--impl reference` times the reference's own CPU implementation of the path: the UNMODIFIED
`TorchOptimizer.optimize_local` through `Map.optimize_local` (oracle/ref_shim.py) when the reference tree is
importable (baseline/_ref or /root/reference), else the restated loop oracle/ref_adam.py (labelled "port").
One reference iteration is ONE ADAM STEP -- the reference has no LM.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "local_ba_lm_iterations_per_s"
UNIT = "LM iter/s"
FP64_PEAK_TFLOPS = 36.6          # scripts/fp64_peak.cu on this pool's B200 (vector DFMA); DMMA: 37.1 (scripts/dmma_peak.cu)


# --------------------------------------------------------------------------------------------
def algorithmic_bytes(n_cam, n_pt, n_obs, dense=True, nnzb=0):
    """SURVEY.md 8d-D4 / BASELINE.md 3 (float32 values, int32 indices; |S| dense or 144 B per block)."""
    S = (6 * n_cam) ** 2 * 4 if dense else 144 * nnzb
    b_rjs = 17 * n_obs + 48 * n_pt + 72 * n_cam + S
    b_iter = 33 * n_obs + 108 * n_pt + 72 * n_cam + 2 * S
    return b_rjs, b_iter


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cfg_name, key="linearize_schur"):
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(cfg_name, {}).get(key)
    except Exception:
        return None


def load_synth():
    """The generator alone, WITHOUT importing the product package (the reference arm must not map our .so)."""
    spec = importlib.util.spec_from_file_location(
        "_bench_synth", os.path.join(ROOT, "visual-slam-core_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_bench_synth"] = mod
    spec.loader.exec_module(mod)
    return mod


class _stdout_to_stderr:
    """Route file descriptor 1 to 2 while a C library that prints banners initialises."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_name(cfg, n_cam, n_pt, n_obs):
    kind = "local BA window" if n_cam <= 40 else "BAL-scale synthetic BA"
    return f"{cfg}: {kind}, {n_cam} keyframes, {n_pt} map points, {n_obs} observations, pinhole, Huber 5px, 5% outliers"


# ============================================================================================
# reference arm
# ============================================================================================
def _compact_cameras(sub):
    """Keep only the cameras the sample observes (plus camera 0, the gauge), renumbered."""
    cams = np.unique(np.concatenate([[0], sub.obs_cam]))
    remap = -np.ones(sub.n_cam, dtype=np.int64)
    remap[cams] = np.arange(len(cams))
    return dict(intr=sub.intr, R=sub.R[cams], t=sub.t[cams], fixed=sub.fixed[cams], X=sub.X,
                obs_cam=remap[sub.obs_cam].astype(np.int32), obs_pt=sub.obs_pt, obs_uv=sub.obs_uv)


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path on a bounded sample of the workload window."""
    if rank != 0:
        return
    import torch
    synth = load_synth()
    from oracle import ref_shim
    ref, ref_err = None, None
    try:
        ref = ref_shim.import_reference()
    except Exception as e:                       # absent on this box: the restated loop takes over
        ref_err = repr(e)
    p = synth.make_problem(args.config, seed=0)
    n_total = max(args.steps + args.warmup, 1)
    # ~1 ms per observation and Adam iteration: size the sample for ~75 s of timed work
    per_obs = 1.0e-3
    n_sample = int(min(p.n_obs, max(400, 75.0 / n_total / per_obs)))
    sub = p.shard_by_point(0, max(1, round(p.n_obs / n_sample)))
    w = _compact_cameras(sub)
    n_s_obs, n_s_pt, n_s_cam = len(w["obs_cam"]), w["X"].shape[0], w["R"].shape[0]
    host = f"torch {torch.__version__} CPU, torch threads {torch.get_num_threads()} (the loop is a single Python thread), host cpus {os.cpu_count()}"
    sub_cfgs = {}
    if ref is not None:
        kind = "reference"
        cam, kfs, mps = ref_shim.graph_from_arrays(ref, w["intr"], w["R"], w["t"], w["X"], w["obs_cam"], w["obs_pt"], w["obs_uv"])
        ref_shim.run_local_ba(ref, kfs, n_iter=max(args.warmup, 1))             # warm-up (lazy torch init)
        ok, secs, losses, info = ref_shim.run_local_ba(ref, kfs, n_iter=args.steps)
        t_iter = secs / max(args.steps, 1)
        how = (f"UNMODIFIED reference ({ref['path']}): Map.optimize_local -> TorchOptimizer.optimize_local(device=cpu), "
               f"one call of n_iter={args.steps} after a warm-up call of n_iter={max(args.warmup, 1)}, ok={ok}")
        if world == 1 and not args.no_sub:
            # BASELINE config 1 timed IN FULL through the unmodified reference (10 Adam iterations)
            p1 = synth.make_problem("C1", seed=0)
            cam1, kfs1, mps1 = ref_shim.graph_from_arrays(ref, p1.intr, p1.R, p1.t, p1.X, p1.obs_cam, p1.obs_pt, p1.obs_uv)
            ok1, secs1, losses1, _ = ref_shim.run_local_ba(ref, kfs1, n_iter=10)
            sub_cfgs["C1"] = {
                "workload": workload_name("C1", p1.n_cam, p1.n_pt, p1.n_obs), "value": 10.0 / secs1, "unit": "Adam iter/s",
                "seconds_for_10_iterations": secs1, "ok": ok1, "logged_losses": losses1[:3],
                "cpu_baseline": {"value": 10.0 / secs1, "unit": "Adam iter/s", "cores": 1, "kind": "reference",
                                 "sample": f"the FULL window, 10 Adam iterations in one Map.optimize_local call; {host}"}}
    else:
        kind = "port"
        from oracle import ref_adam
        ref_adam.adam_local_ba(w["intr"], w["R"], w["t"], w["fixed"], w["X"], w["obs_cam"][:50], w["obs_pt"][:50],
                               w["obs_uv"][:50], 1, 1e-4)
        out = ref_adam.adam_local_ba(w["intr"], w["R"], w["t"], w["fixed"], w["X"], w["obs_cam"], w["obs_pt"], w["obs_uv"],
                                     n_iter=n_total, lr=1e-4, delta=5.0)
        secs_l = out["secs"][args.warmup:]
        t_iter = sum(secs_l) / max(len(secs_l), 1)
        how = (f"oracle/ref_adam.py, the reference's Adam loop restated (reference tree not importable here: {ref_err}); "
               f"{len(secs_l)} timed iterations after {args.warmup} warm-up iterations")
    # one reference iteration over the full window = t_iter * n_obs / sample (linear in n_obs: SURVEY 6)
    value = (n_s_obs / p.n_obs) / t_iter
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_iter, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, p.n_cam, p.n_pt, p.n_obs),
                   "iteration": "ONE ADAM STEP of the reference (it has no Levenberg-Marquardt; SURVEY fact 2) is set against one LM iteration",
                   "step": f"one Adam iteration over a sample of {n_s_obs} observations ({n_s_pt} points, {n_s_cam} keyframes)",
                   "extrapolation": "value = (sample_obs / window_obs) / seconds_per_iteration; time is linear in observations (SURVEY 6)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"{n_s_obs} of {p.n_obs} observations; {how}; {host}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if sub_cfgs:
        line["configs"] = sub_cfgs
    print(json.dumps(line), flush=True)


def reference_pnp(pts3d, pts2d, K):
    """The reference's pose estimate for one frame, call for call
    (`visual_slam/utils/motion_estimation.py:100-131`: cv2.solvePnPRansac, SOLVEPNP_ITERATIVE,
    100 iterations, 3 px, 0.99, zero distortion) -- the reference CPU path of BASELINE config 3."""
    import cv2
    ok, rvec, tvec, inl = cv2.solvePnPRansac(objectPoints=pts3d, imagePoints=pts2d, cameraMatrix=K,
                                             distCoeffs=np.zeros(5), iterationsCount=100,
                                             reprojectionError=3.0, confidence=0.99,
                                             flags=cv2.SOLVEPNP_ITERATIVE)
    if not ok:
        return None, None, False, inl
    R, _ = cv2.Rodrigues(rvec)
    return R, tvec.reshape(3, 1), True, inl


# ============================================================================================
# B200 arm: helpers
# ============================================================================================
class Ctx:
    """Per-process device context."""

    def __init__(self, rank, local_rank, world):
        import torch
        self.torch = torch
        self.rank, self.local_rank, self.world = rank, local_rank, world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the solver has no CPU fallback")
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.stream = torch.cuda.current_stream(self.dev)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            with _stdout_to_stderr():          # NCCL prints its banner on stdout; the contract is ONE JSON line
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def T(self, a):
        return self.torch.as_tensor(np.ascontiguousarray(a)).to(self.dev)

    def pin(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a).copy()).pin_memory().numpy()


class DeviceWindow:
    """A packed window resident in HBM + the in-place state tensors of a solve."""

    def __init__(self, ctx, p):
        T = ctx.T
        self.p = p
        self.R0, self.t0, self.X0 = T(p.R), T(p.t), T(p.X)
        self.R, self.t, self.X = self.R0.clone(), self.t0.clone(), self.X0.clone()
        self.fixed, self.oc, self.op = T(p.fixed), T(p.obs_cam), T(p.obs_pt)
        self.uv, self.off = T(p.obs_uv), T(p.pt_offsets)
        self.flags = ctx.torch.zeros(p.n_obs, dtype=ctx.torch.uint8, device=ctx.dev)

    def reset(self):
        self.R.copy_(self.R0); self.t.copy_(self.t0); self.X.copy_(self.X0)

    def solve(self, solver, stream, sync):
        self.reset()
        return solver.solve_local_device(self.p.intr, self.R, self.t, self.fixed, self.X, self.oc, self.op,
                                         self.uv, self.off, self.flags, stream=stream.cuda_stream, sync=sync)


def time_resident(ctx, win, solver, steps, warmup):
    """K timed solves with CUDA events on the launching stream, L2 flushed between; returns
    (ms summed over the K steps on this rank, last warm-up report, clocks)."""
    torch = ctx.torch
    rep = None
    for _ in range(max(warmup, 3)):
        rep = win.solve(solver, ctx.stream, sync=True)
    sampler = ClockSampler(ctx.local_rank)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    sampler.start()
    for a, b in evs:
        ctx.flush.zero_()
        a.record(ctx.stream)
        win.solve(solver, ctx.stream, sync=False)
        b.record(ctx.stream)
    ctx.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    return ms, rep, sampler.stop()


def time_e2e_abi(ctx, p, solver, steps, warmup):
    """The host-buffer C ABI with pinned host memory: H2D, solve, D2H inside the timed wall-clock region."""
    pin = ctx.pin
    n_calls = max(warmup, 3) + steps
    # the API refines R, t, X in place: every call gets its own pinned copy of the initial state, filled BEFORE the
    # timed region, so that the region holds the call itself (H2D, solve, D2H) and no host-side reset copies
    states = [(pin(p.R), pin(p.t), pin(p.X)) for _ in range(n_calls)] if p.X.nbytes * n_calls < 2e9 else None
    hR, ht, hX = pin(p.R), pin(p.t), pin(p.X)
    hfixed, hoc, hop, huv, hoff = pin(p.fixed), pin(p.obs_cam), pin(p.obs_pt), pin(p.obs_uv), pin(p.pt_offsets)
    hflags = pin(np.zeros(p.n_obs, dtype=np.uint8))
    it = iter(states) if states is not None else None

    def e2e_step():
        if it is not None:
            r_, t_, x_ = next(it)
        else:
            np.copyto(hR, p.R); np.copyto(ht, p.t); np.copyto(hX, p.X)
            r_, t_, x_ = hR, ht, hX
        return solver.solve_local_inplace(p.intr, r_, t_, hfixed, x_, hoc, hop, huv, hoff, hflags)

    res = None
    for _ in range(max(warmup, 3)):
        res = e2e_step()
    ctx.barrier()
    t_a = time.perf_counter()
    for _ in range(steps):
        res = e2e_step()
    ctx.torch.cuda.synchronize(ctx.dev)
    secs = ctx.max_over_ranks(time.perf_counter() - t_a)
    h2d = sum(a.nbytes for a in (hR, ht, hfixed, hX, hoc, huv, hoff))     # obs_pt is expanded from pt_offsets on the device
    d2h = hR.nbytes + ht.nbytes + hX.nbytes + hflags.nbytes + 2240        # + LM control block (report)
    return secs, res, int(h2d), int(d2h)


def profile_kernels(ctx, win, solver_factory, lmcfg, max_iter, reps):
    """CUDA events around every kernel (no graph), same workload: ms per LM iteration per kernel class."""
    from visual_slam_core_b200 import LMConfig
    prof = solver_factory(LMConfig(profile=True, **lmcfg))
    acc = np.zeros(7)
    r = None
    for i in range(reps + 2):
        r = win.solve(prof, ctx.stream, sync=True)
        if i >= 2:
            acc += np.array([r.ms_linearize, r.ms_solve, r.ms_backsub, r.ms_other, r.ms_comm, r.ms_point_blocks, r.ms_cam_blocks])
    acc /= reps
    k_ms = {"linearize_schur": acc[0] / max_iter, "rcs_solve": acc[1] / max_iter,
            "backsub_cost": acc[2] / max_iter, "init_finalize_per_step": acc[3],
            "nccl_allreduce": acc[4] / max_iter}
    if acc[5] > 0:          # panel path: the three kernels of a linearisation, timed separately
        k_ms["linearize_parts"] = {"point_blocks": acc[5] / max_iter, "cam_blocks": acc[6] / max_iter,
                                   "schur_panel": (acc[0] - acc[5] - acc[6]) / max_iter}
    return k_ms, r


def roofline_block(cfg_name, p, k_ms, ms_per_iter, band_blocks=None):
    dense = p.n_cam <= 38
    b_rjs, b_iter = algorithmic_bytes(p.n_cam, p.n_pt, p.n_obs, dense=dense,
                                      nnzb=band_blocks if band_blocks else 33 * p.n_cam)
    peak, peak_src = hbm_peak()
    lin_ms = k_ms["linearize_schur"]
    lin_gbs = b_rjs / (lin_ms * 1e-3) / 1e9
    k_obs = np.diff(p.pt_offsets).astype(np.float64)
    flops = 300.0 * p.n_obs + 216.0 * float((k_obs * k_obs).sum())       # SURVEY 8d-D4 "algorithmic flops"
    fp64_tf = flops / (lin_ms * 1e-3) / 1e12
    dom = max(("linearize_schur", "rcs_solve", "backsub_cost"), key=lambda k: k_ms[k])
    resident = b_iter < 100e6
    parts = None
    if "linearize_parts" in k_ms:
        # what each of the three kernels of the panel path has to move (its own inputs once, its outputs once)
        lp = k_ms["linearize_parts"]
        own = {"point_blocks": 16 * p.n_obs + 24 * p.n_pt + 80 * p.n_pt + 96 * p.n_cam,
               "cam_blocks": 12 * p.n_obs + 48 * p.n_obs + 96 * p.n_cam + 216 * p.n_cam,
               "schur_panel": 16 * p.n_obs + 104 * p.n_pt + 96 * p.n_cam + (b_rjs - 17 * p.n_obs - 48 * p.n_pt - 72 * p.n_cam)}
        parts = {k: {"launch_ms": lp[k], "bytes": own[k], "achieved": own[k] / (lp[k] * 1e-3) / 1e9,
                     "frac": own[k] / (lp[k] * 1e-3) / 1e9 / peak} for k in lp if lp[k] > 0}
        parts["dominant"] = max(lp, key=lambda k: lp[k])
    return {"bound": "hbm", "kernel": "linearise (residual + Jacobian + Schur)", "achieved": lin_gbs, "peak": peak,
            "unit": "GB/s", "frac": lin_gbs / peak, "traffic": ncu_traffic(cfg_name), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": b_rjs, "launch_ms": lin_ms, "dominant_kernel_by_time": dom,
            "kernels_of_the_linearisation": parts,
            "whole_iteration": {"algorithmic_bytes": b_iter, "achieved": b_iter / (ms_per_iter * 1e-3) / 1e9,
                                "frac": b_iter / (ms_per_iter * 1e-3) / 1e9 / peak},
            "fp64": {"algorithmic_flop_per_launch": flops, "achieved_tflops": fp64_tf, "peak_tflops": FP64_PEAK_TFLOPS,
                     "frac": fp64_tf / FP64_PEAK_TFLOPS,
                     "note": "all arithmetic is float64 (vector DFMA + mma.m8n8k4.f64), required by the parity tolerances "
                             "(scripts/f32_jacobian_experiment.py); peak measured with scripts/fp64_peak.cu"},
            "note": ("window is L2-resident (working set << 126 MB L2): bound by FP64 latency, see DESIGN.md" if resident else
                     "float64 arithmetic puts the kernel's floor at the FP64 pipe, not at HBM: see roofline.fp64 and DESIGN.md")}


def plugin_leg(p, device_index, max_iter, reps=5):
    """What a user of the reference calls after swapping the optimizer: `Map.optimize_local(B200Optimizer, window)`
    on the window as an object graph (stand-in KeyFrame / MapPoint types with the reference's accessors):
    window collection + packing + H2D + CUDA solve + D2H + write-back, wall clock."""
    import tempfile
    from visual_slam_core_b200 import B200Optimizer, handlers, packing
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refgraph
    cam, kfs, mps = refgraph.graph_from_problem(p)

    class _Opt:
        lm_max_iter = max_iter
    class _Cfg:
        optimization = _Opt()
    opt = B200Optimizer(_Cfg(), log_dir=tempfile.mkdtemp(prefix="b200ba_bench_"), device=f"cuda:{device_index}")
    m = refgraph.Map()
    R0 = [np.array(kf.R_w2c, dtype=np.float64) for kf in kfs]
    t0 = [np.array(kf.t_w2c, dtype=np.float64).reshape(3, 1) for kf in kfs]
    X0 = [np.array(mp.position, dtype=np.float64) for mp in mps]
    total, inner, site, ok = [], [], [], True

    def reset():
        for kf, R, t in zip(kfs, R0, t0):
            kf.set_pose_Rt(R.copy(), t.copy())
        for mp, x in zip(mps, X0):
            mp.update_position(x.copy())

    for _ in range(reps + 2):                               # the first calls create the solver and capture the graph
        reset()
        t_a = time.perf_counter()
        ok = bool(m.optimize_local(opt, kfs)) and ok        # the reference's own Map.optimize_local (its Python collection loop included)
        total.append(time.perf_counter() - t_a)
    window_points = handlers.collect_window_points(kfs)
    for _ in range(reps):
        reset()
        t_a = time.perf_counter()
        ok = bool(opt.optimize_local(kfs, window_points)) and ok     # the plugin alone: pack + H2D + solve + D2H + write-back
        inner.append(time.perf_counter() - t_a)
        reset()
        t_a = time.perf_counter()
        ok = bool(handlers.local_optimize(opt, kfs)) and ok          # this repo's call site: C collection + the plugin + both mean errors
        site.append(time.perf_counter() - t_a)
    res = opt.last_result
    t_a = time.perf_counter()
    pts = handlers.collect_window_points(kfs)
    t_b = time.perf_counter()
    pw = opt.pack(kfs, pts) if hasattr(opt, "pack") else packing.pack_window(kfs, pts)
    t_c = time.perf_counter()
    packing.write_back(kfs, pts, pw.R, pw.t, pw.X)
    t_d = time.perf_counter()
    timed = total[2:]
    return {"seconds_per_call": statistics.median(timed), "seconds_plugin_only": statistics.median(inner),
            "seconds_call_site": statistics.median(site), "ok": ok, "result": res,
            "breakdown_ms": {"collect": 1e3 * (t_b - t_a), "pack": 1e3 * (t_c - t_b), "write_back": 1e3 * (t_d - t_c)},
            "packer": getattr(opt, "packer_name", "pack_window"),
            "h2d": int(pw.R.nbytes + pw.t.nbytes + pw.fixed.nbytes + pw.X.nbytes + pw.obs_cam.nbytes + pw.obs_uv.nbytes + pw.pt_offsets.nbytes),
            "d2h": int(pw.R.nbytes + pw.t.nbytes + pw.X.nbytes + pw.obs_cam.shape[0] + 2240)}


def oracle_parity(res, ref, tol_cost=1e-5):
    """Trajectory parity of a CUDA solve against the oracle on the same window (north_star tolerances)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import rot_angle, scale_norm_trans_err
    rel = float(np.max(np.abs(res.cost_hist - ref.cost_hist) / ref.cost_hist))
    rot = float(rot_angle(res.R, ref.R).max())
    tr = float(scale_norm_trans_err(res.R, res.t, ref.R, ref.t).max())
    acc = bool(np.array_equal(res.accepted, ref.accepted))
    fl = int((res.flags != ref.flags).sum())
    return {"max_rel_cost_diff": rel, "accept_sequence_equal": acc, "max_rot_err_rad": rot,
            "max_scale_norm_trans_err": tr, "flag_mismatches": fl, "n_flags": int(ref.flags.shape[0]),
            "ok": bool(rel <= tol_cost and rot <= 1e-5 and tr <= 1e-5 and fl == 0 and acc)}


# ============================================================================================
# sub-configs (N = 1): C1, C2 local windows and C3 pose-only
# ============================================================================================
def bench_window(ctx, cfg_name, args, with_cpu=True):
    from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
    max_iter = args.lm_iters
    steps, warmup = max(args.steps, 10), max(args.warmup, 3)
    p = synth.make_problem(cfg_name, seed=0)
    lmcfg = dict(max_iter=max_iter, schur=args.schur)
    solver = LocalBASolver(LMConfig(**lmcfg), device=ctx.local_rank)
    win = DeviceWindow(ctx, p)
    ms, rep, clocks = time_resident(ctx, win, solver, steps, warmup)
    value = steps * max_iter / (ms / 1e3)
    secs, res, h2d, d2h = time_e2e_abi(ctx, p, solver, steps, warmup)
    k_ms, _ = profile_kernels(ctx, win, lambda c: LocalBASolver(c, device=ctx.local_rank), lmcfg, max_iter, 5)
    out = {"workload": workload_name(cfg_name, p.n_cam, p.n_pt, p.n_obs), "metric": METRIC, "unit": UNIT,
           "value": value, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "dtype": "f64",
           "obs_per_s_residual_jacobian_schur": p.n_obs / (k_ms["linearize_schur"] * 1e-3),
           "kernel_ms_per_lm_iteration": k_ms,
           "roofline": roofline_block(cfg_name, p, k_ms, ms / steps / max_iter),
           "e2e_abi": {"value": steps * max_iter / secs, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "api": "LocalBASolver.solve_local_inplace -> ba_solve_local_host (pinned host buffers)"},
           "gpu_launches": int(rep.n_launches) * steps, "clocks": clocks}
    try:
        pl = plugin_leg(p, ctx.local_rank, max_iter)
        out["e2e"] = {"value": max_iter / pl["seconds_per_call"], "unit": UNIT, "ms_per_call": 1e3 * pl["seconds_per_call"],
                      "h2d_bytes_per_step": pl["h2d"], "d2h_bytes_per_step": pl["d2h"], "ok": pl["ok"],
                      "breakdown_ms": pl["breakdown_ms"], "packer": pl["packer"],
                      "api": "Map.optimize_local(B200Optimizer, keyframes): object graph in, objects mutated (the reference's own "
                             "Python collection loop, map/map.py:159-169, + packing + H2D + solve + D2H + write-back, wall clock)",
                      "ms_plugin_only": 1e3 * pl["seconds_plugin_only"],
                      "plugin_only_api": "B200Optimizer.optimize_local(keyframes, points): what the swapped-in class itself costs",
                      "ms_call_site": 1e3 * pl["seconds_call_site"],
                      "call_site_api": "handlers.local_optimize(optimizer, keyframes): collection in C + the plugin + mean "
                                       "reprojection error before / after (LocalHandler.local_optimize restated)"}
    except Exception as e:                                  # never costs the bench line
        out["e2e"] = {"error": repr(e)}
    if with_cpu:
        from oracle import lm_oracle as O
        t_c = time.perf_counter()
        ref = O.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, O.LMParams(max_iter=max_iter))
        dt = time.perf_counter() - t_c
        out["cpu_baseline"] = {"value": max_iter / dt, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"the full window, one solve of {max_iter} LM iterations with oracle/lm_oracle.py "
                                         f"(numpy float64, same algorithm) in {dt:.1f} s"}
        out["parity"] = dict(oracle_parity(res, ref), vs="oracle/lm_oracle.py, full trajectory on the full window")
    solver.close()
    return out


def bench_pose_only(ctx, args, impl_reference=False):
    """BASELINE config 3: batched pose-only BA, 256 frames x 1000 matches, one B200."""
    import ctypes as C
    from visual_slam_core_b200 import LMConfig, LocalBASolver, _lib, synth
    torch = ctx.torch
    pp = synth.make_pose_only(n_frames=256, n_match=1000, seed=0)
    nF, nM = pp.n_frames, pp.n_match
    K = np.array([[pp.intr[0], 0, pp.intr[2]], [0, pp.intr[1], pp.intr[3]], [0, 0, 1.0]])
    max_iter = args.lm_iters
    steps, warmup = max(args.steps, 10), max(args.warmup, 3)
    unit = "frame LM iter/s"
    solver = LocalBASolver(LMConfig(max_iter=max_iter), device=ctx.local_rank)
    lib = solver.lib
    T = ctx.T
    R0, t0 = T(pp.R), T(pp.t)
    R, t = R0.clone(), t0.clone()
    X, uv, fo = T(pp.X), T(pp.uv), T(pp.frame_offsets)
    flags = torch.zeros(nM, dtype=torch.uint8, device=ctx.dev)
    hist = torch.zeros((nF, max_iter + 1), dtype=torch.float64, device=ctx.dev)
    intr = np.ascontiguousarray(pp.intr, dtype=np.float64)
    stream = ctx.stream

    def launch():
        rc = lib.ba_solve_pose_only(solver._h, nF, nM, intr.ctypes.data_as(C.POINTER(C.c_double)), R.data_ptr(),
                                    t.data_ptr(), X.data_ptr(), uv.data_ptr(), fo.data_ptr(), flags.data_ptr(),
                                    hist.data_ptr(), None, C.c_void_p(stream.cuda_stream))
        _lib.check(lib, solver._h, rc, "ba_solve_pose_only")

    for _ in range(warmup):
        R.copy_(R0); t.copy_(t0)
        launch()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    for (a, b), (ka, kb) in zip(evs, kev):
        ctx.flush.zero_()
        a.record(stream)
        R.copy_(R0); t.copy_(t0)
        ka.record(stream)
        launch()
        kb.record(stream)
        b.record(stream)
    ctx.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / steps
    value = steps * nF * max_iter / (ms / 1e3)

    pin = ctx.pin
    hR, ht, hX, huv, hfo = pin(pp.R), pin(pp.t), pin(pp.X), pin(pp.uv), pin(pp.frame_offsets)
    hflags = pin(np.zeros(nM, dtype=np.uint8))
    hhist = pin(np.zeros((nF, max_iter + 1), dtype=np.float64))
    hacc = pin(np.zeros((nF, max(max_iter, 1)), dtype=np.uint8))

    def e2e_step():
        np.copyto(hR, pp.R); np.copyto(ht, pp.t)
        return solver.solve_pose_only(pp.intr, hR, ht, hX, huv, hfo, inplace=True, flags=hflags,
                                      cost_hist=hhist, accepted=hacc)

    for _ in range(warmup):
        out = e2e_step()
    t_a = time.perf_counter()
    for _ in range(steps):
        out = e2e_step()
    torch.cuda.synchronize(ctx.dev)
    e2e_value = steps * nF * max_iter / (time.perf_counter() - t_a)
    h2d = pp.R.nbytes + pp.t.nbytes + pp.X.nbytes + pp.uv.nbytes + pp.frame_offsets.nbytes
    d2h = pp.R.nbytes + pp.t.nbytes + nM + nF * (max_iter + 1) * 8 + nF * max_iter
    b_po = 21 * nM + 96 * nF                                   # SURVEY 8d-D4, one pass over the matches
    peak, peak_src = hbm_peak()
    gbs = b_po / (k_ms * 1e-3) / 1e9
    res = {"workload": f"C3: batched pose-only BA, {nF} frames x {nM // nF} 2D-3D matches, pinhole, Huber 5px, 10% outliers",
           "metric": "pose_only_lm_iterations_per_s", "unit": unit, "value": value, "frames_per_s": value / max_iter,
           "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "dtype": "f64",
           "kernel_ms_per_step": {"pose_only_lm": k_ms},
           "roofline": {"bound": "hbm", "kernel": "k_pose_only (whole LM loop of every frame in one launch)", "achieved": gbs,
                        "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": ncu_traffic("C3", "pose_only"),
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": b_po, "launch_ms": k_ms,
                        "note": f"the matches are read once from HBM and then {2 * max_iter} times from L2 / L1 (linearise + candidate "
                                f"cost per LM iteration): the kernel is bound by FP64 latency, not by HBM"},
           "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "api": "LocalBASolver.solve_pose_only(inplace, pinned host buffers) -> ba_solve_pose_only_host"},
           "gpu_launches": steps}
    if not args.no_cpu:
        from oracle import lm_oracle as O
        t_c = time.perf_counter()
        ref = O.solve_pose_only(pp.intr, pp.R, pp.t, pp.X, pp.uv, pp.frame_offsets, O.LMParams(max_iter=max_iter))
        dt = time.perf_counter() - t_c
        res["oracle_cpu"] = {"value": nF * max_iter / dt, "unit": unit, "cores": 1,
                             "sample": f"all {nF} frames, oracle/lm_oracle.py solve_pose_only (numpy float64) in {dt:.1f} s"}
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import rot_angle
        res["parity"] = {
            "vs": "oracle/lm_oracle.py solve_pose_only, all frames",
            "max_rel_cost_diff": float(np.max(np.abs(out["cost_hist"] - ref["cost_hist"]) / ref["cost_hist"])),
            "accept_sequence_equal": bool(np.array_equal(out["accepted"], ref["accepted"])),
            "max_rot_err_rad": float(rot_angle(out["R"].reshape(-1, 3, 3), ref["R"]).max()),
            "max_trans_err": float(np.abs(out["t"].reshape(-1, 3) - ref["t"]).max()),
            "flag_mismatches": int((out["flags"] != ref["flags"]).sum()), "n_flags": int(nM)}
        res["parity"]["ok"] = bool(res["parity"]["max_rel_cost_diff"] <= 1e-5 and res["parity"]["flag_mismatches"] == 0
                                   and res["parity"]["max_rot_err_rad"] <= 1e-5 and res["parity"]["max_trans_err"] <= 1e-5)
        try:
            import cv2
            cv2.setRNGSeed(42)
            n_r = 48
            t_c = time.perf_counter()
            for f in range(n_r):
                fa, fb = pp.frame_offsets[f], pp.frame_offsets[f + 1]
                reference_pnp(pp.X[fa:fb], pp.uv[fa:fb], K)
            dt_r = time.perf_counter() - t_c
            res["cpu_baseline"] = {"value": n_r / dt_r * max_iter, "unit": unit, "cores": int(cv2.getNumThreads()),
                                   "kind": "reference",
                                   "sample": f"{n_r} of {nF} frames, cv2 {cv2.__version__} solvePnPRansac as called at "
                                             f"motion_estimation.py:107-116 ({1e3 * dt_r / n_r:.2f} ms per frame); value = frames/s x {max_iter}"}
        except Exception as e:  # OpenCV missing on the box: the oracle port is the baseline
            res["cpu_baseline"] = dict(res["oracle_cpu"], kind="port", note=f"cv2 unavailable: {e!r}")
    solver.close()
    return res



def _solve_state(ctx, win, solver, max_iter):
    rep = win.solve(solver, ctx.stream, sync=True)
    return dict(R=win.R.cpu().numpy().copy(), t=win.t.cpu().numpy().copy(), X=win.X.cpu().numpy().copy(),
                flags=win.flags.cpu().numpy().copy(), hist=np.array(rep.cost_hist[: max_iter + 1]),
                acc=np.array(rep.accepted[:max_iter], dtype=np.uint8))


def _compare(ctx, a, b, n_cam):
    """Differences between two solves of the same window (a: this rank's shard or the whole window; b: the whole
    window), reduced over the ranks."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import gauge_aligned_trans_err, rot_angle, scale_norm_trans_err
    terr = (gauge_aligned_trans_err(a["R"], a["t"], b["R"], b["t"])[0] if n_cam > 40
            else scale_norm_trans_err(a["R"], a["t"], b["R"], b["t"]))
    rel = np.abs(a["hist"] - b["hist"]) / b["hist"]
    return {"rel_cost_diff_per_iteration": [ctx.max_over_ranks(float(v)) for v in rel],
            "max_rel_cost_diff": ctx.max_over_ranks(float(rel.max())),
            "accept_sequence_equal": ctx.max_over_ranks(0.0 if np.array_equal(a["acc"], b["acc"]) else 1.0) == 0.0,
            "max_rot_err_rad": ctx.max_over_ranks(float(rot_angle(a["R"], b["R"]).max())),
            "max_trans_err": ctx.max_over_ranks(float(terr.max())),
            "max_point_diff": ctx.max_over_ranks(float(np.abs(a["X"] - b["X"]).max()) if b["X"].size else 0.0),
            "flag_mismatches": int(ctx.sum_over_ranks(float((a["flags"] != b["flags"]).sum())))}


def sharded_parity(ctx, args, full, p, solver, lmcfg, sh):
    """The sharded solve against the single-GPU solve of the same window, run in this job on every rank's own GPU.

    Two windows: (1) `L1`, 1 000 cameras x 50 000 points (tests/golden/make_large_golden.py): the shape of the headline
    window on which ten LM iterations are numerically stable -- there the BASELINE tolerances apply as they stand
    (cost 1e-6, accept sequence, poses 1e-5, flags bit-exact); (2) the headline window itself.  Its ten-iteration
    trajectory is chaotic (lambda sits at its floor 1e-6 from the third iteration on, the 1 000-camera chain amplifies
    rounding by 1-2 orders of magnitude per iteration: two runs of the SAME single-GPU binary differ by 1e-13, 1e-11,
    1e-10, 1e-6 ... 1e-2 in cost, scripts/diag_c4.py), so the sharded run is held to the initial cost and the first two iterations
    (<= 1e-9) and, after that, to 20 x the run-to-run spread of the single-GPU solver measured here.
    In both, the replicated state (poses, cost history) must be bit-identical across the ranks."""
    import b200ba  # noqa: F401
    torch = ctx.torch
    from visual_slam_core_b200 import LMConfig, LocalBASolver
    max_iter = args.lm_iters
    one = LocalBASolver(LMConfig(**lmcfg), device=ctx.local_rank)
    out = {}

    def shard_of(state, prob_full, shard):
        lo, hi = shard.meta["pt_range"]
        o0, o1 = int(prob_full.pt_offsets[lo]), int(prob_full.pt_offsets[hi])
        return dict(state, X=state["X"][lo:hi], flags=state["flags"][o0:o1])

    def identical_across_ranks(state):
        v = torch.as_tensor(np.concatenate([state["R"].reshape(-1), state["t"].reshape(-1), state["hist"]])).to(ctx.dev)
        v0 = v.clone()
        ctx.dist.broadcast(v0, src=0)
        return int(ctx.max_over_ranks(float((v != v0).sum().item())))

    # ---- (1) the stable C4-shaped window
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_large_golden as G
    l1 = G.make_window("L1")
    l1s = l1.shard_by_point(ctx.rank, ctx.world)
    a = _solve_state(ctx, DeviceWindow(ctx, l1s), solver, max_iter)
    b = _solve_state(ctx, DeviceWindow(ctx, l1), one, max_iter)
    c1 = _compare(ctx, a, shard_of(b, l1, l1s), l1.n_cam)
    c1["entries_differing_across_ranks"] = identical_across_ranks(a)
    c1["window"] = f"L1: {l1.n_cam} cameras, {l1.n_pt} points, {l1.n_obs} observations (same co-visibility band as the headline window)"
    c1["ok"] = bool(c1["max_rel_cost_diff"] <= 1e-6 and c1["accept_sequence_equal"] and c1["max_rot_err_rad"] <= 1e-5
                    and c1["max_trans_err"] <= 1e-5 and c1["flag_mismatches"] == 0 and c1["entries_differing_across_ranks"] == 0)
    out["stable_window"] = c1

    # ---- (2) the headline window: two single-GPU runs give the noise floor, the sharded run is set against it
    wfull = DeviceWindow(ctx, full)
    r1 = _solve_state(ctx, wfull, one, max_iter)
    r2 = _solve_state(ctx, wfull, one, max_iter)
    floor = np.abs(r1["hist"] - r2["hist"]) / r1["hist"]
    floor = np.array([ctx.max_over_ranks(float(v)) for v in floor])
    c2 = _compare(ctx, sh, shard_of(r1, full, p), full.n_cam)
    c2["single_gpu_run_to_run_rel_cost_diff_per_iteration"] = floor.tolist()
    c2["entries_differing_across_ranks"] = identical_across_ranks(sh)
    d = np.array(c2["rel_cost_diff_per_iteration"])
    c2["first_iterations_ok"] = bool(np.all(d[:3] <= 1e-9))
    c2["within_20x_noise_floor"] = bool(np.all(d <= np.maximum(1e-6, 20.0 * floor)))
    c2["ok"] = bool(c2["first_iterations_ok"] and c2["within_20x_noise_floor"] and c2["entries_differing_across_ranks"] == 0)
    c2["note"] = ("ten LM iterations of this window are chaotic in float64 (see the run-to-run row); the 1e-6 bound is asserted "
                  "on the stable window above and on the first iterations here")
    out["headline_window"] = c2
    out["vs"] = "the single-GPU solve of the same window, run in this job on every rank's own GPU (N = 1 path of the same library)"
    out["ok"] = bool(c1["ok"] and c2["ok"])
    one.close()
    return out


# ============================================================================================
# B200 arm: the headline window, sharded by point over the ranks
# ============================================================================================
def run_b200(args, rank, local_rank, world):
    ctx = Ctx(rank, local_rank, world)
    torch = ctx.torch
    import b200ba  # noqa: F401
    from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
    from visual_slam_core_b200 import dist as badist

    max_iter = args.lm_iters
    steps, warmup = args.steps, max(args.warmup, 3)
    full = synth.make_problem(args.config, seed=0)
    p = full.shard_by_point(rank, world) if world > 1 else full
    lmcfg = dict(max_iter=max_iter, schur=args.schur, pcg_tol=args.pcg_tol, pcg_max_iter=args.pcg_max_iter)
    solver = LocalBASolver(LMConfig(**lmcfg), device=local_rank)
    if world > 1:
        with _stdout_to_stderr():
            badist.init_solver(solver, rank, world, device=ctx.dev)
    win = DeviceWindow(ctx, p)

    # ---- resident
    ms, rep, clocks = time_resident(ctx, win, solver, steps, warmup)
    ms_max = ctx.max_over_ranks(ms)
    value = steps * max_iter / (ms_max / 1e3)
    launches_per_step = int(rep.n_launches)
    # the state the timed solves end in (every step starts from the same state, so any of them will do)
    rep = win.solve(solver, ctx.stream, sync=True)
    R_sh, t_sh = win.R.cpu().numpy().copy(), win.t.cpu().numpy().copy()
    X_sh, flags_sh = win.X.cpu().numpy().copy(), win.flags.cpu().numpy().copy()
    hist_sh = np.array(rep.cost_hist[: max_iter + 1])
    acc_sh = np.array(rep.accepted[:max_iter], dtype=np.uint8)

    # ---- end to end through the host-buffer C ABI
    secs, res, h2d, d2h = time_e2e_abi(ctx, p, solver, steps, warmup)
    e2e_value = steps * max_iter / secs
    h2d_all, d2h_all = int(ctx.sum_over_ranks(h2d)), int(ctx.sum_over_ranks(d2h))

    # ---- per-kernel device times
    if world > 1:
        def factory(cfg):
            solver.set_config(cfg)
            return solver
    else:
        factory = lambda cfg: LocalBASolver(cfg, device=local_rank)   # noqa: E731
    k_ms, r_prof = profile_kernels(ctx, win, factory, lmcfg, max_iter, 3)
    k_ms = {k: ({kk: ctx.max_over_ranks(vv) for kk, vv in v.items()} if isinstance(v, dict) else ctx.max_over_ranks(v))
            for k, v in k_ms.items()}
    if world > 1:
        solver.set_config(LMConfig(**lmcfg))
    band_blocks = int(r_prof.s_blocks) if int(r_prof.band_width) >= 0 else None

    # ---- parity
    parity = {}
    if world > 1:
        parity = sharded_parity(ctx, args, full, p, solver, lmcfg, dict(R=R_sh, t=t_sh, X=X_sh, flags=flags_sh, hist=hist_sh, acc=acc_sh))
    elif not args.no_cpu:
        # full size, size-independent properties against the oracle's restatement of the reference objective
        from oracle import lm_oracle as O
        c0 = O.reference_cost(full.intr, full.R, full.t, full.X, full.obs_cam, full.obs_pt, full.obs_uv)
        c1 = O.reference_cost(full.intr, R_sh, t_sh, X_sh, full.obs_cam, full.obs_pt, full.obs_uv)
        fl = O.outlier_flags(full.intr, R_sh, t_sh, X_sh, full.obs_cam, full.obs_pt, full.obs_uv)
        parity = {
            "vs": "oracle/lm_oracle.py objective and flag rule evaluated at the CUDA solver's initial / final state, full window "
                  "(trajectory parity at this camera count is asserted against the sparse oracle in tests/test_gpu_large.py)",
            "initial_cost_rel_diff": float(abs(hist_sh[0] - c0) / c0),
            "final_cost_rel_diff": float(abs(hist_sh[-1] - c1) / c1),
            "flag_mismatches": int((fl != flags_sh).sum()), "n_flags": int(full.n_obs),
            "cost_monotone": bool(np.all(np.diff(hist_sh) <= 0)),
            "accepted": acc_sh.tolist(),
        }
        parity["ok"] = bool(parity["initial_cost_rel_diff"] <= 1e-9 and parity["final_cost_rel_diff"] <= 1e-9
                            and parity["flag_mismatches"] == 0 and parity["cost_monotone"])

    lin_ms = k_ms["linearize_schur"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_max / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, full.n_cam, full.n_pt, full.n_obs),
                   "lm_iterations_per_step": max_iter,
                   "parallelism": "single GPU" if world == 1 else
                                  (f"points sharded over {world} GPUs (cameras replicated); every LM iteration the reduced camera system is summed over "
                                   "NVLink peer memory inside the solver's first kernel (CUDA IPC mappings, flag handshake; its time is in rcs_solve) and the "
                                   "candidate cost goes through an NCCL all-reduce (nccl_allreduce)" if int(getattr(r_prof, "peer_exchange", 0)) else
                                   f"points sharded over {world} GPUs (cameras replicated), NCCL all-reduce of the reduced camera system and the costs every LM iteration"),
                   "solver": {"used": {1: "cholesky", 2: "pcg", 3: "band"}.get(int(r_prof.solver_used), "?"),
                              "band_width": int(r_prof.band_width), "s_blocks": int(r_prof.s_blocks)},
                   "l2": "flushed between steps (256 MiB write); the window itself is larger than L2",
                   "schur": args.schur},
        "obs_per_s_residual_jacobian_schur": full.n_obs / (lin_ms * 1e-3),
        "shard_observations": p.n_obs,
        "kernel_ms_per_lm_iteration": k_ms,
        "final_cost": float(hist_sh[-1]),
        "roofline": roofline_block(args.config, p, k_ms, ms_max / steps / max_iter, band_blocks),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                "api": "LocalBASolver.solve_local_inplace -> ba_solve_local_host (pinned host buffers, every rank its shard)"},
        "parity": parity,
        "gpu_launches": launches_per_step * steps,
        "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import lm_oracle as O
        # bounded CPU sample: the first 1/20 of the window's points with every observation of theirs
        frac = 20 if full.n_obs > 2_000_000 else (4 if full.n_obs > 400_000 else 1)
        sub = full.shard_by_point(0, frac)
        t_c = time.perf_counter()
        O.solve_local(sub.intr, sub.R, sub.t, sub.fixed, sub.X, sub.obs_cam, sub.obs_pt, sub.obs_uv,
                      O.LMParams(max_iter=max_iter))
        dt = time.perf_counter() - t_c
        line["cpu_baseline"] = {"value": max_iter / dt * (sub.n_obs / full.n_obs), "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{sub.n_obs} of {full.n_obs} observations ({sub.n_pt} points, the first 1/{frac} of the window), "
                                          f"one solve of {max_iter} LM iterations with oracle/lm_oracle.py (numpy float64 + scipy sparse "
                                          f"direct solve, same algorithm) in {dt:.1f} s; value scaled by sample/window observations"}
    if rank == 0 and world == 1 and not args.no_sub:
        subs = {}
        del win
        torch.cuda.empty_cache()
        for name in ("C1", "C2"):
            try:
                subs[name] = bench_window(ctx, name, args, with_cpu=not args.no_cpu)
            except Exception as e:
                subs[name] = {"error": repr(e)}
        try:
            subs["C3"] = bench_pose_only(ctx, args)
        except Exception as e:
            subs["C3"] = {"error": repr(e)}
        line["configs"] = subs
    if rank == 0:
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


def run_single(args, rank, local_rank, world):
    """`--config C1|C2|C3`: one of the single-GPU configs as the top-level line (replicas for N > 1)."""
    ctx = Ctx(rank, local_rank, world)
    import b200ba  # noqa: F401
    if args.config == "C3":
        sub = bench_pose_only(ctx, args)
    else:
        sub = bench_window(ctx, args.config, args, with_cpu=not args.no_cpu and rank == 0)
    v = ctx.sum_over_ranks(sub["value"])
    line = {"metric": sub["metric"], "value": v, "unit": sub["unit"], "n_gpus": world, "steps": sub["steps"],
            "warmup": sub["warmup"], "ms_per_step": ctx.max_over_ranks(sub["ms_per_step"]), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": sub["workload"], "lm_iterations_per_step": args.lm_iters,
                       "parallelism": "single GPU" if world == 1 else f"{world} independent windows, one per GPU (replicas only: this config does not shard)",
                       "l2": "flushed between steps (256 MiB write); inside a step the window stays L2-resident by construction"}}
    for k in ("obs_per_s_residual_jacobian_schur", "kernel_ms_per_lm_iteration", "kernel_ms_per_step", "roofline", "e2e", "e2e_abi",
              "parity", "cpu_baseline", "oracle_cpu", "gpu_launches", "clocks", "frames_per_s"):
        if k in sub:
            line[k] = sub[k]
    if rank == 0:
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


def run_reference_pose_only(args, rank):
    if rank != 0:
        return
    import cv2
    synth = load_synth()
    pp = synth.make_pose_only(n_frames=256, n_match=1000, seed=0)
    nF, nM = pp.n_frames, pp.n_match
    K = np.array([[pp.intr[0], 0, pp.intr[2]], [0, pp.intr[1], pp.intr[3]], [0, 0, 1.0]])
    max_iter = args.lm_iters
    cv2.setRNGSeed(42)
    total = args.steps + args.warmup
    n_sample = int(min(nF, max(8, 4000 // max(total, 1))))       # ~4 ms per frame: a few seconds per step
    secs = []
    for it in range(total):
        t0 = time.perf_counter()
        for f in range(n_sample):
            a, b = pp.frame_offsets[f], pp.frame_offsets[f + 1]
            reference_pnp(pp.X[a:b], pp.uv[a:b], K)
        secs.append(time.perf_counter() - t0)
    t_step = sum(secs[args.warmup:]) / max(len(secs) - args.warmup, 1)
    fps = n_sample / t_step
    value = fps * max_iter          # one reference pose estimate is set against one 10-iteration LM refinement
    unit = "frame LM iter/s"
    line = {"impl": "reference", "metric": "pose_only_lm_iterations_per_s", "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3: batched pose-only BA, {nF} frames x {nM // nF} 2D-3D matches, pinhole, Huber 5px, 10% outliers",
                       "equivalence": f"value = frames/s x {max_iter}: one cv2.solvePnPRansac call replaces one {max_iter}-iteration pose-only LM"},
            "frames_per_s": fps,
            "cpu_baseline": {"value": value, "unit": unit, "cores": int(cv2.getNumThreads()), "kind": "reference",
                             "sample": f"{n_sample} of {nF} frames per step, cv2 {cv2.__version__} solvePnPRansac exactly as "
                                       f"motion_estimation.py:107-116 calls it, host cpus {os.cpu_count()}"},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", help="C4 (default) / C5: sharded by point over the ranks; C1 / C2 / C3: single-GPU configs")
    ap.add_argument("--lm-iters", type=int, default=10)
    ap.add_argument("--schur", default="auto")
    ap.add_argument("--pcg-tol", type=float, default=1e-10)
    ap.add_argument("--pcg-max-iter", type=int, default=3000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle legs (cpu_baseline, oracle parity)")
    ap.add_argument("--no-sub", action="store_true", help="skip the C1 / C2 / C3 sub-objects of the N = 1 line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.config == "C3":
            run_reference_pose_only(args, rank)
        else:
            run_reference(args, rank, world)
    elif args.config in ("C1", "C2", "C3"):
        run_single(args, rank, local_rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
