#!/usr/bin/env python
"""Benchmark of the local-BA hot path (BASELINE.json metric: LM iterations/s, obs/s through
residual + Jacobian + Schur, against the HBM roofline, with the CPU path timed beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl b200|reference]

A *step* is one complete local-BA solve of the workload window: `max_iter` (10) LM iterations,
each = linearise + Schur -> reduced-camera solve -> back-substitution + candidate cost ->
accept/reject, plus state init and outlier flags.  Every step starts from the same initial state.

  value  = LM iterations/s with inputs resident in HBM (CUDA events around each step, L2 flushed
           between steps, max over ranks; N > 1 = one independent window per GPU, weak scaling).
  e2e    = the same through the public host-buffer API (`LocalBASolver.solve_local_inplace`,
           i.e. the C-ABI `ba_solve_local_host`): pinned host inputs -> H2D -> solve -> D2H of
           poses, points, flags and the LM report, all inside the timed region.
  roofline = dominant kernel (by summed device time, measured live with CUDA events in a
           profile pass of the same workload): algorithmic bytes / launch duration vs measured HBM peak.
  cpu_baseline = oracle/lm_oracle.py (same LM, numpy float64) on this box's host cores.

`--impl reference` times the reference's own CPU algorithm (Adam over a Python per-observation
autograd loop, restated in oracle/ref_adam.py and validated against the unmodified reference in
tests/test_ref_adam.py) on a bounded sample of the same window.
"""
from __future__ import annotations

import argparse
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
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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


# --------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's CPU algorithm on a bounded sample of the workload window."""
    if rank != 0:
        return
    import torch
    import b200ba  # noqa: F401
    from visual_slam_core_b200 import synth
    from oracle import ref_adam
    p = synth.make_problem(args.config, seed=0)
    total = args.steps + args.warmup
    # probe the per-observation cost, then size the sample so the whole run ends in ~2.5 minutes
    probe_n = 60
    sub = p.shard_by_point(0, max(1, p.n_obs // probe_n))
    ref_adam.adam_local_ba(sub.intr, sub.R, sub.t, sub.fixed, sub.X, sub.obs_cam, sub.obs_pt, sub.obs_uv, 1, 1e-4)
    t0 = time.perf_counter()   # the first call pays torch's lazy initialisation: time the second
    ref_adam.adam_local_ba(sub.intr, sub.R, sub.t, sub.fixed, sub.X, sub.obs_cam, sub.obs_pt, sub.obs_uv, 1, 1e-4)
    per_obs = (time.perf_counter() - t0) / max(sub.n_obs, 1)
    budget = 150.0
    n_sample = int(min(p.n_obs, max(200, budget / max(total, 1) / max(per_obs, 1e-6))))
    sub = p.shard_by_point(0, max(1, round(p.n_obs / n_sample)))
    out = ref_adam.adam_local_ba(sub.intr, sub.R, sub.t, sub.fixed, sub.X, sub.obs_cam, sub.obs_pt, sub.obs_uv,
                                 n_iter=total, lr=1e-4, delta=5.0)
    secs = out["secs"][args.warmup:]
    t_step = sum(secs) / max(len(secs), 1)
    # one reference iteration over the full window = t_step * n_obs / sample (linear in n_obs: SURVEY 6)
    value = (sub.n_obs / p.n_obs) / t_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, p), "iteration": "one Adam iteration (the reference has no LM)",
                   "extrapolation": "value = (sample_obs / window_obs) / seconds_per_iteration; time is linear in observations"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{sub.n_obs} of {p.n_obs} observations ({sub.n_pt} points), {len(secs)} timed Adam "
                                   f"iterations, torch {torch.__version__} CPU, torch threads {torch.get_num_threads()} "
                                   f"(the loop is a single Python thread), host cpus {os.cpu_count()}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def plugin_leg(p, device_index, max_iter, reps=3):
    """What a user of the reference sees after swapping the optimizer: `Map.optimize_local(B200Optimizer, window)`
    on the window as an object graph (KeyFrame / MapPoint stand-ins of refgraph.py), i.e. window collection +
    packing + the CUDA solve through the host API + write-back, in wall-clock milliseconds.  Informational: the
    contract's `e2e` is the C-ABI call on host buffers."""
    import tempfile
    from visual_slam_core_b200 import B200Optimizer, handlers, packing, refgraph
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
    total, ok = [], True
    for _ in range(reps + 1):                               # the first call creates the solver and captures the graph
        for kf, R, t in zip(kfs, R0, t0):
            kf.set_pose_Rt(R.copy(), t.copy())
        for mp, x in zip(mps, X0):
            mp.update_position(x.copy())
        t_a = time.perf_counter()
        ok = bool(m.optimize_local(opt, kfs)) and ok
        total.append(1e3 * (time.perf_counter() - t_a))
    t_a = time.perf_counter()
    pts = handlers.collect_window_points(kfs)
    t_b = time.perf_counter()
    pw = packing.pack_window(kfs, pts)
    t_c = time.perf_counter()
    packing.write_back(kfs, pts, pw.R, pw.t, pw.X)
    t_d = time.perf_counter()
    return {"optimize_local_ms": min(total[1:]), "ok": ok, "collect_ms": 1e3 * (t_b - t_a),
            "pack_window_ms": 1e3 * (t_c - t_b), "write_back_ms": 1e3 * (t_d - t_c),
            "packer_loop": "C (csrc/fastpack.c)" if packing._fastpack is not None else "Python",
            "api": "refgraph.Map.optimize_local -> B200Optimizer.optimize_local (object graph in, objects mutated)"}


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


def run_pose_only(args, rank, local_rank, world):
    """BASELINE config 3: batched pose-only BA, 256 frames x 1000 matches, one B200 (replicas for N > 1)."""
    import b200ba  # noqa: F401
    from visual_slam_core_b200 import LMConfig, synth
    pp = synth.make_pose_only(n_frames=256, n_match=1000, seed=rank)
    nF, nM = pp.n_frames, pp.n_match
    K = np.array([[pp.intr[0], 0, pp.intr[2]], [0, pp.intr[1], pp.intr[3]], [0, 0, 1.0]])
    max_iter = args.lm_iters
    metric, unit = "pose_only_lm_iterations_per_s", "frame LM iter/s"
    workload = f"C3: batched pose-only BA, {nF} frames x {nM // nF} 2D-3D matches, pinhole, Huber 5px, 10% outliers"

    if args.impl == "reference":
        if rank != 0:
            return
        import cv2
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
        line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "equivalence": f"value = frames/s x {max_iter}: one cv2.solvePnPRansac call "
                           f"replaces one {max_iter}-iteration pose-only LM"},
                "frames_per_s": fps,
                "cpu_baseline": {"value": value, "unit": unit, "cores": int(cv2.getNumThreads()), "kind": "reference",
                                 "sample": f"{n_sample} of {nF} frames per step, cv2 {cv2.__version__} solvePnPRansac exactly as "
                                           f"motion_estimation.py:107-116 calls it, host cpus {os.cpu_count()}"},
                "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    from visual_slam_core_b200 import LocalBASolver, _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solver has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        with _stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    solver = LocalBASolver(LMConfig(max_iter=max_iter), device=local_rank)
    lib = solver.lib
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    R0, t0 = T(pp.R), T(pp.t)
    R, t = R0.clone(), t0.clone()
    X, uv, fo = T(pp.X), T(pp.uv), T(pp.frame_offsets)
    flags = torch.zeros(nM, dtype=torch.uint8, device=dev)
    hist = torch.zeros((nF, max_iter + 1), dtype=torch.float64, device=dev)
    intr = np.ascontiguousarray(pp.intr, dtype=np.float64)
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        R.copy_(R0); t.copy_(t0)
        rc = lib.ba_solve_pose_only(solver._h, nF, nM, intr.ctypes.data_as(C.POINTER(C.c_double)), R.data_ptr(),
                                    t.data_ptr(), X.data_ptr(), uv.data_ptr(), fo.data_ptr(), flags.data_ptr(),
                                    hist.data_ptr(), None, C.c_void_p(stream.cuda_stream))
        _lib.check(lib, solver._h, rc, "ba_solve_pose_only")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    for (a, b), (ka, kb) in zip(evs, kev):
        flush.zero_()
        a.record(stream)
        R.copy_(R0); t.copy_(t0)
        ka.record(stream)
        lib.ba_solve_pose_only(solver._h, nF, nM, intr.ctypes.data_as(C.POINTER(C.c_double)), R.data_ptr(),
                               t.data_ptr(), X.data_ptr(), uv.data_ptr(), fo.data_ptr(), flags.data_ptr(),
                               hist.data_ptr(), None, C.c_void_p(stream.cuda_stream))
        kb.record(stream)
        b.record(stream)
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    clocks = sampler.stop()
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * args.steps * nF * max_iter / (ms_max / 1e3)

    # end to end: host buffers through ba_solve_pose_only_host (H2D / D2H inside)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).copy()).pin_memory().numpy()
    hR, ht, hX, huv, hfo = pin(pp.R), pin(pp.t), pin(pp.X), pin(pp.uv), pin(pp.frame_offsets)
    hflags = pin(np.zeros(nM, dtype=np.uint8))
    hhist = pin(np.zeros((nF, max_iter + 1), dtype=np.float64))
    hacc = pin(np.zeros((nF, max(max_iter, 1)), dtype=np.uint8))

    def e2e_step():
        np.copyto(hR, pp.R); np.copyto(ht, pp.t)
        return solver.solve_pose_only(pp.intr, hR, ht, hX, huv, hfo, inplace=True, flags=hflags,
                                      cost_hist=hhist, accepted=hacc)

    for _ in range(max(args.warmup, 3)):
        out = e2e_step()
    barrier()
    t_a = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_step()
    torch.cuda.synchronize(dev)
    te = torch.tensor([time.perf_counter() - t_a], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * nF * max_iter / float(te.item())
    h2d = pp.R.nbytes + pp.t.nbytes + pp.X.nbytes + pp.uv.nbytes + pp.frame_offsets.nbytes
    d2h = pp.R.nbytes + pp.t.nbytes + nM + nF * (max_iter + 1) * 8 + nF * max_iter

    b_po = 21 * nM + 96 * nF                                   # SURVEY 8d-D4, one pass over the matches
    peak, peak_src = hbm_peak()
    gbs = b_po / (k_ms * 1e-3) / 1e9
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "lm_iterations_per_step": max_iter,
                       "parallelism": "single GPU" if world == 1 else f"{world} independent batches, one per GPU (replicas)",
                       "l2": "flushed between steps (256 MiB write); inside a step the batch stays L2-resident"},
            "frames_per_s": value / max_iter,
            "kernel_ms_per_step": {"pose_only_lm": k_ms},
            "roofline": {"bound": "hbm", "kernel": "k_pose_only (whole LM loop, one CTA per frame)", "achieved": gbs,
                         "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_po, "launch_ms": k_ms,
                         "note": f"the matches are read once from HBM and then {2 * max_iter} times from L2 / L1 (linearise + candidate "
                                 f"cost per LM iteration): the kernel is bound by FP64 latency, not by HBM"},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "LocalBASolver.solve_pose_only(inplace, pinned host buffers) -> ba_solve_pose_only_host"},
            "gpu_launches": args.steps, "clocks": clocks}
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import lm_oracle as O
        n_s = 16
        a = int(pp.frame_offsets[n_s])
        t_c = time.perf_counter()
        ref = O.solve_pose_only(pp.intr, pp.R[:n_s], pp.t[:n_s], pp.X[:a], pp.uv[:a], pp.frame_offsets[:n_s + 1],
                                O.LMParams(max_iter=max_iter))
        dt = time.perf_counter() - t_c
        line["oracle_cpu"] = {"value": n_s * max_iter / dt, "unit": unit, "cores": 1,
                              "sample": f"{n_s} of {nF} frames, oracle/lm_oracle.py solve_pose_only (numpy float64) in {dt:.1f} s"}
        try:
            import cv2
            cv2.setRNGSeed(42)
            n_r = 48
            t_c = time.perf_counter()
            for f in range(n_r):
                fa, fb = pp.frame_offsets[f], pp.frame_offsets[f + 1]
                reference_pnp(pp.X[fa:fb], pp.uv[fa:fb], K)
            dt_r = time.perf_counter() - t_c
            line["cpu_baseline"] = {"value": n_r / dt_r * max_iter, "unit": unit, "cores": int(cv2.getNumThreads()),
                                    "kind": "reference",
                                    "sample": f"{n_r} of {nF} frames, cv2 {cv2.__version__} solvePnPRansac as called at "
                                              f"motion_estimation.py:107-116 ({1e3 * dt_r / n_r:.2f} ms per frame); value = frames/s x {max_iter}"}
        except Exception as e:  # OpenCV missing on the box: fall back to the oracle port
            line["cpu_baseline"] = dict(line["oracle_cpu"], kind="port", note=f"cv2 unavailable: {e!r}")
        line["parity_vs_oracle"] = {
            "max_rel_cost_diff": float(np.max(np.abs(out["cost_hist"][:n_s] - ref["cost_hist"]) / ref["cost_hist"])),
            "flags_equal": bool(np.array_equal(out["flags"][:a], ref["flags"]))}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def workload_name(cfg, p):
    return f"{cfg}: local BA window, {p.n_cam} keyframes, {p.n_pt} map points, {p.n_obs} observations, pinhole, Huber 5px"


# --------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import b200ba  # noqa: F401
    from visual_slam_core_b200 import LMConfig, LocalBASolver, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solver has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        with _stdout_to_stderr():          # NCCL prints its version banner on stdout; the contract is ONE JSON line
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    max_iter = args.lm_iters
    sharded = args.shard or args.config in ("C4", "C5")
    if sharded:
        # BASELINE configs 4/5: ONE problem, points sharded over the ranks, cameras replicated; the
        # reduced camera system and the costs are NCCL all-reduced inside the CUDA library
        from visual_slam_core_b200 import dist as badist
        full = synth.make_problem(args.config, seed=0)
        p = full.shard_by_point(rank, world) if world > 1 else full
        n_obs_total, n_pt_total = full.n_obs, full.n_pt
        del full
    else:
        p = synth.make_problem(args.config, seed=rank)      # one independent window per GPU
        n_obs_total, n_pt_total = p.n_obs, p.n_pt
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    R0, t0, X0 = T(p.R), T(p.t), T(p.X)
    R, t, X = R0.clone(), t0.clone(), X0.clone()
    fixed, oc, op, uv, off = T(p.fixed), T(p.obs_cam), T(p.obs_pt), T(p.obs_uv), T(p.pt_offsets)
    flags = torch.zeros(p.n_obs, dtype=torch.uint8, device=dev)
    lmcfg = dict(max_iter=max_iter, schur=args.schur, pcg_tol=args.pcg_tol, pcg_max_iter=args.pcg_max_iter)
    solver = LocalBASolver(LMConfig(**lmcfg), device=local_rank)
    if sharded and world > 1:
        with _stdout_to_stderr():
            badist.init_solver(solver, rank, world, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    if sharded:
        args.steps = min(args.steps, 3)
    stream = torch.cuda.current_stream(dev)

    def step(sync=False):
        R.copy_(R0); t.copy_(t0); X.copy_(X0)
        return solver.solve_local_device(p.intr, R, t, fixed, X, oc, op, uv, off, flags,
                                         stream=stream.cuda_stream, sync=sync)

    rep = None
    for _ in range(max(args.warmup, 3)):
        rep = step(sync=True)
    launches_per_step = int(rep.n_launches)

    sampler = ClockSampler(local_rank)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    for a, b in evs:
        flush.zero_()
        a.record(stream)
        step(sync=False)
        b.record(stream)
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop()
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    n_windows = 1 if sharded else world
    value = n_windows * args.steps * max_iter / (ms_max / 1e3)

    # ---- end to end through the host-buffer API (pinned host memory, copies inside the timed region)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).copy()).pin_memory().numpy()
    hR, ht, hX = pin(p.R), pin(p.t), pin(p.X)
    hfixed, hoc, hop, huv, hoff = pin(p.fixed), pin(p.obs_cam), pin(p.obs_pt), pin(p.obs_uv), pin(p.pt_offsets)
    hflags = pin(np.zeros(p.n_obs, dtype=np.uint8))
    if sharded and world > 1:
        e2e_solver = solver                                  # one communicator per process is enough
    else:
        e2e_solver = LocalBASolver(LMConfig(**lmcfg), device=local_rank)

    def e2e_step():
        np.copyto(hR, p.R); np.copyto(ht, p.t); np.copyto(hX, p.X)
        return e2e_solver.solve_local_inplace(p.intr, hR, ht, hfixed, hX, hoc, hop, huv, hoff, hflags)

    for _ in range(max(args.warmup, 3)):
        res = e2e_step()
    barrier()
    t_a = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    torch.cuda.synchronize(dev)
    t_b = time.perf_counter()
    te = torch.tensor([t_b - t_a], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_windows * args.steps * max_iter / float(te.item())
    import ctypes as C
    from visual_slam_core_b200 import _lib
    h2d = sum(a.nbytes for a in (hR, ht, hfixed, hX, hoc, huv, hoff))   # obs_pt is expanded from pt_offsets on the device
    d2h = hR.nbytes + ht.nbytes + hX.nbytes + hflags.nbytes + 2240   # + LM control block (report)
    final_cost = float(res.cost_hist[-1])

    # ---- profile pass: CUDA events around every kernel (no graph), same workload
    if sharded and world > 1:
        prof = solver
        prof.set_config(LMConfig(profile=True, **lmcfg))
    else:
        prof = LocalBASolver(LMConfig(profile=True, **lmcfg), device=local_rank)
    acc = np.zeros(5)
    n_prof = 5 if not sharded else 2
    for i in range(n_prof + 2):
        R.copy_(R0); t.copy_(t0); X.copy_(X0)
        r = prof.solve_local_device(p.intr, R, t, fixed, X, oc, op, uv, off, flags, stream=stream.cuda_stream, sync=True)
        if i >= 2:
            acc += np.array([r.ms_linearize, r.ms_solve, r.ms_backsub, r.ms_other, r.ms_comm])
    acc /= n_prof
    pcg_iters = [int(v) for v in r.pcg_iters[:max_iter]]
    k_ms = {"linearize_schur": acc[0] / max_iter, "rcs_solve": acc[1] / max_iter,
            "backsub_cost": acc[2] / max_iter, "init_finalize_per_step": acc[3],
            "nccl_allreduce": acc[4] / max_iter}
    b_rjs, b_iter = algorithmic_bytes(p.n_cam, p.n_pt, p.n_obs, dense=p.n_cam <= 38, nnzb=33 * p.n_cam)
    peak, peak_src = hbm_peak()
    dom = max(("linearize_schur", "rcs_solve", "backsub_cost"), key=lambda k: k_ms[k])
    lin_gbs = b_rjs / (k_ms["linearize_schur"] * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(args.config, {}).get("linearize_schur")
    except Exception:
        pass
    k_obs = np.diff(p.pt_offsets).astype(np.float64)
    flops = 300.0 * p.n_obs + 216.0 * float((k_obs * k_obs).sum())       # SURVEY 8d-D4 "algorithmic flops"
    fp64_peak = 36.6                                                     # TFLOP/s, scripts/fp64_peak.cu on this pool's B200
    fp64_tf = flops / (k_ms["linearize_schur"] * 1e-3) / 1e12
    resident = b_iter < 100e6
    roofline = {"bound": "hbm", "kernel": "k_linearize (residual+Jacobian+Schur)", "achieved": lin_gbs, "peak": peak,
                "unit": "GB/s", "frac": lin_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_rjs, "launch_ms": k_ms["linearize_schur"],
                "dominant_kernel_by_time": dom,
                "whole_iteration": {"algorithmic_bytes": b_iter,
                                    "achieved": b_iter / (ms_max / 1e3 / (args.steps * max_iter)) / 1e9,
                                    "frac": b_iter / (ms_max / 1e3 / (args.steps * max_iter)) / 1e9 / peak},
                "fp64": {"algorithmic_flop_per_launch": flops, "achieved_tflops": fp64_tf, "peak_tflops": fp64_peak,
                         "frac": fp64_tf / fp64_peak,
                         "note": "all arithmetic is float64 (vector FMA + mma.m8n8k4.f64); measured peak of scripts/fp64_peak.cu"},
                "note": ("window is L2-resident (working set << 126 MB L2): FP64-latency bound, see DESIGN.md" if resident else
                         "float64 arithmetic puts the kernel's floor at the FP64 pipe, not at HBM: see roofline.fp64 and DESIGN.md")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, p) if not sharded else
                   f"{args.config}: {p.n_cam} cameras, {n_pt_total} points, {n_obs_total} observations, sharded by point over {world} GPU(s)",
                   "lm_iterations_per_step": max_iter,
                   "parallelism": "single GPU" if world == 1 else (
                       f"points sharded over {world} GPUs, NCCL all-reduce of the reduced camera system" if sharded
                       else f"{world} independent windows, one per GPU (replicas)"),
                   "pcg": {"tol": args.pcg_tol, "iterations": pcg_iters} if p.n_cam > 38 else None,
                   "l2": "flushed between steps (256 MiB write)" + ("; inside a step the window stays L2-resident by construction" if resident else ""),
                   "schur": args.schur},
        "obs_per_s_residual_jacobian_schur": p.n_obs / (k_ms["linearize_schur"] * 1e-3) * world,
        "shard_observations": p.n_obs,
        "kernel_ms_per_lm_iteration": k_ms,
        "final_cost": final_cost,
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "LocalBASolver.solve_local_inplace -> ba_solve_local_host (pinned host buffers)"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_cpu and p.n_obs <= 400_000:
        from oracle import lm_oracle as O
        t_c = time.perf_counter()
        ref = O.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv,
                            O.LMParams(max_iter=max_iter))
        dt = time.perf_counter() - t_c
        rel = float(np.max(np.abs(res.cost_hist - ref.cost_hist) / ref.cost_hist))
        line["cpu_baseline"] = {"value": max_iter / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"the full window, one solve of {max_iter} LM iterations with "
                                          f"oracle/lm_oracle.py (numpy float64, same algorithm) in {dt:.1f} s"}
        line["parity_vs_oracle"] = {"max_rel_cost_diff": rel,
                                    "accept_sequence_equal": bool(np.array_equal(res.accepted, ref.accepted)),
                                    "flags_equal": bool(np.array_equal(res.flags, ref.flags))}
    if rank == 0 and world == 1 and not args.no_cpu and p.n_obs <= 400_000:
        try:
            line["plugin"] = plugin_leg(p, local_rank, max_iter)
        except Exception as e:                              # informational leg: never costs the bench line
            line["plugin"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--lm-iters", type=int, default=10)
    ap.add_argument("--schur", default="auto")
    ap.add_argument("--shard", action="store_true", help="shard ONE problem by point over the ranks (default for C4/C5)")
    ap.add_argument("--pcg-tol", type=float, default=1e-10)
    ap.add_argument("--pcg-max-iter", type=int, default=3000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config == "C3":
        run_pose_only(args, rank, local_rank, world)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
