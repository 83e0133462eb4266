"""Host-side driver of the CUDA library: owns a `ba_handle`, feeds it packed SoA arrays.

`LocalBASolver.solve_local(...)` is the packed-array public API (host numpy buffers in, refined
numpy buffers out, host<->device copies inside the call).  `solve_local_device(...)` takes torch
CUDA tensors (PyTorch is used only for device memory and the stream) and refines them in place.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib


@dataclass
class LMConfig:
    """LM parameters.  Read from `config.optimization` with getattr defaults."""

    huber_delta: float = 5.0
    chi2_thr: float = 5.991
    lambda0: float = 1e-4
    lambda_up: float = 10.0
    lambda_down: float = 0.1
    lambda_min: float = 1e-6
    lambda_max: float = 1e8
    accept_rtol: float = 1e-8
    max_iter: int = 10
    solver: str = "auto"        # auto | cholesky | pcg | band | band_seq
    schur: str = "auto"         # auto | scatter | tiled | stream | mma | panel
    pcg_max_iter: int = 500
    pcg_tol: float = 1e-10
    use_graph: bool = True
    profile: bool = False       # events around every kernel (disables the graph)

    def to_c(self) -> "_lib.BAParams":
        p = _lib.BAParams()
        p.huber_delta, p.chi2_thr, p.lambda0 = self.huber_delta, self.chi2_thr, self.lambda0
        p.lambda_up, p.lambda_down = self.lambda_up, self.lambda_down
        p.lambda_min, p.lambda_max = self.lambda_min, self.lambda_max
        p.accept_rtol = self.accept_rtol
        p.max_iter = int(self.max_iter)
        p.solver_kind = {"auto": 0, "cholesky": 1, "pcg": 2, "band": 3, "band_seq": 4}[self.solver]
        p.schur_kind = {"auto": 0, "scatter": 1, "tiled": 2, "stream": 3, "mma": 4, "panel": 5}[self.schur]
        p.pcg_max_iter, p.pcg_tol = int(self.pcg_max_iter), float(self.pcg_tol)
        p.use_graph = 1 if self.use_graph else 0
        p.profile = 1 if self.profile else 0
        return p


@dataclass
class SolveResult:
    R: np.ndarray
    t: np.ndarray
    X: np.ndarray
    flags: np.ndarray
    cost_hist: np.ndarray
    cand_cost_hist: np.ndarray
    lambda_hist: np.ndarray
    accepted: np.ndarray
    solve_fail: np.ndarray
    pcg_iters: np.ndarray
    n_iter: int
    gpu_ms: float
    n_launches: int
    extra: dict = field(default_factory=dict)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class LocalBASolver:
    """One CUDA handle (workspace, stream, graph cache).  Not re-entrant: calls are serialised."""

    def __init__(self, cfg: Optional[LMConfig] = None, device: int = 0):
        self.cfg = cfg or LMConfig()
        self.lib = _lib.load()
        self.device = int(device)
        self._h = C.c_void_p()
        self._lock = threading.Lock()
        prm = self.cfg.to_c()
        rc = self.lib.ba_create(C.byref(prm), self.device, C.byref(self._h))
        _lib.check(self.lib, None, rc, "ba_create")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.ba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dist_init(self, nccl_id: bytes, rank: int, world: int):
        """Join the shard-by-point group: every rank then calls the same solve on its point shard."""
        buf = (C.c_ubyte * len(nccl_id)).from_buffer_copy(nccl_id)
        rc = self.lib.ba_dist_init(self._h, buf, int(rank), int(world))
        _lib.check(self.lib, self._h, rc, "ba_dist_init")

    def set_config(self, cfg: LMConfig):
        with self._lock:
            self.cfg = cfg
            prm = cfg.to_c()
            _lib.check(self.lib, self._h, self.lib.ba_set_params(self._h, C.byref(prm)), "ba_set_params")

    # ------------------------------------------------------------------ local BA, host buffers
    def solve_local(self, intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, pt_offsets,
                    want_flags: bool = True) -> SolveResult:
        intr = np.ascontiguousarray(intr, dtype=np.float64)
        R = np.array(R, dtype=np.float64, order="C").reshape(-1, 3, 3)
        t = np.array(t, dtype=np.float64, order="C").reshape(-1, 3)
        X = np.array(X, dtype=np.float64, order="C").reshape(-1, 3)
        fixed = np.ascontiguousarray(fixed, dtype=np.uint8)
        obs_cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
        obs_pt = np.ascontiguousarray(obs_pt, dtype=np.int32)
        obs_uv = np.ascontiguousarray(obs_uv, dtype=np.float32).reshape(-1, 2)
        pt_offsets = np.ascontiguousarray(pt_offsets, dtype=np.int32)
        n_cam, n_pt, n_obs = R.shape[0], X.shape[0], obs_cam.shape[0]
        if t.shape[0] != n_cam or fixed.shape[0] != n_cam or pt_offsets.shape[0] != n_pt + 1 \
                or obs_pt.shape[0] != n_obs or obs_uv.shape[0] != n_obs:
            raise ValueError("inconsistent array sizes")
        flags = np.zeros(n_obs, dtype=np.uint8) if want_flags else None
        rep = _lib.BAReport()
        with self._lock:
            rc = self.lib.ba_solve_local_host(
                self._h, n_cam, n_pt, n_obs, _dptr(intr), _ptr(R), _ptr(t), _ptr(fixed), _ptr(X),
                _ptr(obs_cam), _ptr(obs_pt), _ptr(obs_uv), _ptr(pt_offsets), _ptr(flags),
                C.byref(rep))
            _lib.check(self.lib, self._h, rc, "ba_solve_local_host")
        return self._result(rep, R, t, X, flags)

    def solve_local_inplace(self, intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, pt_offsets,
                            flags=None) -> SolveResult:
        """Same as `solve_local` but refines the given host arrays in place (no defensive copies):
        R,t,X float64 C-contiguous, uv float32, indices int32, fixed/flags uint8.  Hand it pinned
        buffers and the host<->device copies inside the call are DMA transfers."""
        for name, a, dt in (("R", R, np.float64), ("t", t, np.float64), ("X", X, np.float64),
                            ("obs_uv", obs_uv, np.float32), ("obs_cam", obs_cam, np.int32),
                            ("obs_pt", obs_pt, np.int32), ("pt_offsets", pt_offsets, np.int32),
                            ("fixed", fixed, np.uint8)):
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"{name} must be C-contiguous {dt}")
        intr = np.ascontiguousarray(intr, dtype=np.float64)
        n_cam, n_pt, n_obs = R.shape[0], X.shape[0], obs_cam.shape[0]
        rep = _lib.BAReport()
        with self._lock:
            rc = self.lib.ba_solve_local_host(
                self._h, n_cam, n_pt, n_obs, _dptr(intr), _ptr(R), _ptr(t), _ptr(fixed), _ptr(X),
                _ptr(obs_cam), _ptr(obs_pt), _ptr(obs_uv), _ptr(pt_offsets), _ptr(flags),
                C.byref(rep))
            _lib.check(self.lib, self._h, rc, "ba_solve_local_host")
        return self._result(rep, R, t, X, flags)

    def _result(self, rep, R, t, X, flags) -> SolveResult:
        k = int(rep.n_iter)
        return SolveResult(
            R=R, t=t, X=X, flags=flags,
            cost_hist=np.array(rep.cost_hist[: k + 1]),
            cand_cost_hist=np.array(rep.cand_cost_hist[:k]),
            lambda_hist=np.array(rep.lambda_hist[:k]),
            accepted=np.array(rep.accepted[:k], dtype=np.uint8),
            solve_fail=np.array(rep.solve_fail[:k], dtype=np.int32),
            pcg_iters=np.array(rep.pcg_iters[:k], dtype=np.int32),
            n_iter=k, gpu_ms=float(rep.gpu_ms), n_launches=int(rep.n_launches),
            extra=dict(ms_linearize=float(rep.ms_linearize), ms_solve=float(rep.ms_solve),
                       ms_backsub=float(rep.ms_backsub), ms_other=float(rep.ms_other), ms_comm=float(rep.ms_comm),
                       ms_point_blocks=float(rep.ms_point_blocks), ms_cam_blocks=float(rep.ms_cam_blocks),
                       solver_used={1: "cholesky", 2: "pcg", 3: "band"}.get(int(rep.solver_used), "?"),
                       band_width=int(rep.band_width), s_blocks=int(rep.s_blocks),
                       peer_exchange=int(rep.peer_exchange)))

    # ------------------------------------------------------------------ local BA, torch CUDA tensors
    def solve_local_device(self, intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, pt_offsets,
                           flags=None, stream: Optional[int] = None, sync: bool = True):
        """In-place refinement of torch CUDA tensors (float64 R,t,X; float32 uv; int32 indices)."""
        import torch
        for name, ten, dt in (("R", R, torch.float64), ("t", t, torch.float64), ("X", X, torch.float64),
                              ("obs_uv", obs_uv, torch.float32), ("obs_cam", obs_cam, torch.int32),
                              ("obs_pt", obs_pt, torch.int32), ("pt_offsets", pt_offsets, torch.int32),
                              ("fixed", fixed, torch.uint8)):
            if not (ten.is_cuda and ten.is_contiguous() and ten.dtype == dt):
                raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {dt}")
        intr = np.ascontiguousarray(intr, dtype=np.float64)
        n_cam, n_pt, n_obs = R.shape[0], X.shape[0], obs_cam.shape[0]
        if stream is None:
            stream = torch.cuda.current_stream(R.device).cuda_stream
        rep = _lib.BAReport() if sync else None
        with self._lock:
            rc = self.lib.ba_solve_local(
                self._h, n_cam, n_pt, n_obs, _dptr(intr), R.data_ptr(), t.data_ptr(),
                fixed.data_ptr(), X.data_ptr(), obs_cam.data_ptr(), obs_pt.data_ptr(),
                obs_uv.data_ptr(), pt_offsets.data_ptr(),
                None if flags is None else flags.data_ptr(), C.c_void_p(stream),
                C.byref(rep) if sync else None)
            _lib.check(self.lib, self._h, rc, "ba_solve_local")
        return rep

    # ------------------------------------------------------------------ pose-only
    def solve_pose_only(self, intr, R, t, X, uv, frame_offsets, want_hist: bool = True, inplace: bool = False,
                        flags=None, cost_hist=None, accepted=None, chi2_thr: Optional[float] = None):
        """Batched motion-only LM (one free pose per frame, points fixed).

        `inplace=True` refines the caller's float64 C-contiguous `R` (n,3,3) / `t` (n,3) where they are;
        `flags` (uint8 [n_match]), `cost_hist` (float64 [n, max_iter+1]) and `accepted` (uint8 [n, max_iter])
        may be passed in to be filled.  With all of them (and X, uv, frame_offsets) in pinned host memory the
        host<->device copies inside the call are plain DMA transfers.  `chi2_thr` overrides the flag threshold
        for this call only (applied and restored under the solver's lock)."""
        intr = np.ascontiguousarray(intr, dtype=np.float64)
        if inplace:
            for name, a in (("R", R), ("t", t)):
                if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError(f"{name} must be C-contiguous float64 for inplace=True")
            R = R.reshape(-1, 3, 3)
            t = t.reshape(-1, 3)
        else:
            R = np.array(R, dtype=np.float64, order="C").reshape(-1, 3, 3)
            t = np.array(t, dtype=np.float64, order="C").reshape(-1, 3)
        X = np.ascontiguousarray(X, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        fo = np.ascontiguousarray(frame_offsets, dtype=np.int32)
        nF, nM = R.shape[0], X.shape[0]
        if fo.shape[0] != nF + 1 or uv.shape[0] != nM or (nF and int(fo[-1]) != nM):
            raise ValueError("inconsistent array sizes")
        mi = self.cfg.max_iter

        def _out(a, shape, name):
            if a is None:
                return np.zeros(shape, dtype=np.float64 if name == "cost_hist" else np.uint8)
            want = np.float64 if name == "cost_hist" else np.uint8
            if a.dtype != want or not a.flags["C_CONTIGUOUS"] or a.size != int(np.prod(shape)):
                raise ValueError(f"{name} must be a C-contiguous {np.dtype(want).name} array of shape {shape}")
            return a

        flags = _out(flags, (nM,), "flags")
        hist = _out(cost_hist, (nF, mi + 1), "cost_hist") if (want_hist or cost_hist is not None) else None
        acc = _out(accepted, (nF, max(mi, 1)), "accepted") if (want_hist or accepted is not None) else None
        ms = C.c_float(0.0)
        with self._lock:
            swapped = chi2_thr is not None and float(chi2_thr) != self.cfg.chi2_thr
            if swapped:
                import dataclasses
                tmp = dataclasses.replace(self.cfg, chi2_thr=float(chi2_thr)).to_c()
                _lib.check(self.lib, self._h, self.lib.ba_set_params(self._h, C.byref(tmp)), "ba_set_params")
            try:
                rc = self.lib.ba_solve_pose_only_host(
                    self._h, nF, nM, _dptr(intr), _ptr(R), _ptr(t), _ptr(X), _ptr(uv), _ptr(fo),
                    _ptr(flags), _ptr(hist), _ptr(acc), C.byref(ms))
            finally:
                if swapped:
                    old = self.cfg.to_c()
                    self.lib.ba_set_params(self._h, C.byref(old))
            _lib.check(self.lib, self._h, rc, "ba_solve_pose_only_host")
        return dict(R=R, t=t, flags=flags, cost_hist=None if hist is None else hist.reshape(nF, mi + 1),
                    accepted=None if acc is None else acc.reshape(nF, max(mi, 1))[:, :mi], gpu_ms=float(ms.value))

    def pose_ransac(self, intr, X, uv, n_hypotheses: int = 100, threshold_px: float = 3.0, seed: int = 42):
        """Pose of one frame without a prior: RANSAC over minimal P3P samples on the device (`ba_pose_ransac_host`).
        Returns dict(R (3,3), t (3,), n_inliers) or None when no hypothesis is admissible."""
        intr = np.ascontiguousarray(intr, dtype=np.float64)
        X = np.ascontiguousarray(X, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        if X.shape[0] != uv.shape[0] or X.shape[0] < 4:
            return None
        R = np.zeros(9, dtype=np.float64)
        t = np.zeros(3, dtype=np.float64)
        n_in = C.c_int32(0)
        with self._lock:
            rc = self.lib.ba_pose_ransac_host(self._h, X.shape[0], _dptr(intr), _ptr(X), _ptr(uv), int(n_hypotheses),
                                              float(threshold_px), int(seed) & 0xffffffff, _dptr(R), _dptr(t), C.byref(n_in))
            _lib.check(self.lib, self._h, rc, "ba_pose_ransac_host")
        if n_in.value <= 0:
            return None
        return dict(R=R.reshape(3, 3), t=t, n_inliers=int(n_in.value))

    # ------------------------------------------------------------------ inspection (tests)
    def debug_linearize(self, intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, pt_offsets, lam):
        """Run init + one linearise/Schur on device and return (S lower-valid, b, diagU, cost)."""
        import torch
        dev = torch.device("cuda", self.device)
        intr = np.ascontiguousarray(intr, dtype=np.float64)
        tens = dict(
            R=torch.as_tensor(np.ascontiguousarray(R, dtype=np.float64)).to(dev),
            t=torch.as_tensor(np.ascontiguousarray(t, dtype=np.float64)).to(dev),
            fixed=torch.as_tensor(np.ascontiguousarray(fixed, dtype=np.uint8)).to(dev),
            X=torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64)).to(dev),
            oc=torch.as_tensor(np.ascontiguousarray(obs_cam, dtype=np.int32)).to(dev),
            op=torch.as_tensor(np.ascontiguousarray(obs_pt, dtype=np.int32)).to(dev),
            uv=torch.as_tensor(np.ascontiguousarray(obs_uv, dtype=np.float32)).to(dev),
            off=torch.as_tensor(np.ascontiguousarray(pt_offsets, dtype=np.int32)).to(dev),
        )
        n_cam, n_pt, n_obs = tens["R"].shape[0], tens["X"].shape[0], tens["oc"].shape[0]
        n = 6 * n_cam
        Sp = np.zeros(n_cam * (n_cam + 1) // 2 * 36); b = np.zeros(n); dU = np.zeros(n); cost = np.zeros(1)
        stream = torch.cuda.current_stream(dev).cuda_stream
        torch.cuda.synchronize(dev)
        with self._lock:
            rc = self.lib.ba_debug_linearize(
                self._h, n_cam, n_pt, n_obs, _dptr(intr), tens["R"].data_ptr(), tens["t"].data_ptr(),
                tens["fixed"].data_ptr(), tens["X"].data_ptr(), tens["oc"].data_ptr(),
                tens["op"].data_ptr(), tens["uv"].data_ptr(), tens["off"].data_ptr(), float(lam),
                C.c_void_p(stream), _ptr(Sp), _ptr(b), _ptr(dU), _ptr(cost))
            _lib.check(self.lib, self._h, rc, "ba_debug_linearize")
        S = np.zeros((n, n))                 # unpack the block-packed lower triangle
        for i in range(n_cam):
            for j in range(i + 1):
                o = (i * (i + 1) // 2 + j) * 36
                S[6 * i:6 * i + 6, 6 * j:6 * j + 6] = Sp[o:o + 36].reshape(6, 6)
        return S, b, dU, float(cost[0])
