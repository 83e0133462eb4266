"""Synthetic bundle-adjustment problems.

One generator for every BASELINE.json config.  The output is the packed
structure-of-arrays form that the reference's packing loop
(`/root/reference/visual_slam/optimization/torch_optimizer.py:43-95`) would
produce for the equivalent KeyFrame/MapPoint graph: observations ordered by
map point (CSR offsets), poses world->camera, camera 0 fixed.

All random draws come from ``numpy.random.default_rng(seed)`` (PCG64).
Poses and points are rounded through float32 once, exactly as the reference
does when it builds its tensors (`torch_optimizer.py:51-52,65-68`), so the
values handed to the solver are float32-representable.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

# pinhole 1280x720, fx = fy = 700 (notebook frames)
WIDTH, HEIGHT = 1280.0, 720.0
INTR = np.array([700.0, 700.0, 640.0, 360.0], dtype=np.float64)

CONFIGS: Dict[str, dict] = {
    # name: n_cam, n_pt, k_mean, k_max, full (all-see-all)
    "C1": dict(n_cam=10, n_pt=2_000, k_mean=None, k_max=10, full=True),
    "C2": dict(n_cam=20, n_pt=10_000, k_mean=15.0, k_max=20, full=False),
    "C4": dict(n_cam=1_000, n_pt=1_000_000, k_mean=8.0, k_max=16, full=False),
    "C5": dict(n_cam=10_000, n_pt=4_000_000, k_mean=7.5, k_max=16, full=False),
    # small cases for parity tests
    "T0": dict(n_cam=4, n_pt=60, k_mean=3.0, k_max=4, full=False),
    "T1": dict(n_cam=6, n_pt=300, k_mean=4.0, k_max=6, full=False),
    "T2": dict(n_cam=12, n_pt=1_500, k_mean=8.0, k_max=12, full=False),
    "T3": dict(n_cam=40, n_pt=4_000, k_mean=6.0, k_max=12, full=False),
    "T4": dict(n_cam=200, n_pt=20_000, k_mean=6.0, k_max=12, full=False),
}


@dataclass
class BAProblem:
    """Packed local-BA problem (host arrays)."""

    intr: np.ndarray          # (4,) f64  fx fy cx cy
    R: np.ndarray             # (n_cam,3,3) f64 world->camera
    t: np.ndarray             # (n_cam,3) f64
    fixed: np.ndarray         # (n_cam,) u8
    X: np.ndarray             # (n_pt,3) f64, float32-representable
    obs_cam: np.ndarray       # (n_obs,) i32
    obs_pt: np.ndarray        # (n_obs,) i32, non-decreasing
    obs_uv: np.ndarray        # (n_obs,2) f32
    pt_offsets: np.ndarray    # (n_pt+1,) i32 CSR over obs
    R_true: Optional[np.ndarray] = None
    t_true: Optional[np.ndarray] = None
    X_true: Optional[np.ndarray] = None
    outlier: Optional[np.ndarray] = None   # (n_obs,) bool, injected outliers
    meta: dict = field(default_factory=dict)

    @property
    def n_cam(self) -> int:
        return int(self.R.shape[0])

    @property
    def n_pt(self) -> int:
        return int(self.X.shape[0])

    @property
    def n_obs(self) -> int:
        return int(self.obs_cam.shape[0])

    def copy(self) -> "BAProblem":
        return BAProblem(
            self.intr.copy(), self.R.copy(), self.t.copy(), self.fixed.copy(),
            self.X.copy(), self.obs_cam.copy(), self.obs_pt.copy(),
            self.obs_uv.copy(), self.pt_offsets.copy(),
            None if self.R_true is None else self.R_true.copy(),
            None if self.t_true is None else self.t_true.copy(),
            None if self.X_true is None else self.X_true.copy(),
            None if self.outlier is None else self.outlier.copy(),
            dict(self.meta),
        )

    def shard_by_point(self, rank: int, world: int) -> "BAProblem":
        """Contiguous point range balanced by observation count.

        Cameras are replicated; the shard keeps every observation of its points.
        """
        lo, hi = shard_point_range(self.pt_offsets, rank, world)
        o0, o1 = int(self.pt_offsets[lo]), int(self.pt_offsets[hi])
        return BAProblem(
            self.intr.copy(), self.R.copy(), self.t.copy(), self.fixed.copy(),
            self.X[lo:hi].copy(), self.obs_cam[o0:o1].copy(),
            (self.obs_pt[o0:o1] - lo).astype(np.int32), self.obs_uv[o0:o1].copy(),
            (self.pt_offsets[lo:hi + 1] - o0).astype(np.int32),
            meta=dict(self.meta, shard=(rank, world), pt_range=(lo, hi)),
        )


def shard_point_range(pt_offsets: np.ndarray, rank: int, world: int):
    """Point range [lo, hi) of `rank`: cut the CSR at equal observation counts."""
    n_pt = len(pt_offsets) - 1
    n_obs = int(pt_offsets[-1])
    cuts = [0]
    for r in range(1, world):
        target = (n_obs * r) // world
        cuts.append(int(np.searchsorted(pt_offsets, target, side="left")))
    cuts.append(n_pt)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n_pt))
    return int(cuts[rank]), int(cuts[rank + 1])


def so3_exp(w: np.ndarray) -> np.ndarray:
    """Rodrigues formula, batched: w (...,3) -> R (...,3,3).

    Uses sin(th)/th and (1-cos th)/th^2 = 0.5*(sin(th/2)/(th/2))^2, both
    evaluated through a sinc with a Taylor branch below 1e-4, so it is stable
    at th -> 0 (the reference's `so3_exp`, torch_optimizer.py:119-133, adds
    1e-8 to th instead).
    """
    w = np.asarray(w, dtype=np.float64)
    th2 = np.sum(w * w, axis=-1)
    th = np.sqrt(th2)

    def sinc(x):
        x2 = x * x
        small = x < 1e-4
        xs = np.where(small, 1.0, x)
        return np.where(small, 1.0 - x2 / 6.0 + x2 * x2 / 120.0, np.sin(xs) / xs)

    A = sinc(th)
    h = sinc(0.5 * th)
    B = 0.5 * h * h
    wx, wy, wz = w[..., 0], w[..., 1], w[..., 2]
    Kw = np.zeros(w.shape[:-1] + (3, 3))
    Kw[..., 0, 1] = -wz
    Kw[..., 0, 2] = wy
    Kw[..., 1, 0] = wz
    Kw[..., 1, 2] = -wx
    Kw[..., 2, 0] = -wy
    Kw[..., 2, 1] = wx
    K2 = Kw @ Kw
    return np.eye(3) + A[..., None, None] * Kw + B[..., None, None] * K2


def trajectory(n_cam: int):
    """Camera i: centre (0.4 i, 0, 0.05 i), yaw 0.01 i about +y; world->camera."""
    i = np.arange(n_cam, dtype=np.float64)
    c = np.stack([0.4 * i, np.zeros(n_cam), 0.05 * i], axis=1)
    yaw = 0.01 * i
    cy, sy = np.cos(yaw), np.sin(yaw)
    R_c2w = np.zeros((n_cam, 3, 3))
    R_c2w[:, 0, 0] = cy
    R_c2w[:, 0, 2] = sy
    R_c2w[:, 1, 1] = 1.0
    R_c2w[:, 2, 0] = -sy
    R_c2w[:, 2, 2] = cy
    R = np.transpose(R_c2w, (0, 2, 1))
    t = -np.einsum("nij,nj->ni", R, c)
    return R, t, c


def _project(R, t, Xw, intr):
    Xc = np.einsum("...ij,...j->...i", R, Xw) + t
    z = Xc[..., 2]
    zs = np.where(np.abs(z) < 1e-12, 1e-12, z)
    u = intr[0] * Xc[..., 0] / zs + intr[2]
    v = intr[1] * Xc[..., 1] / zs + intr[3]
    return u, v, z


def _sample_points(rng, anchors, R, t):
    """Points in the anchor camera's frustum: depth U[6,30], lateral +-0.4/0.25 depth."""
    n = anchors.shape[0]
    depth = rng.uniform(6.0, 30.0, n)
    xl = rng.uniform(-0.4, 0.4, n) * depth
    yl = rng.uniform(-0.25, 0.25, n) * depth
    Xc = np.stack([xl, yl, depth], axis=1)
    Ra, ta = R[anchors], t[anchors]
    # Xw = R^T (Xc - t)
    return np.einsum("nji,nj->ni", Ra, Xc - ta)


def make_problem(
    name: Optional[str] = None,
    *,
    n_cam: Optional[int] = None,
    n_pt: Optional[int] = None,
    k_mean: Optional[float] = None,
    k_max: Optional[int] = None,
    full: bool = False,
    seed: int = 0,
    outlier_frac: float = 0.05,
    noise_px: float = 1.0,
    chunk: int = 250_000,
    sort_points: Optional[bool] = None,
    init_scale: float = 1.0,
) -> BAProblem:
    """Generate one of the named configs (or a custom size).

    `sort_points` (default: on when n_cam > 40) orders the points by their anchor camera, the
    order a SLAM map has naturally: `Map.optimize_local` collects points keyframe by keyframe
    (map/map.py:159-169), so consecutive points are seen by neighbouring keyframes.
    `init_scale` scales the perturbation of the initial estimate (1.0 = SURVEY 8d-D2: 0.05 m / 0.01 rad).
    """
    if name is not None:
        cfg = dict(CONFIGS[name])
        n_cam = cfg["n_cam"] if n_cam is None else n_cam
        n_pt = cfg["n_pt"] if n_pt is None else n_pt
        k_mean = cfg["k_mean"] if k_mean is None else k_mean
        k_max = cfg["k_max"] if k_max is None else k_max
        full = cfg["full"]
    assert n_cam is not None and n_pt is not None
    k_max = min(int(k_max or n_cam), n_cam)
    rng = np.random.default_rng(seed)
    R_true, t_true, centres = trajectory(n_cam)

    win = n_cam if full else min(n_cam, 2 * k_max + 1)
    X_true = np.empty((n_pt, 3))
    counts = np.zeros(n_pt, dtype=np.int64)
    anchor_all = np.zeros(n_pt, dtype=np.int64)
    cam_lists = []   # per chunk: (k_per_point, cams flat)

    for p0 in range(0, n_pt, chunk):
        p1 = min(n_pt, p0 + chunk)
        m = p1 - p0
        anchors = rng.integers(0, n_cam, m)
        if full:
            k_target = np.full(m, n_cam)
        else:
            k_target = np.clip(rng.poisson(k_mean, m), 2, k_max)
        Xw = _sample_points(rng, anchors, R_true, t_true)
        need = np.ones(m, dtype=bool)
        sel_mask = np.zeros((m, win), dtype=bool)
        cand_all = np.zeros((m, win), dtype=np.int64)
        for _attempt in range(200):
            idx = np.nonzero(need)[0]
            if idx.size == 0:
                break
            a = anchors[idx]
            # candidate cameras ordered by |i - a| (ties: lower index first)
            lo = np.clip(a - win // 2, 0, n_cam - win)
            cand = lo[:, None] + np.arange(win)[None, :]
            dist = np.abs(cand - a[:, None]) * 2 + (cand > a[:, None])
            order = np.argsort(dist, axis=1, kind="stable")
            cand = np.take_along_axis(cand, order, axis=1)
            u, v, z = _project(R_true[cand], t_true[cand], Xw[idx][:, None, :], INTR)
            vis = (z > 1.0) & (u >= 0) & (u < WIDTH) & (v >= 0) & (v < HEIGHT)
            rank_vis = np.cumsum(vis, axis=1)
            sel = vis & (rank_vis <= k_target[idx][:, None])
            nsel = sel.sum(axis=1)
            ok = (nsel == n_cam) if full else (nsel >= 2)
            good = idx[ok]
            sel_mask[good] = sel[ok]
            cand_all[good] = cand[ok]
            need[good] = False
            bad = idx[~ok]
            if bad.size:
                anchors[bad] = rng.integers(0, n_cam, bad.size)
                Xw[bad] = _sample_points(rng, anchors[bad], R_true, t_true)
        else:
            raise RuntimeError("could not place all points")
        X_true[p0:p1] = Xw
        anchor_all[p0:p1] = anchors
        k = sel_mask.sum(axis=1)
        counts[p0:p1] = k
        # selected cameras, per point sorted ascending by camera index
        cams = np.where(sel_mask, cand_all, n_cam + 1)
        cams.sort(axis=1)
        cam_lists.append(cams[cams <= n_cam])

    pt_offsets = np.zeros(n_pt + 1, dtype=np.int64)
    np.cumsum(counts, out=pt_offsets[1:])
    n_obs = int(pt_offsets[-1])
    assert n_obs < 2**31
    obs_cam = np.concatenate(cam_lists).astype(np.int32)
    if sort_points is None:
        sort_points = n_cam > 40
    if sort_points:
        order = np.argsort(anchor_all, kind="stable")
        new_counts = counts[order]
        new_off = np.zeros(n_pt + 1, dtype=np.int64)
        np.cumsum(new_counts, out=new_off[1:])
        src = np.repeat(pt_offsets[:-1][order] - new_off[:-1], new_counts) + np.arange(n_obs, dtype=np.int64)
        obs_cam = obs_cam[src]
        X_true = X_true[order]
        counts, pt_offsets = new_counts, new_off
    obs_pt = np.repeat(np.arange(n_pt, dtype=np.int32), counts)

    # measurements: exact projection + N(0,1px); 5% get an extra N(0,30px)
    obs_uv = np.empty((n_obs, 2), dtype=np.float32)
    outlier = np.empty(n_obs, dtype=bool)
    ochunk = 4_000_000
    for o0 in range(0, n_obs, ochunk):
        o1 = min(n_obs, o0 + ochunk)
        c, p = obs_cam[o0:o1], obs_pt[o0:o1]
        u, v, _ = _project(R_true[c], t_true[c], X_true[p], INTR)
        uv = np.stack([u, v], axis=1) + rng.normal(0.0, noise_px, (o1 - o0, 2))
        out = rng.random(o1 - o0) < outlier_frac
        uv += out[:, None] * rng.normal(0.0, 30.0, (o1 - o0, 2))
        obs_uv[o0:o1] = uv.astype(np.float32)
        outlier[o0:o1] = out

    # initial estimate
    X0 = X_true + rng.normal(0.0, 0.05 * init_scale, (n_pt, 3))
    t0 = t_true + rng.normal(0.0, 0.05 * init_scale, (n_cam, 3))
    dR = so3_exp(rng.normal(0.0, 0.01 * init_scale, (n_cam, 3)))
    R0 = dR @ R_true
    R0[0], t0[0] = R_true[0], t_true[0]
    fixed = np.zeros(n_cam, dtype=np.uint8)
    fixed[0] = 1

    return BAProblem(
        intr=INTR.copy(),
        R=R0.astype(np.float32).astype(np.float64),
        t=t0.astype(np.float32).astype(np.float64),
        fixed=fixed,
        X=X0.astype(np.float32).astype(np.float64),
        obs_cam=obs_cam, obs_pt=obs_pt, obs_uv=obs_uv,
        pt_offsets=pt_offsets.astype(np.int32),
        R_true=R_true, t_true=t_true, X_true=X_true, outlier=outlier,
        meta=dict(name=name or "custom", seed=seed, n_cam=n_cam, n_pt=n_pt, n_obs=n_obs),
    )


@dataclass
class PoseOnlyProblem:
    """Batched pose-only (motion-only) refinement: frames x 2D-3D matches (config 3)."""

    intr: np.ndarray           # (4,)
    R: np.ndarray              # (n_frames,3,3) f64 initial pose
    t: np.ndarray              # (n_frames,3)
    X: np.ndarray              # (n_match,3) f32 fixed 3D points (world)
    uv: np.ndarray             # (n_match,2) f32
    frame_offsets: np.ndarray  # (n_frames+1,) i32
    R_true: Optional[np.ndarray] = None
    t_true: Optional[np.ndarray] = None
    outlier: Optional[np.ndarray] = None
    meta: dict = field(default_factory=dict)

    @property
    def n_frames(self) -> int:
        return int(self.R.shape[0])

    @property
    def n_match(self) -> int:
        return int(self.X.shape[0])


def make_pose_only(
    n_frames: int = 256,
    n_match: int = 1000,
    seed: int = 0,
    outlier_frac: float = 0.10,
    ragged: bool = False,
) -> PoseOnlyProblem:
    """SURVEY §8d-D2 'C3': independent frames, 10% outliers, perturbed initial pose."""
    rng = np.random.default_rng(seed)
    R_true, t_true, _ = trajectory(n_frames)
    if ragged:
        counts = rng.integers(max(4, n_match // 4), n_match + 1, n_frames)
    else:
        counts = np.full(n_frames, n_match)
    off = np.zeros(n_frames + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    frame_of = np.repeat(np.arange(n_frames), counts)
    Xw = _sample_points(rng, frame_of, R_true, t_true)
    u, v, _ = _project(R_true[frame_of], t_true[frame_of], Xw, INTR)
    uv = np.stack([u, v], axis=1) + rng.normal(0.0, 1.0, (n, 2))
    out = rng.random(n) < outlier_frac
    uv += out[:, None] * rng.normal(0.0, 30.0, (n, 2))
    X = Xw + rng.normal(0.0, 0.02, (n, 3))
    dR = so3_exp(rng.normal(0.0, 0.02, (n_frames, 3)))
    R0 = dR @ R_true
    t0 = t_true + rng.normal(0.0, 0.1, (n_frames, 3))
    return PoseOnlyProblem(
        intr=INTR.copy(),
        R=R0.astype(np.float32).astype(np.float64),
        t=t0.astype(np.float32).astype(np.float64),
        X=X.astype(np.float32), uv=uv.astype(np.float32),
        frame_offsets=off.astype(np.int32),
        R_true=R_true, t_true=t_true, outlier=out,
        meta=dict(name="C3", seed=seed, n_frames=n_frames, n_match=n),
    )
