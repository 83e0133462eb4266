"""Map-point / observation packing: reference object graph -> structure-of-arrays.

Reproduces, order and skip rules included, what the reference's optimizer builds before it
iterates (`visual_slam/optimization/torch_optimizer.py:43-95`):

  * gauge: `keyframes[0]` is the fixed keyframe (`:43`, identity test `kf is kf_fixed` `:54`)
  * intrinsics from `keyframes[0]._camera.K` only, through float32 (`:46`, used `:170-171`)
  * per keyframe R_w2c / t_w2c through float32 (`:51-52`); poses are keyed by `keyframe_id` (`:61`)
  * points `mp.position` through float32 (`:64-68`)
  * observations: for point j, for (kf_id, cams) in `mp.get_observations()` insertion order,
    skip kf_id not in the window (`:82-83`), for (cam_id, kp_idx) skip cam_id >= len(kf.keypoints)
    and kp_idx >= len(kf.keypoints[cam_id]) (`:90-93`); uv = `kf.keypoints[cam_id][kp_idx].pt` (`:94`)

The outer loop is over points, so the list is already sorted by point: `pt_offsets` is its CSR.
"""
from __future__ import annotations

import operator
from dataclasses import dataclass
from typing import Any, List

import numpy as np

MIN_OBSERVATIONS = 10   # torch_optimizer.py:99-101


def _load_fastpack():
    """The observation loop as a CPython extension (csrc/fastpack.c), built in-tree by `_build.build_fastpack`
    (`__graft_entry__.build()` does it).  Optional: without it the Python loop below does the same work."""
    import importlib.util
    import os
    try:
        from . import _build
        path = _build.fastpack_path()
        if not os.path.exists(path):
            return None
        spec = importlib.util.spec_from_file_location("_fastpack", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        return None


_fastpack = _load_fastpack()


@dataclass
class PackedWindow:
    intr: np.ndarray        # (4,) f64 (float32-valued)
    R: np.ndarray           # (n_cam,3,3) f64 (float32-valued)
    t: np.ndarray           # (n_cam,3)
    fixed: np.ndarray       # (n_cam,) u8
    X: np.ndarray           # (n_pt,3) f64 (float32-valued)
    obs_cam: np.ndarray     # (n_obs,) i32  index into `keyframes`
    obs_pt: np.ndarray      # (n_obs,) i32
    obs_uv: np.ndarray      # (n_obs,2) f32
    obs_cam_id: np.ndarray  # (n_obs,) i32  the reference's cam_id (sensor index inside the keyframe)
    pt_offsets: np.ndarray  # (n_pt+1,) i32

    @property
    def n_obs(self) -> int:
        return int(self.obs_cam.shape[0])


def pack_window(keyframes: List[Any], map_points: List[Any]) -> PackedWindow:
    kf_fixed = keyframes[0]
    K = np.asarray(kf_fixed._camera.K, dtype=np.float32)
    intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)

    n_cam = len(keyframes)
    R = np.empty((n_cam, 3, 3), dtype=np.float64)
    t = np.empty((n_cam, 3), dtype=np.float64)
    fixed = np.zeros(n_cam, dtype=np.uint8)
    pose_index = {}                       # keyframe_id -> row (a later duplicate id overwrites, as the dict at :61 does)
    for i, kf in enumerate(keyframes):
        R[i] = np.asarray(kf.R_w2c, dtype=np.float32)
        t[i] = np.asarray(kf.t_w2c, dtype=np.float32).reshape(3)
        fixed[i] = 1 if kf is kf_fixed else 0
        pose_index[kf.keyframe_id] = i

    n_pt = len(map_points)
    if n_pt:
        X = np.array([mp.position for mp in map_points], dtype=np.float32).reshape(n_pt, 3).astype(np.float64)
    else:
        X = np.zeros((0, 3), dtype=np.float64)

    kps_of = [keyframes[i].keypoints for i in range(n_cam)]
    packed = None
    if _fastpack is not None and isinstance(map_points, list):
        # the same loop in C; None when a container is not an exact dict / list (then the loop below runs)
        packed = _fastpack.pack_observations(map_points, pose_index, kps_of)
    if packed is not None:
        cam_b, cid_b, uv_b, counts_b = packed
        counts_a = np.frombuffer(counts_b, dtype=np.int64)
        pt_offsets = np.zeros(n_pt + 1, dtype=np.int32)
        if n_pt:
            np.cumsum(counts_a, out=pt_offsets[1:])
        return PackedWindow(
            intr=intr, R=R, t=t, fixed=fixed, X=X,
            obs_cam=np.frombuffer(cam_b, dtype=np.int32).copy(),
            obs_pt=np.repeat(np.arange(n_pt, dtype=np.int32), counts_a) if n_pt else np.zeros(0, dtype=np.int32),
            obs_uv=np.frombuffer(uv_b, dtype=np.float32).reshape(-1, 2).copy(),
            obs_cam_id=np.frombuffer(cid_b, dtype=np.int32).copy(), pt_offsets=pt_offsets)

    # the reference's nested loop (:74-95), kept loop for loop; what can leave the interpreter does: the point
    # index of an observation comes from the per-point counts, the appends are bound once
    obs_cam: List[int] = []
    obs_cid: List[int] = []
    obs_uv: List[Any] = []
    counts: List[int] = []
    add_cam, add_cid, add_uv, add_count = obs_cam.append, obs_cid.append, obs_uv.append, counts.append
    index_of = pose_index.get
    for mp in map_points:
        c0 = len(obs_cam)
        for kf_id, cams in mp.get_observations().items():
            i = index_of(kf_id)
            if i is None:
                continue
            kps = kps_of[i]
            n_sensors = len(kps)
            for cam_id, kp_idx in cams.items():
                if cam_id >= n_sensors:
                    continue
                kp = kps[cam_id]
                if kp_idx >= len(kp):
                    continue
                add_uv(kp[kp_idx].pt)
                add_cam(i)
                add_cid(cam_id)
        add_count(len(obs_cam) - c0)

    pt_offsets = np.zeros(n_pt + 1, dtype=np.int32)
    if n_pt:
        np.cumsum(np.asarray(counts, dtype=np.int64), out=pt_offsets[1:])
    obs_pt = np.repeat(np.arange(n_pt, dtype=np.int32), np.asarray(counts, dtype=np.int64)) if n_pt \
        else np.zeros(0, dtype=np.int32)
    return PackedWindow(
        intr=intr, R=R, t=t, fixed=fixed, X=X,
        obs_cam=np.asarray(obs_cam, dtype=np.int32), obs_pt=obs_pt,
        obs_uv=np.asarray(obs_uv, dtype=np.float32).reshape(-1, 2),
        obs_cam_id=np.asarray(obs_cid, dtype=np.int32), pt_offsets=pt_offsets)


def write_back(keyframes: List[Any], map_points: List[Any], R: np.ndarray, t: np.ndarray,
               X: np.ndarray) -> None:
    """Deliver results the way the reference does (`torch_optimizer.py:188-203`): every keyframe
    (the fixed one included) gets `set_pose_Rt(R 3x3, t 3x1)`, every point `update_position`."""
    for i, kf in enumerate(keyframes):
        kf.set_pose_Rt(np.array(R[i], dtype=np.float64), np.array(t[i], dtype=np.float64).reshape(3, 1))
    for mp, pos in zip(map_points, X):
        mp.update_position(np.asarray(pos))


# ---------------------------------------------------------------------------------------------
# two-view initial refinement: the selection rule of TorchOptimizer.optimize_initial
# ---------------------------------------------------------------------------------------------
@dataclass
class PackedInitial:
    window: PackedWindow     # two cameras (reference fixed, current free), selected points only
    selected: np.ndarray     # indices into the `map_points` list handed in, in selection order
    uv_ref: np.ndarray       # (n_sel, 2) f32
    uv_cur: np.ndarray       # (n_sel, 2) f32


def pack_initial(keyframes: List[Any], map_points: List[Any]) -> PackedInitial:
    """What `optimize_initial` collects before it iterates (`torch_optimizer.py:222-272`):
    `kf_ref, kf_cur = keyframes[:2]`; intrinsics and both poses through float32; a point is used iff
    it has an observation in BOTH keyframes with `cam_id == 0` (`:240-250`) and both keypoint
    indices are in range (`:252-258`); pixel = `kf.keypoints[0][kp_idx].pt`."""
    kf_ref, kf_cur = keyframes[:2]
    K = np.asarray(kf_ref._camera.K, dtype=np.float32)
    intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)
    R = np.stack([np.asarray(kf_ref.R_w2c, dtype=np.float32), np.asarray(kf_cur.R_w2c, dtype=np.float32)]).astype(np.float64)
    t = np.stack([np.asarray(kf_ref.t_w2c, dtype=np.float32).reshape(3),
                  np.asarray(kf_cur.t_w2c, dtype=np.float32).reshape(3)]).astype(np.float64)
    kps_ref, kps_cur = kf_ref.keypoints[0], kf_cur.keypoints[0]
    sel, pos, uvr, uvc = [], [], [], []
    for j, mp in enumerate(map_points):
        obs = mp.get_observations()
        if kf_ref.keyframe_id not in obs or kf_cur.keyframe_id not in obs:
            continue
        o_ref, o_cur = obs[kf_ref.keyframe_id], obs[kf_cur.keyframe_id]
        if 0 not in o_ref or 0 not in o_cur:
            continue
        i_ref, i_cur = o_ref[0], o_cur[0]
        if i_ref >= len(kps_ref) or i_cur >= len(kps_cur):
            continue
        sel.append(j)
        pos.append(mp.position)
        uvr.append(kps_ref[i_ref].pt)
        uvc.append(kps_cur[i_cur].pt)
    n = len(sel)
    X = np.array(pos, dtype=np.float32).reshape(n, 3).astype(np.float64)
    uv_ref = np.asarray(uvr, dtype=np.float32).reshape(n, 2)
    uv_cur = np.asarray(uvc, dtype=np.float32).reshape(n, 2)
    obs_cam = np.tile(np.array([0, 1], dtype=np.int32), n)
    obs_pt = np.repeat(np.arange(n, dtype=np.int32), 2)
    obs_uv = np.stack([uv_ref, uv_cur], axis=1).reshape(2 * n, 2)
    win = PackedWindow(intr=intr, R=R, t=t, fixed=np.array([1, 0], dtype=np.uint8), X=X,
                       obs_cam=obs_cam, obs_pt=obs_pt, obs_uv=obs_uv,
                       obs_cam_id=np.zeros(2 * n, dtype=np.int32),
                       pt_offsets=(2 * np.arange(n + 1)).astype(np.int32))
    return PackedInitial(window=win, selected=np.asarray(sel, dtype=np.int64), uv_ref=uv_ref, uv_cur=uv_cur)


# ---------------------------------------------------------------------------------------------
# fast path: an incremental structure-of-arrays mirror of the observation graph
# ---------------------------------------------------------------------------------------------
_ID_OF = operator.attrgetter("id")


def _stable_order_by(keys: np.ndarray, n_keys: int) -> np.ndarray:
    """argsort(keys, kind="stable") for small non-negative integer keys: a counting sort through scipy's
    COO -> CSR conversion when scipy is there (it keeps the input order inside a row)."""
    try:
        from scipy.sparse import coo_matrix
        n = keys.shape[0]
        csr = coo_matrix((np.arange(1, n + 1, dtype=np.int32), (keys, np.arange(n, dtype=np.int32))),
                         shape=(max(n_keys, 1), max(n, 1))).tocsr()
        if csr.nnz == n:
            return csr.indices.astype(np.int64)
    except Exception:                                   # no scipy: numpy's merge sort
        pass
    return np.argsort(keys, kind="stable")


class ObservationTable:
    """Append-only SoA mirror of `KeyFrame.add_map_point` / `MapPoint.add_observation`
    (`map/keyframe.py:82-86`, `map/map_point.py:37-39`): one row (keyframe_id, point id, cam_id,
    kp_idx, u, v) per recorded observation, in insertion order.

    With it `pack()` builds the same `PackedWindow` as `pack_window` with numpy selections and one
    stable sort instead of a Python loop over every observation (~1-2 us each, which is > 99 % of
    `optimize_local`'s wall time once the solve takes a millisecond).  The table follows the
    reference's first-write-wins rule (`map/observation.py:10-22`): a second (keyframe, cam_id)
    observation of the same point is ignored.  Removals (`KeyFrame.remove_map_point`,
    `map/keyframe.py:93-100`) tombstone the row.
    """

    def __init__(self, capacity: int = 1 << 16):
        self._n = 0
        self._kf = np.empty(capacity, dtype=np.int64)
        self._mp = np.empty(capacity, dtype=np.int64)
        self._cam = np.empty(capacity, dtype=np.int32)
        self._kp = np.empty(capacity, dtype=np.int32)
        self._uv = np.empty((capacity, 2), dtype=np.float32)
        self._alive = np.empty(capacity, dtype=bool)
        self._grp = np.empty(capacity, dtype=np.int64)   # row of the first observation of the same (keyframe, point)
        self._seen = {}          # (kf_id, point key, cam_id) -> row
        self._first = {}         # (kf_id, point key) -> first row (the reference groups a point's observations by keyframe)

    def __len__(self) -> int:
        return self._n

    @staticmethod
    def point_key(mp) -> int:
        return int(getattr(mp, "id", id(mp)))

    def _grow(self):
        for name in ("_kf", "_mp", "_cam", "_kp", "_uv", "_alive", "_grp"):
            a = getattr(self, name)
            b = np.empty((2 * a.shape[0],) + a.shape[1:], dtype=a.dtype)
            b[: self._n] = a[: self._n]
            setattr(self, name, b)

    def add(self, kf, cam_id: int, kp_idx: int, mp) -> bool:
        key = (int(kf.keyframe_id), self.point_key(mp), int(cam_id))
        if key in self._seen:
            return False
        if self._n == self._kf.shape[0]:
            self._grow()
        i = self._n
        kps = kf.keypoints
        in_range = cam_id < len(kps) and kp_idx < len(kps[cam_id])
        self._kf[i], self._mp[i], self._cam[i], self._kp[i] = key[0], key[1], cam_id, kp_idx
        self._uv[i] = kps[cam_id][kp_idx].pt if in_range else (np.nan, np.nan)
        self._alive[i] = in_range        # out-of-range entries are skipped by the reference's packer (:90-93)
        self._grp[i] = self._first.setdefault(key[:2], i)
        self._seen[key] = i
        self._n += 1
        return True

    def remove(self, kf, cam_id: int, mp) -> None:
        row = self._seen.pop((int(kf.keyframe_id), self.point_key(mp), int(cam_id)), None)
        if row is not None:
            self._alive[row] = False

    def install(self, keyframe_cls) -> None:
        """Hook `keyframe_cls.add_map_point` / `remove_map_point` so the table follows the map."""
        table = self
        orig_add = keyframe_cls.add_map_point
        orig_remove = getattr(keyframe_cls, "remove_map_point", None)

        def add_map_point(kf, cam_id, kp_idx, map_point):
            orig_add(kf, cam_id, kp_idx, map_point)
            table.add(kf, cam_id, kp_idx, map_point)

        keyframe_cls.add_map_point = add_map_point
        if orig_remove is not None:
            def remove_map_point(kf, cam_id, kp_idx):
                mp = kf.map_points.get((cam_id, kp_idx)) if hasattr(kf, "map_points") else None
                orig_remove(kf, cam_id, kp_idx)
                if mp is not None:
                    table.remove(kf, cam_id, mp)

            keyframe_cls.remove_map_point = remove_map_point
        self._uninstall = (keyframe_cls, orig_add, orig_remove)

    def uninstall(self) -> None:
        cls, orig_add, orig_remove = self._uninstall
        cls.add_map_point = orig_add
        if orig_remove is not None:
            cls.remove_map_point = orig_remove

    @staticmethod
    def _lookup(window_keys: np.ndarray, row_keys: np.ndarray, last_wins: bool) -> np.ndarray:
        """Window position of every table row's key (int32, -1: not in the window)."""
        m = window_keys.shape[0]
        out = np.full(row_keys.shape[0], -1, dtype=np.int32)
        if m == 0 or row_keys.shape[0] == 0:
            return out
        lo, hi = int(window_keys.min()), int(window_keys.max())
        if lo >= 0 and hi < 8 * (m + row_keys.shape[0]) + (1 << 20):
            lut = np.full(hi + 2, -1, dtype=np.int32)
            pos = np.arange(m, dtype=np.int32)
            if last_wins:
                lut[window_keys] = pos                  # duplicate ids: the later assignment stays
            else:
                lut[window_keys[::-1]] = pos[::-1]      # duplicate ids: the first occurrence stays
            inside = (row_keys >= 0) & (row_keys <= hi)
            return lut[np.where(inside, row_keys, hi + 1)]          # lut[hi + 1] is -1
        order = np.argsort(window_keys, kind="stable")
        skeys = window_keys[order]
        pos = np.searchsorted(skeys, row_keys, side="right" if last_wins else "left")
        if last_wins:
            pos -= 1
        ok = (pos >= 0) & (pos < m)
        ok &= skeys[np.clip(pos, 0, m - 1)] == row_keys
        out[ok] = order[pos[ok]]
        return out

    def pack(self, keyframes: List[Any], map_points: List[Any]) -> PackedWindow:
        """Same result as `pack_window(keyframes, map_points)` for a graph that was built through the
        hooked `add_map_point` (observation order inside a point = insertion order, as
        `mp.get_observations()` yields it)."""
        kf_fixed = keyframes[0]
        K = np.asarray(kf_fixed._camera.K, dtype=np.float32)
        intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)
        n_cam, n_pt = len(keyframes), len(map_points)
        R = np.empty((n_cam, 3, 3), dtype=np.float64)
        t = np.empty((n_cam, 3), dtype=np.float64)
        fixed = np.zeros(n_cam, dtype=np.uint8)
        kf_ids = np.empty(n_cam, dtype=np.int64)
        for i, kf in enumerate(keyframes):
            R[i] = np.asarray(kf.R_w2c, dtype=np.float32)
            t[i] = np.asarray(kf.t_w2c, dtype=np.float32).reshape(3)
            fixed[i] = 1 if kf is kf_fixed else 0
            kf_ids[i] = kf.keyframe_id
        if n_pt:
            try:                                        # positions are (3,) arrays: one concatenation, one rounding
                X = np.concatenate([mp.position for mp in map_points]).astype(np.float32).astype(np.float64).reshape(n_pt, 3)
            except ValueError:                          # a position that is not a flat array of three
                X = np.array([mp.position for mp in map_points], dtype=np.float32).reshape(n_pt, 3).astype(np.float64)
        else:
            X = np.zeros((0, 3), dtype=np.float64)
        try:
            mp_keys = np.fromiter(map(_ID_OF, map_points), dtype=np.int64, count=n_pt)
        except AttributeError:                          # points without an `id`: the object identity is the key
            mp_keys = np.fromiter((self.point_key(mp) for mp in map_points), dtype=np.int64, count=n_pt)

        n = self._n
        alive = self._alive[:n]
        # key -> window row through a lookup table when the ids are small integers (the reference's global
        # counters), through a sorted search otherwise (object identities)
        cam_of = self._lookup(kf_ids, self._kf[:n], last_wins=True)      # a later duplicate id wins, as the dict at :61 does
        pt_of = self._lookup(mp_keys, self._mp[:n], last_wins=False)
        rows = np.flatnonzero(alive & (cam_of >= 0) & (pt_of >= 0))
        cam_row, pt_row = cam_of[rows], pt_of[rows]
        # group by point; inside a point by keyframe in first-insertion order, then by insertion (= the nested
        # dict iteration of the reference's packing loop, :80-95).  Rows are in insertion order already, so
        # when no (keyframe, point) pair was observed twice a stable counting sort by point is all it takes.
        if np.array_equal(self._grp[rows], rows):
            o = _stable_order_by(pt_row, n_pt)
        else:
            o = np.lexsort((rows, self._grp[rows], pt_row))
        rows, cam_row, pt_row = rows[o], cam_row[o], pt_row[o]
        pt_offsets = np.zeros(n_pt + 1, dtype=np.int32)
        np.cumsum(np.bincount(pt_row, minlength=n_pt), out=pt_offsets[1:])
        return PackedWindow(intr=intr, R=R, t=t, fixed=fixed, X=X,
                            obs_cam=cam_row.astype(np.int32), obs_pt=pt_row.astype(np.int32),
                            obs_uv=np.ascontiguousarray(self._uv[rows]),
                            obs_cam_id=np.ascontiguousarray(self._cam[rows]), pt_offsets=pt_offsets)
