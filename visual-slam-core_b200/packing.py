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
import threading
from dataclasses import dataclass
from typing import Any, List, Optional

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
    write_back_points(map_points, X)


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


def gather_points(map_points: List[Any]):
    """Positions (n,3) float64 and integer keys (`mp.id`, else the object identity) of a list of map points:
    one C loop when the extension is built, a comprehension otherwise."""
    n_pt = len(map_points)
    if n_pt == 0:
        return np.zeros((0, 3), dtype=np.float64), np.zeros(0, dtype=np.int64)
    if _fastpack is not None and hasattr(_fastpack, "gather_points") and isinstance(map_points, list):
        out = _fastpack.gather_points(map_points)
        if out is not None:
            return (np.frombuffer(out[0], dtype=np.float64).reshape(n_pt, 3),
                    np.frombuffer(out[1], dtype=np.int64))
    try:
        X = np.concatenate([mp.position for mp in map_points]).astype(np.float64).reshape(n_pt, 3)
    except ValueError:                                  # a position that is not a flat array of three
        X = np.array([mp.position for mp in map_points], dtype=np.float64).reshape(n_pt, 3)
    try:
        keys = np.fromiter(map(_ID_OF, map_points), dtype=np.int64, count=n_pt)
    except AttributeError:                              # points without an `id`: the object identity is the key
        keys = np.fromiter((ObservationTable.point_key(mp) for mp in map_points), dtype=np.int64, count=n_pt)
    return X, keys


def _update_position_template(self, new_position):
    with self._lock:
        self.position = np.asarray(new_position, dtype=float)


_PLAIN_UPDATE = {}


def _plain_update_position(cls) -> bool:
    """True when `cls.update_position` is, instruction for instruction, the reference's
    `with self._lock: self.position = np.asarray(new_position, dtype=float)` (`map/map_point.py:61-63`): the
    write-back may then store the float64 row as `mp.position` directly (what `np.asarray` would return; a single
    attribute store, atomic under the GIL, which is all the lock around it protects).  Any other implementation --
    a subclass that records the change, another dtype -- keeps the per-point method call."""
    hit = _PLAIN_UPDATE.get(cls)
    if hit is None:
        fn = getattr(cls, "update_position", None)
        code, ref = getattr(fn, "__code__", None), _update_position_template.__code__
        hit = bool(code is not None and code.co_code == ref.co_code and code.co_names == ref.co_names
                   and code.co_consts == ref.co_consts and code.co_varnames == ref.co_varnames
                   and getattr(getattr(fn, "__globals__", {}).get("np"), "__name__", "") == "numpy"
                   and not hasattr(cls, "__slots__") and "position" not in vars(cls))
        _PLAIN_UPDATE[cls] = hit
    return hit


def write_back_points(map_points: List[Any], X: np.ndarray) -> None:
    """`mp.update_position(X[j])` for every point (`torch_optimizer.py:201-203`), the loop in C when it can be."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    if _fastpack is not None and hasattr(_fastpack, "scatter_points") and isinstance(map_points, list) and map_points:
        cls = type(map_points[0])
        try:
            if _fastpack.scatter_points(map_points, X, cls if _plain_update_position(cls) else None):
                return
        except TypeError:                               # an older build of the extension: two arguments
            if _fastpack.scatter_points(map_points, X):
                return
    for mp, pos in zip(map_points, X):
        mp.update_position(pos)


class ObservationTable:
    """Append-only SoA mirror of the map's observation dictionaries: one row (point key, keyframe_id, cam_id,
    kp_idx) per `MapPoint.add_observation` (`map/map_point.py:37-39`, reached from `KeyFrame.add_map_point`,
    `map/keyframe.py:82-86`), in insertion order; `remove_observation` / `clear_observations` tombstone rows.
    Pixels are not stored per row: each keyframe's keypoints are mirrored ONCE as a float32 (N_kp, 2) array per
    sensor (they do not change after the keyframe is created) and gathered at pack time.

    `pack()` then builds the same `PackedWindow` as `pack_window` -- array for array, order and skip rules of
    `torch_optimizer.py:74-95` included -- with numpy selections and one stable counting sort instead of a Python
    loop over every observation (~1-2 us each, > 90 % of `optimize_local`'s wall time once the solve takes a
    millisecond).

    A point is mirrored either from its first observation on (hook) or, if it existed before the hooks were
    installed, by one walk over its own dictionary the first time a window contains it (`ingest`).  Anything
    the table cannot represent (a point handed in twice, a container it does not know) makes `pack()` return
    None and the caller falls back to `pack_window`.  All entry points take the table's lock: the map is mutated
    from the tracking / local-mapping threads while `pack()` runs in the local-BA thread.
    """

    def __init__(self, capacity: int = 1 << 16):
        self._lock = threading.RLock()
        self._n = 0
        self._kf = np.empty(capacity, dtype=np.int64)
        self._mp = np.empty(capacity, dtype=np.int64)
        self._cam = np.empty(capacity, dtype=np.int32)
        self._kp = np.empty(capacity, dtype=np.int64)
        self._alive = np.empty(capacity, dtype=bool)
        self._grp = np.empty(capacity, dtype=np.int64)   # row of the first live observation of the same (point, keyframe)
        self._pairs = {}         # (point key, kf_id) -> [first row, {cam_id: row}]   live observations only
        self._known = set()      # point keys whose dictionaries are mirrored in full
        self._kp_cache = {}      # kf_id -> (tuple of sensor lengths, [float32 (N,2) per sensor])
        self._hooks = []
        self.stats = {"ingested_points": 0, "packs": 0, "fallbacks": 0}

    def __len__(self) -> int:
        return self._n

    @staticmethod
    def point_key(mp) -> int:
        return int(getattr(mp, "id", id(mp)))

    def _grow(self):
        for name in ("_kf", "_mp", "_cam", "_kp", "_alive", "_grp"):
            a = getattr(self, name)
            b = np.empty((2 * a.shape[0],) + a.shape[1:], dtype=a.dtype)
            b[: self._n] = a[: self._n]
            setattr(self, name, b)

    # ----------------------------------------------------------------------------- rows
    def _add_row(self, pt: int, kf_id: int, cam_id: int, kp_idx: int) -> bool:
        pair = self._pairs.get((pt, kf_id))
        if pair is not None and cam_id in pair[1]:
            return False                                  # first write wins (map/observation.py:10-22)
        if self._n == self._kf.shape[0]:
            self._grow()
        i = self._n
        self._kf[i], self._mp[i], self._cam[i], self._kp[i] = kf_id, pt, cam_id, kp_idx
        self._alive[i] = True
        if pair is None:
            self._pairs[(pt, kf_id)] = [i, {cam_id: i}]
            self._grp[i] = i
        else:
            pair[1][cam_id] = i
            self._grp[i] = pair[0]
        self._n += 1
        return True

    def _remove_rows(self, pt: int, kf_id: int, cam_id) -> None:
        pair = self._pairs.get((pt, kf_id))
        if pair is None:
            return
        if cam_id is None:                                # the whole keyframe entry goes (map/observation.py:36-38)
            for row in pair[1].values():
                self._alive[row] = False
            del self._pairs[(pt, kf_id)]
            return
        row = pair[1].pop(cam_id, None)
        if row is not None:
            self._alive[row] = False
            if not pair[1]:                               # a later re-add starts a new dictionary entry, at the end
                del self._pairs[(pt, kf_id)]

    def add(self, kf, cam_id: int, kp_idx: int, mp) -> bool:
        """Record one observation by hand (what the hooks do); marks the point as mirrored."""
        with self._lock:
            pt = self.point_key(mp)
            self._known.add(pt)
            return self._add_row(pt, int(kf.keyframe_id), int(cam_id), int(kp_idx))

    def remove(self, kf, cam_id, mp) -> None:
        with self._lock:
            self._remove_rows(self.point_key(mp), int(kf.keyframe_id), None if cam_id is None else int(cam_id))

    def ingest(self, mp) -> None:
        """Mirror a point that existed before the hooks: one walk over its own dictionary, in its order."""
        with self._lock:
            pt = self.point_key(mp)
            if pt in self._known:
                return
            for kf_id, cams in mp.get_observations().items():
                for cam_id, kp_idx in cams.items():
                    self._add_row(pt, int(kf_id), int(cam_id), int(kp_idx))
            self._known.add(pt)
            self.stats["ingested_points"] += 1

    def _forget(self, pt: int) -> None:
        n = self._n
        self._alive[:n][self._mp[:n] == pt] = False
        for key in [k for k in self._pairs if k[0] == pt]:
            del self._pairs[key]

    # ----------------------------------------------------------------------------- hooks
    def _on_add(self, mp, kf_id, cam_id, kp_idx, n_before: int) -> None:
        with self._lock:
            pt = self.point_key(mp)
            if pt not in self._known:
                if n_before != 0:
                    return                                # older point, not mirrored yet: `ingest` will take all of it
                self._known.add(pt)
            self._add_row(pt, int(kf_id), int(cam_id), int(kp_idx))

    def install(self, cls) -> None:
        """Hook a MapPoint class (`add_observation` / `remove_observation` / `clear_observations`) so the table
        follows the map.  A KeyFrame class may be given instead (older call sites): its `add_map_point` /
        `remove_map_point` are hooked, which covers a map that is only ever changed through them."""
        table = self
        if hasattr(cls, "add_observation"):
            orig_add = cls.add_observation
            orig_remove = getattr(cls, "remove_observation", None)
            orig_clear = getattr(cls, "clear_observations", None)

            def add_observation(mp, kf_idx, cam_id, kp_idx):
                n_before = table._n_observations(mp)
                ok = orig_add(mp, kf_idx, cam_id, kp_idx)
                if ok or ok is None:
                    table._on_add(mp, kf_idx, cam_id, kp_idx, n_before)
                return ok

            cls.add_observation = add_observation
            restore = [("add_observation", orig_add)]
            if orig_remove is not None:
                def remove_observation(mp, kf_id, cam_id=None):
                    out = orig_remove(mp, kf_id, cam_id)
                    with table._lock:
                        table._remove_rows(table.point_key(mp), int(kf_id), None if cam_id is None else int(cam_id))
                    return out

                cls.remove_observation = remove_observation
                restore.append(("remove_observation", orig_remove))
            if orig_clear is not None:
                def clear_observations(mp):
                    out = orig_clear(mp)
                    with table._lock:
                        table._forget(table.point_key(mp))
                    return out

                cls.clear_observations = clear_observations
                restore.append(("clear_observations", orig_clear))
            self._hooks.append((cls, restore))
            return
        orig_add = cls.add_map_point
        orig_remove = getattr(cls, "remove_map_point", None)

        def add_map_point(kf, cam_id, kp_idx, map_point):
            n_before = table._n_observations(map_point)
            orig_add(kf, cam_id, kp_idx, map_point)
            table._on_add(map_point, kf.keyframe_id, cam_id, kp_idx, n_before)

        cls.add_map_point = add_map_point
        restore = [("add_map_point", orig_add)]
        if orig_remove is not None:
            def remove_map_point(kf, cam_id, kp_idx):
                mp = kf.map_points.get((cam_id, kp_idx)) if hasattr(kf, "map_points") else None
                orig_remove(kf, cam_id, kp_idx)
                if mp is not None:
                    table.remove(kf, cam_id, mp)

            cls.remove_map_point = remove_map_point
            restore.append(("remove_map_point", orig_remove))
        self._hooks.append((cls, restore))

    @staticmethod
    def _n_observations(mp) -> int:
        try:
            f = getattr(mp, "num_observations", None)
            if f is not None:
                return int(f())
            return len(mp.get_observations())
        except Exception:
            return 1                                      # unknown: treat as an older point, `ingest` decides

    def installed_on(self, cls) -> bool:
        return any(c is cls for c, _ in self._hooks)

    def uninstall(self) -> None:
        for cls, restore in reversed(self._hooks):
            for name, orig in restore:
                setattr(cls, name, orig)
        self._hooks = []

    # ----------------------------------------------------------------------------- pack
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

    def _keypoint_arrays(self, kf):
        """float32 (N,2) pixel array per sensor of a keyframe, built once (keypoints do not change after the
        keyframe is created; the cache is keyed by id and checked against the sensor lengths)."""
        kps = kf.keypoints
        lens = tuple(len(s) for s in kps)
        hit = self._kp_cache.get(kf.keyframe_id)
        if hit is not None and hit[0] == lens:
            return hit[1]
        arrs = []
        for s in kps:
            if len(s) == 0:
                arrs.append(np.zeros((0, 2), dtype=np.float32))
                continue
            a = None
            try:
                import cv2
                if isinstance(s[0], cv2.KeyPoint):
                    a = cv2.KeyPoint_convert(list(s)).astype(np.float32).reshape(-1, 2)
            except Exception:
                a = None
            if a is None:
                a = np.array([kp.pt for kp in s], dtype=np.float32).reshape(-1, 2)
            arrs.append(a)
        self._kp_cache[kf.keyframe_id] = (lens, arrs)
        if len(self._kp_cache) > 4096:                   # bounded: drop the oldest entries
            for k in list(self._kp_cache)[:1024]:
                del self._kp_cache[k]
        return arrs

    def pack(self, keyframes: List[Any], map_points: List[Any]) -> Optional[PackedWindow]:
        """Same result as `pack_window(keyframes, map_points)`, or None when the window has something the table
        does not represent (a point listed twice) -- the caller then walks the graph."""
        with self._lock:
            self.stats["packs"] += 1
            kf_fixed = keyframes[0]
            K = np.asarray(kf_fixed._camera.K, dtype=np.float32)
            intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)
            n_cam, n_pt = len(keyframes), len(map_points)
            R = np.empty((n_cam, 3, 3), dtype=np.float64)
            t = np.empty((n_cam, 3), dtype=np.float64)
            fixed = np.zeros(n_cam, dtype=np.uint8)
            kf_ids = np.empty(n_cam, dtype=np.int64)
            uv_parts, n_sens = [], np.zeros(n_cam, dtype=np.int64)
            max_sens = max((len(kf.keypoints) for kf in keyframes), default=0)
            base = np.zeros((n_cam, max(max_sens, 1)), dtype=np.int64)
            count = np.zeros((n_cam, max(max_sens, 1)), dtype=np.int64)
            at = 0
            for i, kf in enumerate(keyframes):
                R[i] = np.asarray(kf.R_w2c, dtype=np.float32)
                t[i] = np.asarray(kf.t_w2c, dtype=np.float32).reshape(3)
                fixed[i] = 1 if kf is kf_fixed else 0
                kf_ids[i] = kf.keyframe_id
                arrs = self._keypoint_arrays(kf)
                n_sens[i] = len(arrs)
                for c, a in enumerate(arrs):
                    base[i, c], count[i, c] = at, a.shape[0]
                    uv_parts.append(a)
                    at += a.shape[0]
            Xd, mp_keys = gather_points(map_points)
            X = Xd.astype(np.float32).astype(np.float64)
            if n_pt:
                known = self._known
                key_list = mp_keys.tolist()
                if not known.issuperset(key_list):
                    for j, k in enumerate(key_list):
                        if k not in known:
                            self.ingest(map_points[j])
            uv_all = np.concatenate(uv_parts) if uv_parts else np.zeros((0, 2), dtype=np.float32)
            n = self._n
            if _fastpack is not None and hasattr(_fastpack, "table_select"):
                # selection, skip rules and the counting sort by point in one C pass over the table's rows
                sel = _fastpack.table_select(self._kf, self._mp, self._cam, self._kp, self._alive.view(np.uint8), self._grp, n,
                                             kf_ids, np.ascontiguousarray(mp_keys), n_sens, base, count, base.shape[1],
                                             np.ascontiguousarray(uv_all))
                if sel is not None:
                    pt_offsets = np.frombuffer(sel[3], dtype=np.int32)
                    return PackedWindow(intr=intr, R=R, t=t, fixed=fixed, X=X,
                                        obs_cam=np.frombuffer(sel[0], dtype=np.int32),
                                        obs_pt=np.repeat(np.arange(n_pt, dtype=np.int32), np.diff(pt_offsets)),
                                        obs_uv=np.frombuffer(sel[2], dtype=np.float32).reshape(-1, 2),
                                        obs_cam_id=np.frombuffer(sel[1], dtype=np.int32), pt_offsets=pt_offsets)
            if n_pt and np.unique(mp_keys).shape[0] != n_pt:        # a point handed in twice: its rows would be needed twice
                self.stats["fallbacks"] += 1
                return None

            n = self._n
            # key -> window row through a lookup table when the ids are small integers (the reference's global
            # counters), through a sorted search otherwise (object identities)
            cam_of = self._lookup(kf_ids, self._kf[:n], last_wins=True)      # a later duplicate id wins, as the dict at :61 does
            pt_of = self._lookup(mp_keys, self._mp[:n], last_wins=False)
            rows = np.flatnonzero(self._alive[:n] & (cam_of >= 0) & (pt_of >= 0))
            cam_row, pt_row = cam_of[rows], pt_of[rows]
            cid, kpi = self._cam[rows], self._kp[rows]
            # the packer's skip rules (:90-93): sensor index and keypoint index must be in range
            ns = n_sens[cam_row]
            ok = (cid < ns) & (cid >= -ns)
            cc = np.where(ok, np.where(cid < 0, cid + ns, cid), 0)     # Python's negative indexing, as kps[cam_id] would
            ok &= (kpi < count[cam_row, cc]) & (kpi >= -count[cam_row, cc])
            if not ok.all():
                rows, cam_row, pt_row, cid, kpi, cc = rows[ok], cam_row[ok], pt_row[ok], cid[ok], kpi[ok], cc[ok]
            kpi = np.where(kpi < 0, kpi + count[cam_row, cc], kpi)           # Python's negative indexing, as kps[idx] would
            # group by point; inside a point by keyframe in first-insertion order, then by insertion (= the nested
            # dict iteration of the reference's packing loop, :80-95).  Rows are in insertion order already, so
            # when no (keyframe, point) pair was observed twice a stable counting sort by point is all it takes.
            if np.array_equal(self._grp[rows], rows):
                o = _stable_order_by(pt_row, n_pt)
            else:
                o = np.lexsort((rows, self._grp[rows], pt_row))
            cam_row, pt_row, cid = cam_row[o], pt_row[o], cid[o]
            src = (base[cam_row, cc[o]] + kpi[o])
            pt_offsets = np.zeros(n_pt + 1, dtype=np.int32)
            if n_pt:
                np.cumsum(np.bincount(pt_row, minlength=n_pt), out=pt_offsets[1:])
            return PackedWindow(intr=intr, R=R, t=t, fixed=fixed, X=X,
                                obs_cam=cam_row.astype(np.int32), obs_pt=pt_row.astype(np.int32),
                                obs_uv=np.ascontiguousarray(uv_all[src]) if src.size else np.zeros((0, 2), dtype=np.float32),
                                obs_cam_id=np.ascontiguousarray(cid.astype(np.int32)), pt_offsets=pt_offsets)


_SHARED_TABLE: Optional[ObservationTable] = None
_SHARED_LOCK = threading.Lock()


def shared_table() -> ObservationTable:
    """The process-wide table the plugin uses by default: one set of hooks per map-point class, however many
    optimizer instances there are."""
    global _SHARED_TABLE
    with _SHARED_LOCK:
        if _SHARED_TABLE is None:
            _SHARED_TABLE = ObservationTable()
        return _SHARED_TABLE
