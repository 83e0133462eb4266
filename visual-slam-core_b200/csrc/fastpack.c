/* The observation loop of the window packer in C (CPython API): same iteration order, same skip rules as
 * `packing.pack_window`, i.e. as the reference's packing loop
 * (visual_slam/optimization/torch_optimizer.py:74-95):
 *
 *   for mp in map_points:
 *       for kf_id, cams in mp.get_observations().items():      # insertion order
 *           i = pose_index.get(kf_id);  skip if the keyframe is not in the window        (:82-83)
 *           for cam_id, kp_idx in cams.items():
 *               skip if cam_id >= len(kf.keypoints) or kp_idx >= len(kf.keypoints[cam_id])  (:90-93)
 *               uv = kf.keypoints[cam_id][kp_idx].pt                                     (:94)
 *
 * The interpreter spends ~1.5 us per observation on this loop; the solve it feeds takes a millisecond.
 * Only exact dict containers and exact list / tuple keypoint sequences take this path; anything else makes the
 * function return None and the caller runs the Python loop.  Host-side helper: no CUDA here, and nothing of the product depends on it being
 * built (packing.py falls back to the Python loop).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t* cam;
  int32_t* cid;
  float* uv;
  size_t n, cap;
} Rows;

static int rows_push(Rows* r, int32_t cam, int32_t cid, float u, float v) {
  if (r->n == r->cap) {
    size_t cap = r->cap ? 2 * r->cap : 4096;
    int32_t* a = (int32_t*)realloc(r->cam, cap * sizeof(int32_t));
    if (!a) return -1;
    r->cam = a;
    int32_t* b = (int32_t*)realloc(r->cid, cap * sizeof(int32_t));
    if (!b) return -1;
    r->cid = b;
    float* c = (float*)realloc(r->uv, 2 * cap * sizeof(float));
    if (!c) return -1;
    r->uv = c;
    r->cap = cap;
  }
  r->cam[r->n] = cam;
  r->cid[r->n] = cid;
  r->uv[2 * r->n] = u;
  r->uv[2 * r->n + 1] = v;
  r->n++;
  return 0;
}

static void rows_free(Rows* r) {
  free(r->cam);
  free(r->cid);
  free(r->uv);
}

/* exact list (1) or exact tuple (2) -- cv2's detectors hand out tuples of KeyPoint -- anything else 0 */
static int seq_kind(PyObject* o) { return PyList_CheckExact(o) ? 1 : (PyTuple_CheckExact(o) ? 2 : 0); }
static Py_ssize_t seq_len(PyObject* o, int kind) { return kind == 1 ? PyList_GET_SIZE(o) : PyTuple_GET_SIZE(o); }

/* seq[index] with Python's rules for negative indices; NULL + IndexError when out of range (borrowed ref) */
static PyObject* seq_item(PyObject* o, int kind, Py_ssize_t idx) {
  const Py_ssize_t n = seq_len(o, kind);
  if (idx < 0) idx += n;
  if (idx < 0 || idx >= n) {
    PyErr_SetString(PyExc_IndexError, kind == 1 ? "list index out of range" : "tuple index out of range");
    return NULL;
  }
  return kind == 1 ? PyList_GET_ITEM(o, idx) : PyTuple_GET_ITEM(o, idx);
}

/* pack_observations(map_points, pose_index, keypoints_of) -> (cam, cam_id, uv, counts) as bytes, or None */
static PyObject* pack_observations(PyObject* self, PyObject* args) {
  PyObject *map_points, *pose_index, *kps_of;
  if (!PyArg_ParseTuple(args, "OOO", &map_points, &pose_index, &kps_of)) return NULL;
  if (!PyList_CheckExact(map_points) || !PyDict_CheckExact(pose_index) || !PyList_CheckExact(kps_of)) Py_RETURN_NONE;

  static PyObject *s_get = NULL, *s_pt = NULL;
  if (!s_get) s_get = PyUnicode_InternFromString("get_observations");
  if (!s_pt) s_pt = PyUnicode_InternFromString("pt");
  if (!s_get || !s_pt) return NULL;

  const Py_ssize_t n_pt = PyList_GET_SIZE(map_points);
  Rows rows = {0};
  int64_t* counts = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_pt > 0 ? n_pt : 1));
  if (!counts) return PyErr_NoMemory();
  int unsupported = 0;
  PyObject* obs = NULL;

  for (Py_ssize_t j = 0; j < n_pt && !unsupported; ++j) {
    PyObject* mp = PyList_GET_ITEM(map_points, j);
    const size_t c0 = rows.n;
    obs = PyObject_CallMethodNoArgs(mp, s_get);
    if (!obs) goto fail;
    if (!PyDict_CheckExact(obs)) { unsupported = 1; break; }
    Py_ssize_t p1 = 0;
    PyObject *kf_id, *cams;
    while (PyDict_Next(obs, &p1, &kf_id, &cams)) {
      PyObject* row = PyDict_GetItemWithError(pose_index, kf_id);      /* borrowed */
      if (!row) {
        if (PyErr_Occurred()) goto fail;
        continue;                                                      /* keyframe not in the window */
      }
      const long i = PyLong_AsLong(row);
      if (i == -1 && PyErr_Occurred()) goto fail;
      if (i < 0 || i >= PyList_GET_SIZE(kps_of)) { PyErr_SetString(PyExc_IndexError, "pose index out of range"); goto fail; }
      PyObject* kps = PyList_GET_ITEM(kps_of, i);
      const int kps_kind = seq_kind(kps);
      if (!PyDict_CheckExact(cams) || !kps_kind) { unsupported = 1; break; }
      const Py_ssize_t n_sensors = seq_len(kps, kps_kind);
      Py_ssize_t p2 = 0;
      PyObject *cam_id_o, *kp_idx_o;
      while (PyDict_Next(cams, &p2, &cam_id_o, &kp_idx_o)) {
        const Py_ssize_t cam_id = PyLong_AsSsize_t(cam_id_o);
        if (cam_id == -1 && PyErr_Occurred()) { PyErr_Clear(); unsupported = 1; break; }
        if (cam_id >= n_sensors) continue;
        PyObject* kp_list = seq_item(kps, kps_kind, cam_id);
        if (!kp_list) goto fail;
        const int kp_kind = seq_kind(kp_list);
        if (!kp_kind) { unsupported = 1; break; }
        const Py_ssize_t kp_idx = PyLong_AsSsize_t(kp_idx_o);
        if (kp_idx == -1 && PyErr_Occurred()) { PyErr_Clear(); unsupported = 1; break; }
        if (kp_idx >= seq_len(kp_list, kp_kind)) continue;
        PyObject* kp = seq_item(kp_list, kp_kind, kp_idx);
        if (!kp) goto fail;
        PyObject* pt = PyObject_GetAttr(kp, s_pt);
        if (!pt) goto fail;
        double u, v;
        if (PyTuple_CheckExact(pt) && PyTuple_GET_SIZE(pt) == 2) {
          u = PyFloat_AsDouble(PyTuple_GET_ITEM(pt, 0));
          v = PyFloat_AsDouble(PyTuple_GET_ITEM(pt, 1));
        } else {
          PyObject* a = PySequence_GetItem(pt, 0);
          PyObject* b = a ? PySequence_GetItem(pt, 1) : NULL;
          u = a ? PyFloat_AsDouble(a) : -1.0;
          v = b ? PyFloat_AsDouble(b) : -1.0;
          Py_XDECREF(a);
          Py_XDECREF(b);
          if (!a || !b) { Py_DECREF(pt); goto fail; }
        }
        Py_DECREF(pt);
        if ((u == -1.0 || v == -1.0) && PyErr_Occurred()) goto fail;
        if (cam_id > INT32_MAX || cam_id < INT32_MIN) { PyErr_SetString(PyExc_OverflowError, "cam_id"); goto fail; }
        if (rows_push(&rows, (int32_t)i, (int32_t)cam_id, (float)u, (float)v) != 0) { PyErr_NoMemory(); goto fail; }
      }
      if (unsupported) break;
    }
    Py_CLEAR(obs);
    counts[j] = (int64_t)(rows.n - c0);
  }
  Py_CLEAR(obs);
  if (unsupported) {
    rows_free(&rows);
    free(counts);
    Py_RETURN_NONE;
  }
  {
    /* "y#" turns a NULL pointer into None: an empty result still has to be (empty) bytes */
    const char* none = "";
    PyObject* out = Py_BuildValue(
        "(y#y#y#y#)", rows.n ? (const char*)rows.cam : none, (Py_ssize_t)(rows.n * sizeof(int32_t)),
        rows.n ? (const char*)rows.cid : none, (Py_ssize_t)(rows.n * sizeof(int32_t)),
        rows.n ? (const char*)rows.uv : none, (Py_ssize_t)(2 * rows.n * sizeof(float)),
        (const char*)counts, (Py_ssize_t)((size_t)n_pt * sizeof(int64_t)));
    rows_free(&rows);
    free(counts);
    return out;
  }
fail:
  Py_XDECREF(obs);
  rows_free(&rows);
  free(counts);
  return NULL;
}


/* gather_points(map_points) -> (positions f64[n*3], keys i64[n]) as bytes, or None when a position is not three numbers.
 * key = mp.id when the point has one (the reference's global counter, map/map_point.py:22-24), else the object identity. */
static PyObject* gather_points(PyObject* self, PyObject* args) {
  PyObject* map_points;
  if (!PyArg_ParseTuple(args, "O", &map_points)) return NULL;
  if (!PyList_CheckExact(map_points)) Py_RETURN_NONE;
  static PyObject *s_pos = NULL, *s_id = NULL;
  if (!s_pos) s_pos = PyUnicode_InternFromString("position");
  if (!s_id) s_id = PyUnicode_InternFromString("id");
  if (!s_pos || !s_id) return NULL;
  const Py_ssize_t n = PyList_GET_SIZE(map_points);
  PyObject* pos_b = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)(n * 3 * sizeof(double)));
  PyObject* key_b = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)(n * sizeof(int64_t)));
  if (!pos_b || !key_b) { Py_XDECREF(pos_b); Py_XDECREF(key_b); return NULL; }
  double* X = (double*)PyBytes_AS_STRING(pos_b);
  int64_t* keys = (int64_t*)PyBytes_AS_STRING(key_b);
  int unsupported = 0;
  for (Py_ssize_t j = 0; j < n; ++j) {
    PyObject* mp = PyList_GET_ITEM(map_points, j);
    PyObject* pos = PyObject_GetAttr(mp, s_pos);
    if (!pos) goto fail;
    int done = 0;
    Py_buffer view;
    if (PyObject_CheckBuffer(pos) && PyObject_GetBuffer(pos, &view, PyBUF_FORMAT | PyBUF_C_CONTIGUOUS) == 0) {
      if (view.format && view.format[0] == 'd' && view.format[1] == 0 && view.len == 3 * (Py_ssize_t)sizeof(double)) {
        memcpy(X + 3 * j, view.buf, 3 * sizeof(double));
        done = 1;
      } else if (view.format && view.format[0] == 'f' && view.format[1] == 0 && view.len == 3 * (Py_ssize_t)sizeof(float)) {
        const float* f = (const float*)view.buf;
        X[3 * j] = f[0]; X[3 * j + 1] = f[1]; X[3 * j + 2] = f[2];
        done = 1;
      }
      PyBuffer_Release(&view);
    } else {
      PyErr_Clear();
    }
    if (!done) {
      /* any flat sequence of three numbers */
      if (PySequence_Check(pos) && PySequence_Size(pos) == 3) {
        done = 1;
        for (int c = 0; c < 3; ++c) {
          PyObject* it = PySequence_GetItem(pos, c);
          if (!it) { Py_DECREF(pos); goto fail; }
          const double v = PyFloat_AsDouble(it);
          Py_DECREF(it);
          if (v == -1.0 && PyErr_Occurred()) { PyErr_Clear(); done = 0; break; }
          X[3 * j + c] = v;
        }
      } else {
        PyErr_Clear();
      }
    }
    Py_DECREF(pos);
    if (!done) { unsupported = 1; break; }
    PyObject* ido = PyObject_GetAttr(mp, s_id);
    if (ido) {
      const long long k = PyLong_AsLongLong(ido);
      Py_DECREF(ido);
      if (k == -1 && PyErr_Occurred()) { PyErr_Clear(); keys[j] = (int64_t)(intptr_t)mp; }
      else keys[j] = (int64_t)k;
    } else {
      PyErr_Clear();
      keys[j] = (int64_t)(intptr_t)mp;                 /* id(mp) */
    }
  }
  if (unsupported) { Py_DECREF(pos_b); Py_DECREF(key_b); Py_RETURN_NONE; }
  {
    PyObject* out = PyTuple_Pack(2, pos_b, key_b);
    Py_DECREF(pos_b);
    Py_DECREF(key_b);
    return out;
  }
fail:
  Py_DECREF(pos_b);
  Py_DECREF(key_b);
  return NULL;
}

/* scatter_points(map_points, X[, plain_type]) -> True: mp.update_position(X[j]) for every point, X any sequence of rows
 * (a C-contiguous (n,3) float64 numpy array hands out row views).  The write-back of torch_optimizer.py:201-203.
 * `plain_type`: a class whose update_position is known (packing._plain_update_position) to be exactly
 * `with self._lock: self.position = np.asarray(new_position, dtype=float)` (map/map_point.py:61-63); for points of
 * exactly that type the float64 row is stored as `mp.position` directly -- the same object np.asarray would
 * return, one attribute store under the GIL -- without the per-point method call. */
static PyObject* scatter_points(PyObject* self, PyObject* args) {
  PyObject *map_points, *X, *plain = NULL;
  if (!PyArg_ParseTuple(args, "OO|O", &map_points, &X, &plain)) return NULL;
  if (!PyList_CheckExact(map_points) || !PySequence_Check(X)) Py_RETURN_FALSE;
  if (plain == Py_None) plain = NULL;
  static PyObject *s_upd = NULL, *s_pos = NULL;
  if (!s_upd) s_upd = PyUnicode_InternFromString("update_position");
  if (!s_pos) s_pos = PyUnicode_InternFromString("position");
  if (!s_upd || !s_pos) return NULL;
  const Py_ssize_t n = PyList_GET_SIZE(map_points);
  if (PySequence_Size(X) < n) { PyErr_SetString(PyExc_ValueError, "fewer rows than points"); return NULL; }
  for (Py_ssize_t j = 0; j < n; ++j) {
    PyObject* mp = PyList_GET_ITEM(map_points, j);
    PyObject* row = PySequence_GetItem(X, j);
    if (!row) return NULL;
    if (plain && (PyObject*)Py_TYPE(mp) == plain) {
      const int rc = PyObject_SetAttr(mp, s_pos, row);
      Py_DECREF(row);
      if (rc != 0) return NULL;
      continue;
    }
    PyObject* r = PyObject_CallMethodOneArg(mp, s_upd, row);
    Py_DECREF(row);
    if (!r) return NULL;
    Py_DECREF(r);
  }
  Py_RETURN_TRUE;
}

/* ------------------------------------------------------------------------------------------------------------
 * open-addressing set / map of 64-bit keys (window collection and the observation table's selection)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t* key;
  int32_t* val;
  uint8_t* used;
  size_t mask;
} I64Map;

static inline size_t i64_hash(int64_t k) {
  uint64_t x = (uint64_t)k;
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (size_t)x;
}

static int i64map_init(I64Map* m, size_t n_keys) {
  size_t cap = 16;
  while (cap < 2 * n_keys + 2) cap <<= 1;
  m->key = (int64_t*)malloc(cap * sizeof(int64_t));
  m->val = (int32_t*)malloc(cap * sizeof(int32_t));
  m->used = (uint8_t*)calloc(cap, 1);
  m->mask = cap - 1;
  return (m->key && m->val && m->used) ? 0 : -1;
}

static void i64map_free(I64Map* m) { free(m->key); free(m->val); free(m->used); }

/* returns the slot of k; *found says whether it was there (an empty slot otherwise) */
static inline size_t i64map_slot(const I64Map* m, int64_t k, int* found) {
  size_t h = i64_hash(k) & m->mask;
  while (m->used[h]) {
    if (m->key[h] == k) { *found = 1; return h; }
    h = (h + 1) & m->mask;
  }
  *found = 0;
  return h;
}

/* collect_points(keyframes) -> list: the points of every keyframe's get_all_mappoints() dictionary in order,
 * deduplicated by identity, first occurrence kept (Map.optimize_local's collection, map/map.py:159-169);
 * None when a container is not an exact dict. */
static PyObject* collect_points(PyObject* self, PyObject* args) {
  PyObject* keyframes;
  if (!PyArg_ParseTuple(args, "O", &keyframes)) return NULL;
  if (!PyList_CheckExact(keyframes)) Py_RETURN_NONE;
  static PyObject* s_all = NULL;
  if (!s_all) s_all = PyUnicode_InternFromString("get_all_mappoints");
  if (!s_all) return NULL;
  const Py_ssize_t n_kf = PyList_GET_SIZE(keyframes);
  PyObject* dicts = PyList_New(0);                 /* keeps the dictionaries (and so their values) alive */
  PyObject* out = NULL;
  if (!dicts) return NULL;
  size_t total = 0;
  for (Py_ssize_t i = 0; i < n_kf; ++i) {
    PyObject* d = PyObject_CallMethodNoArgs(PyList_GET_ITEM(keyframes, i), s_all);
    if (!d) { Py_DECREF(dicts); return NULL; }
    if (!PyDict_CheckExact(d)) { Py_DECREF(d); Py_DECREF(dicts); Py_RETURN_NONE; }
    total += (size_t)PyDict_GET_SIZE(d);
    const int rc = PyList_Append(dicts, d);
    Py_DECREF(d);
    if (rc != 0) { Py_DECREF(dicts); return NULL; }
  }
  /* growable open-addressing set of object addresses (0 = empty): sized by the distinct points, not by the walk */
  size_t cap = 1024, used = 0;
  uintptr_t* set = (uintptr_t*)calloc(cap, sizeof(uintptr_t));
  if (!set) { Py_DECREF(dicts); return PyErr_NoMemory(); }
  (void)total;
  out = PyList_New(0);
  if (!out) goto done;
  for (Py_ssize_t i = 0; i < n_kf; ++i) {
    PyObject* d = PyList_GET_ITEM(dicts, i);
    Py_ssize_t pos = 0;
    PyObject *k, *mp;
    while (PyDict_Next(d, &pos, &k, &mp)) {
      const uintptr_t key = (uintptr_t)mp;
      size_t h = i64_hash((int64_t)key) & (cap - 1);
      while (set[h] && set[h] != key) h = (h + 1) & (cap - 1);
      if (set[h]) continue;
      set[h] = key;
      if (PyList_Append(out, mp) != 0) { Py_CLEAR(out); goto done; }
      if (2 * ++used > cap) {
        const size_t cap2 = 2 * cap;
        uintptr_t* s2 = (uintptr_t*)calloc(cap2, sizeof(uintptr_t));
        if (!s2) { Py_CLEAR(out); PyErr_NoMemory(); goto done; }
        for (size_t q = 0; q < cap; ++q) {
          if (!set[q]) continue;
          size_t g = i64_hash((int64_t)set[q]) & (cap2 - 1);
          while (s2[g]) g = (g + 1) & (cap2 - 1);
          s2[g] = set[q];
        }
        free(set);
        set = s2;
        cap = cap2;
      }
    }
  }
done:
  free(set);
  Py_DECREF(dicts);
  return out;
}

/* table_select(kf i64[n], mp i64[n], cam i32[n], kp i64[n], alive u8[n], grp i64[n], n,
 *              kf_ids i64[n_cam], mp_keys i64[n_pt], n_sens i64[n_cam], base i64[n_cam*S], count i64[n_cam*S], S,
 *              uv_all f32[M*2])
 *   -> (obs_cam i32, obs_cam_id i32, obs_uv f32, pt_offsets i32[n_pt+1]) as bytearrays, observations grouped by
 *      window point and in table (= insertion) order inside a point;
 *      None when the selection needs the general path: a point key listed twice, or a (point, keyframe) pair seen
 *      from more than one sensor with another pair in between (ObservationTable.pack then sorts with numpy).
 * The selection and skip rules are those of ObservationTable.pack (packing.py), i.e. of the reference's packing
 * loop (torch_optimizer.py:80-95). */
static PyObject* table_select(PyObject* self, PyObject* args) {
  Py_buffer b_kf, b_mp, b_cam, b_kp, b_alive, b_grp, b_kfids, b_keys, b_ns, b_base, b_count, b_uv;
  Py_ssize_t n, S;
  if (!PyArg_ParseTuple(args, "y*y*y*y*y*y*ny*y*y*y*y*ny*", &b_kf, &b_mp, &b_cam, &b_kp, &b_alive, &b_grp, &n, &b_kfids,
                        &b_keys, &b_ns, &b_base, &b_count, &S, &b_uv))
    return NULL;
  PyObject* result = NULL;
  const int64_t* kf = (const int64_t*)b_kf.buf;
  const int64_t* mp = (const int64_t*)b_mp.buf;
  const int32_t* cam = (const int32_t*)b_cam.buf;
  const int64_t* kp = (const int64_t*)b_kp.buf;
  const uint8_t* alive = (const uint8_t*)b_alive.buf;
  const int64_t* grp = (const int64_t*)b_grp.buf;
  const int64_t* kf_ids = (const int64_t*)b_kfids.buf;
  const int64_t* keys = (const int64_t*)b_keys.buf;
  const int64_t* n_sens = (const int64_t*)b_ns.buf;
  const int64_t* base = (const int64_t*)b_base.buf;
  const int64_t* count = (const int64_t*)b_count.buf;
  const float* uv = (const float*)b_uv.buf;
  const Py_ssize_t n_cam = b_kfids.len / (Py_ssize_t)sizeof(int64_t), n_pt = b_keys.len / (Py_ssize_t)sizeof(int64_t);
  const Py_ssize_t n_uv = b_uv.len / (Py_ssize_t)(2 * sizeof(float));
  I64Map cmap = {0}, pmap = {0};
  int32_t *s_pt = NULL, *s_cam = NULL, *s_cid = NULL, *cnt = NULL;
  int64_t* s_src = NULL;
  int general = 0;
  if (n < 0 || S < 1 || b_kf.len < n * 8 || b_mp.len < n * 8 || b_cam.len < n * 4 || b_kp.len < n * 8 || b_alive.len < n ||
      b_grp.len < n * 8 || b_ns.len < n_cam * 8 || b_base.len < n_cam * S * 8 || b_count.len < n_cam * S * 8) {
    PyErr_SetString(PyExc_ValueError, "table_select: buffer sizes");
    goto out;
  }
  if (i64map_init(&cmap, (size_t)n_cam) != 0 || i64map_init(&pmap, (size_t)n_pt) != 0) { PyErr_NoMemory(); goto out; }
  for (Py_ssize_t i = 0; i < n_cam; ++i) {            /* a later duplicate keyframe id wins, as the dict at :61 does */
    int found;
    const size_t h = i64map_slot(&cmap, kf_ids[i], &found);
    cmap.used[h] = 1; cmap.key[h] = kf_ids[i]; cmap.val[h] = (int32_t)i;
  }
  for (Py_ssize_t j = 0; j < n_pt; ++j) {
    int found;
    const size_t h = i64map_slot(&pmap, keys[j], &found);
    if (found) { general = 1; break; }                /* a point handed in twice */
    pmap.used[h] = 1; pmap.key[h] = keys[j]; pmap.val[h] = (int32_t)j;
  }
  if (general) { result = Py_None; Py_INCREF(result); goto out; }
  s_pt = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  s_cam = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  s_cid = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  s_src = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  cnt = (int32_t*)calloc((size_t)n_pt + 2, sizeof(int32_t));
  if (!s_pt || !s_cam || !s_cid || !s_src || !cnt) { PyErr_NoMemory(); goto out; }
  {
    size_t m = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
      if (!alive[i]) continue;
      int found;
      size_t h = i64map_slot(&pmap, mp[i], &found);
      if (!found) continue;
      const int32_t p = pmap.val[h];
      h = i64map_slot(&cmap, kf[i], &found);
      if (!found) continue;
      const int32_t c = cmap.val[h];
      /* the packer's skip rules (:90-93), with Python's negative indexing as kps[cam_id][kp_idx] would apply it */
      const int64_t ns = n_sens[c];
      int64_t cid = cam[i];
      if (!(cid < ns && cid >= -ns)) continue;
      const int64_t cc = cid < 0 ? cid + ns : cid;
      if (cc >= S) continue;
      const int64_t nk = count[c * S + cc];
      int64_t k = kp[i];
      if (!(k < nk && k >= -nk)) continue;
      if (k < 0) k += nk;
      if (grp[i] != (int64_t)i) { general = 1; break; }      /* second sensor of a (point, keyframe) pair */
      const int64_t src = base[c * S + cc] + k;
      if (src < 0 || src >= n_uv) { PyErr_SetString(PyExc_IndexError, "table_select: keypoint offset"); goto out; }
      s_pt[m] = p; s_cam[m] = c; s_cid[m] = (int32_t)cam[i]; s_src[m] = src;
      cnt[p + 1]++;
      ++m;
    }
    if (general) { result = Py_None; Py_INCREF(result); goto out; }
    for (Py_ssize_t j = 0; j < n_pt; ++j) cnt[j + 1] += cnt[j];        /* offsets */
    PyObject* o_cam = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(m * sizeof(int32_t)));
    PyObject* o_cid = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(m * sizeof(int32_t)));
    PyObject* o_uv = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(m * 2 * sizeof(float)));
    PyObject* o_off = PyByteArray_FromStringAndSize((const char*)cnt, (Py_ssize_t)(((size_t)n_pt + 1) * sizeof(int32_t)));
    if (o_cam && o_cid && o_uv && o_off) {
      int32_t* d_cam = (int32_t*)PyByteArray_AS_STRING(o_cam);
      int32_t* d_cid = (int32_t*)PyByteArray_AS_STRING(o_cid);
      float* d_uv = (float*)PyByteArray_AS_STRING(o_uv);
      for (size_t r = 0; r < m; ++r) {                                 /* stable counting sort by point */
        const int32_t at = cnt[s_pt[r]]++;
        d_cam[at] = s_cam[r];
        d_cid[at] = s_cid[r];
        d_uv[2 * at] = uv[2 * s_src[r]];
        d_uv[2 * at + 1] = uv[2 * s_src[r] + 1];
      }
      result = PyTuple_Pack(4, o_cam, o_cid, o_uv, o_off);
    }
    Py_XDECREF(o_cam); Py_XDECREF(o_cid); Py_XDECREF(o_uv); Py_XDECREF(o_off);
  }
out:
  free(s_pt); free(s_cam); free(s_cid); free(s_src); free(cnt);
  i64map_free(&cmap); i64map_free(&pmap);
  PyBuffer_Release(&b_kf); PyBuffer_Release(&b_mp); PyBuffer_Release(&b_cam); PyBuffer_Release(&b_kp);
  PyBuffer_Release(&b_alive); PyBuffer_Release(&b_grp); PyBuffer_Release(&b_kfids); PyBuffer_Release(&b_keys);
  PyBuffer_Release(&b_ns); PyBuffer_Release(&b_base); PyBuffer_Release(&b_count); PyBuffer_Release(&b_uv);
  return result;
}

static PyMethodDef methods[] = {
    {"pack_observations", pack_observations, METH_VARARGS,
     "pack_observations(map_points, pose_index, keypoints_of) -> (cam i32, cam_id i32, uv f32, counts i64) bytes, "
     "or None when a container is not an exact dict / list"},
    {"gather_points", gather_points, METH_VARARGS,
     "gather_points(map_points) -> (positions f64 bytes, keys i64 bytes), or None when a position is not three numbers"},
    {"collect_points", collect_points, METH_VARARGS, "collect_points(keyframes) -> list of the window's points, first occurrence kept; None for foreign containers"},
    {"table_select", table_select, METH_VARARGS, "selection + counting sort of ObservationTable.pack; None when the general (numpy) path is needed"},
    {"scatter_points", scatter_points, METH_VARARGS, "scatter_points(map_points, X) -> True after mp.update_position(X[j]) for all j"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "window packer inner loop", -1, methods};

PyMODINIT_FUNC PyInit__fastpack(void) { return PyModule_Create(&moddef); }
