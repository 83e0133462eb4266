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

/* scatter_points(map_points, X) -> True: mp.update_position(X[j]) for every point, X any sequence of rows
 * (a C-contiguous (n,3) float64 numpy array hands out row views).  The write-back of torch_optimizer.py:201-203. */
static PyObject* scatter_points(PyObject* self, PyObject* args) {
  PyObject *map_points, *X;
  if (!PyArg_ParseTuple(args, "OO", &map_points, &X)) return NULL;
  if (!PyList_CheckExact(map_points) || !PySequence_Check(X)) Py_RETURN_FALSE;
  static PyObject* s_upd = NULL;
  if (!s_upd) s_upd = PyUnicode_InternFromString("update_position");
  if (!s_upd) return NULL;
  const Py_ssize_t n = PyList_GET_SIZE(map_points);
  if (PySequence_Size(X) < n) { PyErr_SetString(PyExc_ValueError, "fewer rows than points"); return NULL; }
  for (Py_ssize_t j = 0; j < n; ++j) {
    PyObject* row = PySequence_GetItem(X, j);
    if (!row) return NULL;
    PyObject* r = PyObject_CallMethodOneArg(PyList_GET_ITEM(map_points, j), s_upd, row);
    Py_DECREF(row);
    if (!r) return NULL;
    Py_DECREF(r);
  }
  Py_RETURN_TRUE;
}

static PyMethodDef methods[] = {
    {"pack_observations", pack_observations, METH_VARARGS,
     "pack_observations(map_points, pose_index, keypoints_of) -> (cam i32, cam_id i32, uv f32, counts i64) bytes, "
     "or None when a container is not an exact dict / list"},
    {"gather_points", gather_points, METH_VARARGS,
     "gather_points(map_points) -> (positions f64 bytes, keys i64 bytes), or None when a position is not three numbers"},
    {"scatter_points", scatter_points, METH_VARARGS, "scatter_points(map_points, X) -> True after mp.update_position(X[j]) for all j"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "window packer inner loop", -1, methods};

PyMODINIT_FUNC PyInit__fastpack(void) { return PyModule_Create(&moddef); }
