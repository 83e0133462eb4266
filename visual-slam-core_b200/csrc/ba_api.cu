// C ABI (include/b200ba.h): handle, workspace, LM loop orchestration, CUDA-graph capture,
// host-pointer variants, NCCL plumbing for the shard-by-point multi-GPU path.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/b200ba.h"
#include "ba_kernels.cuh"

using namespace ba;

// ------------------------------------------------------------------------------------------
// NCCL through dlopen: the single-GPU path must not depend on libnccl being present
// ------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm*, int, nccl_uid, int);
typedef int (*fn_destroy)(nccl_comm);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t);
typedef int (*fn_broadcast)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef const char* (*fn_errstr)(int);
struct NcclApi {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_destroy destroy = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_allgather allgather = nullptr;
  fn_broadcast broadcast = nullptr;
  fn_errstr errstr = nullptr;
};
NcclApi g_nccl;
bool load_nccl() {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    void* h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!h) continue;
    g_nccl.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
    g_nccl.allgather = (fn_allgather)dlsym(h, "ncclAllGather");
    g_nccl.broadcast = (fn_broadcast)dlsym(h, "ncclBroadcast");
    g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    if (g_nccl.get_uid && g_nccl.init_rank && g_nccl.destroy && g_nccl.allreduce && g_nccl.allgather && g_nccl.broadcast) {
      g_nccl.lib = h;
      return true;
    }
    dlclose(h);
  }
  return false;
}
constexpr int kNcclDouble = 8, kNcclSum = 0, kNcclMax = 2, kNcclUint32 = 3, kNcclInt32 = 2;
}  // namespace

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct ba_handle {
  int device = 0;
  int sm_count = 148;
  ba_params prm;
  std::string err;
  // workspace
  DevBuf camstate, X0, X1, red, ctl, cand, dbg, part;
  DevBuf rec, cs_cnt, cs_blk_off, cs_blk_cam, cs_pt, cs_uv, cs_nblk;   // panel linearise path
  DevBuf ro_cnt, ro_perm, ro_off, ro_cam, ro_pt, ro_uv, ro_range;      // panel path: class order of the points, the shard's camera range   // camstate = pose[0] | pose[1] | dp, contiguous (one broadcast)
  // sparse path: structure + PCG workspace
  DevBuf bits, wpre, row_nnz, row_ptr, col_idx, tpos, pcg, bits_all;
  // banded path: factor columns, forward-substituted rhs, inverse pivots, bandwidth scalar
  DevBuf band_L, band_y, band_dinv, band_wbuf, bcr;
  int last_solver = BA_SOLVER_CHOLESKY, last_band_w = -1, last_s_blocks = 0;
  int* h_nnzb = nullptr;          // pinned
  // staging for the host-pointer API
  DevBuf d_R, d_t, d_fixed, d_X, d_ocam, d_opt, d_ouv, d_off, d_flags, d_hist, d_acc, d_foff;
  LmCtl* h_ctl = nullptr;         // pinned
  double* h_small = nullptr;      // pinned staging of R | t | fixed for the host-pointer API
  size_t h_small_cap = 0;         // bytes
  DevBuf d_small;
  cudaStream_t stream = nullptr;  // used by the host API
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_val = nullptr;
  DevBuf d_bad;                   // verdict of k_validate
  // CUDA graph cache of the LM loop
  cudaGraphExec_t graph = nullptr;
  struct GraphKey {
    int n_cam = -1, n_pt = -1, n_obs = -1, max_iter = -1, schur = -1, world = -1;
    const void* ptrs[10] = {0};
    double scal[12] = {0};
    bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
  } graph_key;
  int graph_launches = 0;
  // multi-GPU
  nccl_comm comm = nullptr;
  int rank = 0, world = 1;
  // peer-memory exchange of the reduction buffer (banded path): CUDA IPC mappings of every rank's `red` and flag array
  struct Peer {
    bool open = false, off = false;
    void* red_exported = nullptr;            // the `red` pointer the peers have mapped
    void* red[8] = {nullptr};                // every rank's reduction buffer (own at [rank])
    void* flags[8] = {nullptr};              // every rank's flag array: flags[q][r] = last epoch rank r posted to rank q
    DevBuf flagbuf, xbuf;                    // own flag array; scratch for the handle exchange
    unsigned long long solves = 0;           // epoch base = 64 * solves (all ranks call in lock step)
  } peer;
  // profile mode: events between kernels, tag of the kernel that follows each event
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_tag;
  int prof_n = 0;
  float prof_ms[7] = {0, 0, 0, 0, 0, 0, 0};
};

enum { TAG_LIN = 0, TAG_SOLVE = 1, TAG_BACK = 2, TAG_OTHER = 3, TAG_COMM = 4, TAG_LIN_PB = 5, TAG_LIN_CAM = 6 };

static void prof_mark(ba_handle* h, int tag, cudaStream_t s) {
  if (!h->prm.profile) return;
  if (h->prof_n >= (int)h->prof_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->prof_ev.push_back(e);
    h->prof_tag.push_back(0);
  }
  h->prof_tag[h->prof_n] = tag;
  cudaEventRecord(h->prof_ev[h->prof_n], s);
  h->prof_n++;
}

// reserve an event slot (recorded by the callee) that closes the interval of `tag_before` and opens one of `tag`
static cudaEvent_t prof_slot(ba_handle* h, int tag) {
  if (h->prof_n >= (int)h->prof_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->prof_ev.push_back(e);
    h->prof_tag.push_back(0);
  }
  h->prof_tag[h->prof_n] = tag;
  return h->prof_ev[h->prof_n++];
}

static void prof_collect(ba_handle* h) {
  for (int i = 0; i < 7; ++i) h->prof_ms[i] = 0.f;
  for (int i = 0; i + 1 < h->prof_n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]) == cudaSuccess)
      h->prof_ms[h->prof_tag[i]] += ms;
  }
}

// cudaSetDevice for the duration of an API call; the caller's current device comes back on every exit path
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
      return BA_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

static int ensure(ba_handle* h, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap && b.p) return BA_OK;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  size_t cap = bytes + bytes / 4 + 256;
  CK(cudaMalloc(&b.p, cap));
  b.cap = cap;
  if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; h->graph_key = ba_handle::GraphKey(); }
  return BA_OK;
}

// ---- peer-memory exchange (sharded banded path) --------------------------------------------------------------
static bool peer_wanted(const ba_handle* h) {
  static int v = -1;                 // diagnostics: BA_PEER_REDUCE=0 keeps the NCCL all-reduce in front of the solver
  if (v < 0) { const char* e = getenv("BA_PEER_REDUCE"); v = (e && !atoi(e)) ? 0 : 1; }
  return v != 0 && h->world > 1 && h->world <= kMaxPeers && !h->peer.off;
}

// unmap the peers' buffers; every rank gets here in the same call (the buffer sizes depend on the window only), and the
// collective makes sure nobody frees a buffer a peer still has mapped
static int peer_close(ba_handle* h, cudaStream_t s) {
  if (!h->peer.open) return BA_OK;
  for (int q = 0; q < h->world; ++q) {
    if (q == h->rank) continue;
    if (h->peer.red[q]) cudaIpcCloseMemHandle(h->peer.red[q]);
    if (h->peer.flags[q]) cudaIpcCloseMemHandle(h->peer.flags[q]);
    h->peer.red[q] = h->peer.flags[q] = nullptr;
  }
  h->peer.open = false;
  h->peer.red_exported = nullptr;
  if (h->comm) {
    if (ensure(h, h->peer.xbuf, 4096)) return BA_ERR_CUDA;
    int r = g_nccl.allreduce(h->peer.xbuf.p, h->peer.xbuf.p, 1, kNcclInt32, kNcclSum, h->comm, s);
    if (r != 0) { h->err = "ncclAllReduce(peer close)"; return BA_ERR_NCCL; }
    CK(cudaStreamSynchronize(s));
  }
  return BA_OK;
}

// map every rank's reduction buffer and flag array (IPC handles travel through an NCCL all-gather); the ranks agree
// on the outcome, a failure anywhere switches the peer path off everywhere (NCCL all-reduce instead)
static int peer_open(ba_handle* h, cudaStream_t s) {
  const int W = h->world;
  int rc;
  if ((rc = ensure(h, h->peer.xbuf, 4096))) return rc;
  if (!h->peer.flagbuf.p) {
    if ((rc = ensure(h, h->peer.flagbuf, sizeof(long long) * 64))) return rc;
    CK(cudaMemset(h->peer.flagbuf.p, 0, sizeof(long long) * 64));
  }
  struct Pair { cudaIpcMemHandle_t red, flags; };
  static_assert(sizeof(Pair) == 128, "IPC handle size");
  std::vector<Pair> all(W);
  Pair mine;
  int ok = 1;
  if (cudaIpcGetMemHandle(&mine.red, h->red.p) != cudaSuccess) ok = 0;
  if (cudaIpcGetMemHandle(&mine.flags, h->peer.flagbuf.p) != cudaSuccess) ok = 0;
  cudaGetLastError();
  char* xb = (char*)h->peer.xbuf.p;                      // [0, 128): mine; [256, 256 + 128 W): everyone's
  CK(cudaMemcpyAsync(xb, &mine, sizeof(Pair), cudaMemcpyHostToDevice, s));
  int r = g_nccl.allgather(xb, xb + 256, sizeof(Pair), /*ncclChar*/ 0, h->comm, s);
  if (r != 0) { h->err = "ncclAllGather(IPC handles)"; return BA_ERR_NCCL; }
  CK(cudaMemcpyAsync(all.data(), xb + 256, sizeof(Pair) * W, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int q = 0; q < W && ok; ++q) {
    if (q == h->rank) { h->peer.red[q] = h->red.p; h->peer.flags[q] = h->peer.flagbuf.p; continue; }
    if (cudaIpcOpenMemHandle(&h->peer.red[q], all[q].red, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
    else if (cudaIpcOpenMemHandle(&h->peer.flags[q], all[q].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
  }
  cudaGetLastError();
  // agreement: min over the ranks
  int* d_ok = (int*)(xb + 2048);
  CK(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, s));
  r = g_nccl.allreduce(d_ok, d_ok, 1, kNcclInt32, /*ncclMin*/ 3, h->comm, s);
  if (r != 0) { h->err = "ncclAllReduce(peer agreement)"; return BA_ERR_NCCL; }
  int all_ok = 0;
  CK(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  h->peer.open = true;                                   // so that peer_close unmaps whatever was mapped
  h->peer.red_exported = h->red.p;
  if (!all_ok) {
    h->peer.off = true;
    return peer_close(h, s);
  }
  return BA_OK;
}

// unmap without a collective (tear-down: the ranks do not arrive together)
static void peer_unmap(ba_handle* h) {
  for (int q = 0; q < kMaxPeers; ++q) {
    if (q == h->rank || !h->peer.open) continue;
    if (h->peer.red[q]) cudaIpcCloseMemHandle(h->peer.red[q]);
    if (h->peer.flags[q]) cudaIpcCloseMemHandle(h->peer.flags[q]);
  }
  for (int q = 0; q < kMaxPeers; ++q) h->peer.red[q] = h->peer.flags[q] = nullptr;
  h->peer.open = false;
  h->peer.red_exported = nullptr;
}

static PeerView peer_view(const ba_handle* h) {
  PeerView pv;
  memset(&pv, 0, sizeof(pv));
  pv.world = h->world; pv.rank = h->rank;
  for (int q = 0; q < h->world; ++q) {
    pv.red[q] = (const double*)h->peer.red[q];
    pv.flag_remote[q] = (long long*)h->peer.flags[q] + 4 * h->rank;
  }
  pv.flag_local = (const long long*)h->peer.flagbuf.p;
  pv.cam_range = (const int*)h->ro_range.p;
  return pv;
}

namespace ba { int mma_max_active_clusters(int cluster); }
extern "C" int ba_debug_max_active_clusters(int cluster) { return ba::mma_max_active_clusters(cluster); }

extern "C" int ba_version(void) { return 100; }

extern "C" const char* ba_error_string(int code) {
  switch (code) {
    case BA_OK: return "ok";
    case BA_ERR_INVALID: return "invalid argument";
    case BA_ERR_CUDA: return "CUDA error";
    case BA_ERR_TOO_LARGE: return "problem too large for the selected solver";
    case BA_ERR_NCCL: return "NCCL error";
    case BA_ERR_NO_DEVICE: return "no CUDA device (this solver has no CPU fallback)";
    default: return "unknown error";
  }
}

static std::string g_create_err;   // ba_create has no handle to hang its message on
extern "C" const char* ba_last_error(const ba_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" void ba_default_params(ba_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->huber_delta = 5.0;
  p->chi2_thr = 5.991;
  p->lambda0 = 1e-4;
  p->lambda_up = 10.0;
  p->lambda_down = 0.1;
  p->lambda_min = 1e-6;
  p->lambda_max = 1e8;
  p->accept_rtol = 1e-8;
  p->max_iter = 10;
  p->solver_kind = BA_SOLVER_AUTO;
  p->schur_kind = BA_SCHUR_AUTO;
  p->pcg_max_iter = 500;
  p->pcg_tol = 1e-10;
  p->use_graph = 1;
  p->profile = 0;
}

static int check_params(const ba_params* p) {
  if (!p) return BA_ERR_INVALID;
  if (p->max_iter < 0 || p->max_iter > BA_MAX_ITER) return BA_ERR_INVALID;
  if (!(p->huber_delta > 0) || !(p->lambda0 > 0) || !(p->lambda_up >= 1) || !(p->lambda_down > 0))
    return BA_ERR_INVALID;
  return BA_OK;
}

extern "C" int ba_create(const ba_params* params, int device, ba_handle** out) {
  if (!out) return BA_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return BA_ERR_NO_DEVICE;
  if (device < 0 || device >= ndev) return BA_ERR_INVALID;
  ba_params p;
  if (params) p = *params; else ba_default_params(&p);
  if (check_params(&p) != BA_OK) return BA_ERR_INVALID;
  ba_handle* h = new ba_handle();
  h->device = device;
  h->prm = p;
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  cudaError_t ce = cudaSetDevice(device);
  if (ce != cudaSuccess) { g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(ce); delete h; return BA_ERR_CUDA; }
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev == device ? -1 : prev_dev};
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  int ke = configure_kernels();
  if (ke != 0) {
    g_create_err = std::string("kernel attributes: ") + cudaGetErrorString((cudaError_t)ke);
    cudaGetLastError();
    delete h;
    return BA_ERR_CUDA;
  }
  if ((ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (ce = cudaEventCreate(&h->ev0)) != cudaSuccess || (ce = cudaEventCreate(&h->ev1)) != cudaSuccess ||
      (ce = cudaEventCreateWithFlags(&h->ev_val, cudaEventDisableTiming)) != cudaSuccess ||
      (ce = cudaMallocHost((void**)&h->h_ctl, sizeof(LmCtl))) != cudaSuccess ||
      (ce = cudaMallocHost((void**)&h->h_nnzb, sizeof(int))) != cudaSuccess) {
    g_create_err = std::string("handle resources: ") + cudaGetErrorString(ce);
    delete h;
    return BA_ERR_CUDA;
  }
  *out = h;
  return BA_OK;
}

extern "C" int ba_set_params(ba_handle* h, const ba_params* params) {
  if (!h || check_params(params) != BA_OK) return BA_ERR_INVALID;
  // the captured LM loop depends on the solver / Schur kind, the iteration count and the PCG settings; the LM
  // scalars (damping, Huber delta, chi2 threshold ...) travel with k_init outside the graph
  const bool regraph = params->max_iter != h->prm.max_iter || params->solver_kind != h->prm.solver_kind ||
                       params->schur_kind != h->prm.schur_kind || params->use_graph != h->prm.use_graph ||
                       params->profile != h->prm.profile || params->pcg_max_iter != h->prm.pcg_max_iter ||
                       params->pcg_tol != h->prm.pcg_tol;
  h->prm = *params;
  if (regraph && h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; h->graph_key = ba_handle::GraphKey(); }
  return BA_OK;
}

extern "C" int ba_destroy(ba_handle* h) {
  if (!h) return BA_OK;
  DeviceGuard dev_guard(h->device);
  if (h->graph) cudaGraphExecDestroy(h->graph);
  DevBuf* bufs[] = {&h->camstate, &h->X0, &h->X1, &h->red, &h->ctl, &h->cand, &h->rec, &h->cs_cnt, &h->cs_blk_off, &h->cs_blk_cam, &h->cs_pt, &h->cs_uv, &h->cs_nblk,
                    &h->d_R, &h->d_t, &h->d_fixed, &h->d_X, &h->d_ocam, &h->d_opt, &h->d_ouv,
                    &h->d_off, &h->d_flags, &h->d_hist, &h->d_acc, &h->d_foff, &h->dbg, &h->part,
                    &h->bits, &h->wpre, &h->row_nnz, &h->row_ptr, &h->col_idx, &h->tpos, &h->pcg,
                    &h->bits_all, &h->band_L, &h->band_y, &h->band_dinv, &h->band_wbuf, &h->bcr,
                    &h->ro_cnt, &h->ro_perm, &h->ro_off, &h->ro_cam, &h->ro_pt, &h->ro_uv,
                    &h->peer.flagbuf, &h->peer.xbuf, &h->ro_range, &h->d_bad};
  peer_unmap(h);
  for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
  if (h->comm && g_nccl.destroy) g_nccl.destroy(h->comm);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->h_small) cudaFreeHost(h->h_small);
  if (h->d_small.p) cudaFree(h->d_small.p);
  if (h->h_nnzb) cudaFreeHost(h->h_nnzb);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_val) cudaEventDestroy(h->ev_val);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return BA_OK;
}

// ------------------------------------------------------------------------------------------
// local BA
// ------------------------------------------------------------------------------------------
enum { MODE_DENSE = 0, MODE_PCG = 1, MODE_BAND = 2, MODE_NONE = 3 };   // NONE: no reduced camera matrix (cost evaluation)

// co-visibility bandwidth in keyframes (max over points of the camera-index spread); one host sync
static int band_width(ba_handle* h, int n_pt, const int32_t* obs_cam, const int32_t* pt_off,
                      cudaStream_t s, int* out) {
  int rc;
  if ((rc = ensure(h, h->band_wbuf, sizeof(int)))) return rc;
  Work w;
  memset(&w, 0, sizeof(w));
  w.n_pt = n_pt; w.obs_cam = obs_cam; w.pt_off = pt_off;
  launch_band_width(w, (int*)h->band_wbuf.p, s);
  if (h->world > 1) {
    int r = g_nccl.allreduce(h->band_wbuf.p, h->band_wbuf.p, 1, kNcclInt32, kNcclMax, h->comm, s);
    if (r != 0) { h->err = std::string("ncclAllReduce(bandwidth): ") + (g_nccl.errstr ? g_nccl.errstr(r) : "?"); return BA_ERR_NCCL; }
  }
  CK(cudaMemcpyAsync(h->h_nnzb, h->band_wbuf.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  *out = *h->h_nnzb;
  return BA_OK;
}

// which layout / solver this window gets
static int pick_mode(ba_handle* h, int n_cam, int n_pt, const int32_t* obs_cam, const int32_t* pt_off,
                     cudaStream_t s, int* mode, int* bw) {
  const int kind = h->prm.solver_kind;
  const bool dense_fits = 6 * n_cam <= rcs_cholesky_max_n();
  *bw = -1;
  if (kind == BA_SOLVER_PCG) { *mode = MODE_PCG; return BA_OK; }
  if (kind == BA_SOLVER_CHOLESKY || (kind == BA_SOLVER_AUTO && dense_fits)) {
    if (!dense_fits) {
      h->err = "reduced camera system larger than the in-shared-memory Cholesky; use solver auto, band or pcg";
      return BA_ERR_TOO_LARGE;
    }
    *mode = MODE_DENSE;
    return BA_OK;
  }
  int wv = 0;
  int rc = band_width(h, n_pt, obs_cam, pt_off, s, &wv);
  if (rc) return rc;
  if (wv >= n_cam) wv = n_cam - 1;
  if (wv <= band_max_w()) { *mode = MODE_BAND; *bw = wv; return BA_OK; }
  if (kind == BA_SOLVER_BAND || kind == BA_SOLVER_BAND_SEQ) {
    h->err = "co-visibility bandwidth exceeds the banded solver (32 keyframes); use solver auto or pcg";
    return BA_ERR_TOO_LARGE;
  }
  *mode = MODE_PCG;
  return BA_OK;
}

// workspace + kernel view.  sparse = block-sparse S + PCG (needs one host sync for the block count)
static int make_work(ba_handle* h, int n_cam, int n_pt, int n_obs, const double intr[4],
                     const uint8_t* fixed, const int32_t* obs_cam, const int32_t* obs_pt,
                     const float* obs_uv, const int32_t* pt_off, Work& w, int mode = MODE_DENSE,
                     cudaStream_t s = nullptr, int bw = -1) {
  const int n = 6 * n_cam;
  int rc;
  memset(&w, 0, sizeof(w));
  w.band_w = -1;
  if ((rc = ensure(h, h->camstate, sizeof(double) * (30 * (size_t)n_cam + 2)))) return rc;
  if ((rc = ensure(h, h->X0, sizeof(double) * 3 * (size_t)(n_pt + 1)))) return rc;
  if ((rc = ensure(h, h->X1, sizeof(double) * 3 * (size_t)(n_pt + 1)))) return rc;
  if ((rc = ensure(h, h->ctl, sizeof(LmCtl)))) return rc;
  if ((rc = ensure(h, h->cand, sizeof(double) * 2))) return rc;
  w.n_cam = n_cam; w.n_pt = n_pt; w.n_obs = n_obs; w.n = n;
  w.K = Intr{intr[0], intr[1], intr[2], intr[3]};
  w.obs_cam = obs_cam; w.obs_pt = obs_pt; w.obs_uv = obs_uv; w.pt_off = pt_off; w.fixed = fixed;
  w.pose[0] = (double*)h->camstate.p; w.pose[1] = w.pose[0] + 12 * (size_t)n_cam;
  w.X[0] = (double*)h->X0.p; w.X[1] = (double*)h->X1.p;
  w.cost_cand = (double*)h->cand.p;
  w.dp = w.pose[1] + 12 * (size_t)n_cam;
  w.ctl = (LmCtl*)h->ctl.p;
  w.world = h->world;
  if (pt_off != nullptr && n_pt > 0) {
    w.part_n = mma_grid(n_cam, n_pt, h->sm_count);
    if ((rc = ensure(h, h->part, sizeof(int) * (size_t)(w.part_n + 1)))) return rc;
    w.part = (int*)h->part.p;
    if (n_obs > 0 && (h->prm.schur_kind == BA_SCHUR_AUTO || h->prm.schur_kind == BA_SCHUR_PANEL)) {
      // panel path: one record per point, camera-sorted copy of the observations (cameras padded to 32 entries)
      w.cs_nblk_max = panel_cs_blocks_max(n_cam, n_obs);
      if ((rc = ensure(h, h->rec, sizeof(double) * panel_rec_doubles(n_pt)))) return rc;
      if ((rc = ensure(h, h->cs_cnt, sizeof(int) * (size_t)(n_cam + 1)))) return rc;
      if ((rc = ensure(h, h->cs_blk_off, sizeof(int) * (size_t)(n_cam + 2)))) return rc;
      if ((rc = ensure(h, h->cs_blk_cam, sizeof(int) * (size_t)(w.cs_nblk_max + 1)))) return rc;
      if ((rc = ensure(h, h->cs_pt, sizeof(int) * 32 * (size_t)(w.cs_nblk_max + 1)))) return rc;
      if ((rc = ensure(h, h->cs_uv, sizeof(float) * 64 * (size_t)(w.cs_nblk_max + 1)))) return rc;
      if ((rc = ensure(h, h->cs_nblk, sizeof(int) * 4))) return rc;
      w.rec = (double*)h->rec.p;
      w.cs_cnt = (int*)h->cs_cnt.p; w.cs_blk_off = (int*)h->cs_blk_off.p; w.cs_blk_cam = (int*)h->cs_blk_cam.p;
      w.cs_pt = (int*)h->cs_pt.p; w.cs_uv = (float2*)h->cs_uv.p; w.cs_nblk = (int*)h->cs_nblk.p;
    }
  }
  w.dbg = nullptr;
  static int want_dbg = -1;
  if (want_dbg < 0) { const char* e = getenv("BA_DBG_TIMES"); want_dbg = (e && atoi(e)) ? 1 : 0; }
  if (want_dbg) {
    if ((rc = ensure(h, h->dbg, sizeof(long long) * 4096))) return rc;
    w.dbg = (long long*)h->dbg.p;
  }
  if (mode == MODE_NONE) {
    w.s_len = 0;
  } else if (mode == MODE_DENSE) {
    w.s_len = s_doubles(n_cam);
  } else if (mode == MODE_BAND) {
    w.band_w = bw;
    w.s_len = (size_t)n_cam * (bw + 1) * 36;
    if ((rc = ensure(h, h->band_L, sizeof(double) * w.s_len))) return rc;
    if ((rc = ensure(h, h->band_y, sizeof(double) * (size_t)n))) return rc;
    if ((rc = ensure(h, h->band_dinv, sizeof(double) * (size_t)n))) return rc;
    w.band_L = (double*)h->band_L.p;
    w.band_y = (double*)h->band_y.p;
    w.band_dinv = (double*)h->band_dinv.p;
    // cyclic reduction over 15-keyframe super-blocks when the band fits one (BA_BAND=seq keeps the sequential solver)
    static int band_mode = -1;
    if (band_mode < 0) { const char* e = getenv("BA_BAND"); band_mode = (e && !strcmp(e, "seq")) ? 1 : 0; }
    if (band_mode == 0 && h->prm.solver_kind != BA_SOLVER_BAND_SEQ && bcr_supported(n_cam, bw)) {
      if ((rc = ensure(h, h->bcr, sizeof(double) * bcr_workspace_doubles(n_cam)))) return rc;
      w.bcr_ws = (double*)h->bcr.p;
    }
  } else {
    // symbolic phase: co-visibility bitmap -> block-CSR structure
    const int W = (n_cam + 31) / 32;
    if ((rc = ensure(h, h->bits, sizeof(unsigned) * (size_t)n_cam * W))) return rc;
    if ((rc = ensure(h, h->wpre, sizeof(int) * (size_t)n_cam * W))) return rc;
    if ((rc = ensure(h, h->row_nnz, sizeof(int) * (size_t)(n_cam + 1)))) return rc;
    if ((rc = ensure(h, h->row_ptr, sizeof(int) * (size_t)(n_cam + 1)))) return rc;
    w.bsr_W = W;
    w.bsr_bits = (const unsigned*)h->bits.p;
    w.bsr_wpre = (const int*)h->wpre.p;
    w.bsr_row_ptr = (const int*)h->row_ptr.p;
    if (h->world > 1) {
      // every rank must build the same structure: OR the co-visibility bitmaps of all shards
      const size_t words = (size_t)n_cam * W;
      if ((rc = ensure(h, h->bits_all, sizeof(unsigned) * words * h->world))) return rc;
      launch_symbolic_mark(w, (unsigned*)h->bits.p, s);
      int r = g_nccl.allgather(h->bits.p, h->bits_all.p, words, kNcclUint32, h->comm, s);
      if (r != 0) { h->err = std::string("ncclAllGather(bitmap): ") + (g_nccl.errstr ? g_nccl.errstr(r) : "?"); return BA_ERR_NCCL; }
      launch_symbolic_or(w, (const unsigned*)h->bits_all.p, h->world, (unsigned*)h->bits.p, s);
      launch_symbolic_count(w, (unsigned*)h->bits.p, (int*)h->wpre.p, (int*)h->row_nnz.p, (int*)h->row_ptr.p, s);
    } else {
      launch_symbolic(w, (unsigned*)h->bits.p, (int*)h->wpre.p, (int*)h->row_nnz.p, (int*)h->row_ptr.p, s);
    }
    CK(cudaMemcpyAsync(h->h_nnzb, (int*)h->row_ptr.p + n_cam, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int nnzb = *h->h_nnzb;
    if (nnzb <= 0) { h->err = "symbolic phase produced no blocks"; return BA_ERR_CUDA; }
    if ((rc = ensure(h, h->col_idx, sizeof(int) * (size_t)nnzb))) return rc;
    if ((rc = ensure(h, h->tpos, sizeof(int) * (size_t)nnzb))) return rc;
    w.nnzb = nnzb;
    w.bsr_col = (const int*)h->col_idx.p;
    w.bsr_tpos = (const int*)h->tpos.p;
    launch_symbolic_fill(w, (int*)h->col_idx.p, (int*)h->tpos.p, s);
    w.s_len = (size_t)nnzb * 36;
    // PCG workspace: r, z, p0, p1, Ap (n each), Minv (36 n_cam), accumulators
    w.pcg_max_iter = h->prm.pcg_max_iter > 0 ? h->prm.pcg_max_iter : 200;
    w.pcg_tol = h->prm.pcg_tol > 0 ? h->prm.pcg_tol : 1e-10;
    const size_t nacc = 3 * (size_t)(w.pcg_max_iter + 2);
    if ((rc = ensure(h, h->pcg, sizeof(double) * (5 * (size_t)n + 36 * (size_t)n_cam + nacc + 8)))) return rc;
    double* q = (double*)h->pcg.p;
    w.pcg_r = q; q += n;
    w.pcg_z = q; q += n;
    w.pcg_p[0] = q; q += n;
    w.pcg_p[1] = q; q += n;
    w.pcg_Ap = q; q += n;
    w.pcg_Minv = q; q += 36 * (size_t)n_cam;
    w.pcg_acc = q;
  }
  {
    const size_t need = sizeof(double) * (w.s_len + 2 * (size_t)n + 1);
    // peers have this buffer mapped: they let go before it is reallocated (same call on every rank)
    if (h->peer.open && !(need <= h->red.cap && h->red.p) && (rc = peer_close(h, s))) return rc;
    if ((rc = ensure(h, h->red, need))) return rc;
  }
  w.S = (double*)h->red.p;
  w.b = w.S + w.s_len;
  w.diagU = w.b + n;
  w.cost_lin = w.diagU + n;
  return BA_OK;
}

static bool panel_reorder_enabled() {           // diagnostics: BA_PANEL_REORDER=0 keeps the caller's point order
  static int v = -1;
  if (v < 0) { const char* e = getenv("BA_PANEL_REORDER"); v = (e && !atoi(e)) ? 0 : 1; }
  return v != 0;
}

static int pick_schur(const ba_handle* h, const Work& w, bool ignore_alignment = false) {
  int kind = h->prm.schur_kind;
  if (kind == BA_SCHUR_AUTO) {
    // lane-per-keyframe staging needs the cameras of consecutive points inside a 20-keyframe window:
    // whole local windows, or sequential (banded) co-visibility
    const bool seq = w.band_w >= 0 && w.band_w < mma_window();
    // whole windows (<= 20 keyframes): the round-1 kernel (its lane-per-keyframe staging has no set-up passes; measured
    // C1 25 vs 39 us, C2 55 vs 75 us per linearisation); longer banded windows: the panel path (C4 3.7 -> 2.1 ms)
    kind = (w.n_cam <= mma_window()) ? BA_SCHUR_MMA : (seq ? BA_SCHUR_PANEL : BA_SCHUR_STREAM);
  }
  // the panel path needs its workspace and 16-byte aligned observation arrays (bulk copies); the round-1 kernel takes anything
  if (kind == BA_SCHUR_PANEL && !ignore_alignment && !panel_supported(w)) kind = BA_SCHUR_MMA;
  return kind;
}

static int launch_linearize(const Work& w, int schur, int sm_count, cudaStream_t s, cudaEvent_t* marks = nullptr) {
  if (schur == BA_SCHUR_PANEL) return launch_linearize_panel(w, sm_count, s, marks);
  if (schur == BA_SCHUR_MMA) return launch_linearize_mma(w, sm_count, s);
  if (schur == BA_SCHUR_STREAM) return launch_linearize_stream(w, sm_count, s);
  if (schur == BA_SCHUR_TILED) return launch_linearize_tiled(w, sm_count, s);
  return launch_linearize_scatter(w, sm_count, s);
}

// enqueue the LM iterations (no init / finalize); returns kernels launched or < 0
static int enqueue_lm_loop(ba_handle* h, const Work& w, int schur, cudaStream_t s) {
  int launches = 0;
  const size_t nred = w.s_len + 2 * (size_t)w.n + 1;
  // sharded banded path: the solver's first kernel sums the ranks' reduction buffers over peer memory itself
  const bool peers = h->world > 1 && h->peer.open && w.band_w >= 0 && w.bcr_ws != nullptr;
  const PeerView pv = peers ? peer_view(h) : PeerView();
  const bool sparse = w.bsr_bits != nullptr;
  const int pcg_grid = sparse ? pcg_grid_size(w.n_cam, h->sm_count) : 0;
  for (int it = 0; it < h->prm.max_iter; ++it) {
    {
      // profile mode, panel path: the three kernels of a linearisation are timed separately
      cudaEvent_t marks[2];
      const bool split = h->prm.profile && schur == BA_SCHUR_PANEL;
      prof_mark(h, split ? TAG_LIN_PB : TAG_LIN, s);
      if (split) { marks[0] = prof_slot(h, TAG_LIN_CAM); marks[1] = prof_slot(h, TAG_LIN); }
      const int nl = launch_linearize(w, schur, h->sm_count, s, split ? marks : nullptr);
      if (nl < 0) { h->err = std::string("linearise launch: ") + cudaGetErrorString(cudaGetLastError()); return BA_ERR_CUDA; }
      launches += nl;
    }
    if (h->world > 1 && !peers) {
      prof_mark(h, TAG_COMM, s);
      int r = g_nccl.allreduce(w.S, w.S, nred, kNcclDouble, kNcclSum, h->comm, s);
      if (r != 0) { h->err = std::string("ncclAllReduce(S): ") + (g_nccl.errstr ? g_nccl.errstr(r) : "?"); return BA_ERR_NCCL; }
    }
    prof_mark(h, TAG_SOLVE, s);
    if (w.band_w >= 0) {
      launches += w.bcr_ws ? launch_rcs_bcr(w, w.bcr_ws, s, h->sm_count, peers ? &pv : nullptr, it) : launch_rcs_band(w, s);
    } else if (sparse) {
      if (launch_rcs_pcg(w, pcg_grid, s) < 0) { h->err = std::string("cooperative PCG launch: ") + cudaGetErrorString(cudaGetLastError()); return BA_ERR_CUDA; }
      launches += 1;
    } else {
      launches += launch_rcs_cholesky(w, s);
    }
    static int no_bcast = -1;       // diagnostics only (BA_NO_BCAST=1): how much of the iteration is the broadcast
    if (no_bcast < 0) { const char* e = getenv("BA_NO_BCAST"); no_bcast = (e && atoi(e)) ? 1 : 0; }
    // (the cyclic-reduction solver is deterministic -- gathered updates, no REDs -- and its inputs are identical bits on
    // every rank: no broadcast there)
    if (h->world > 1 && !no_bcast && !(w.band_w >= 0 && w.bcr_ws != nullptr)) {
      // the replicated solve ends in float64 REDs whose order differs from rank to rank: rank 0's camera step
      // and candidate poses go to everyone, so the replicated state stays bit-identical across ranks
      prof_mark(h, TAG_COMM, s);
      int r = g_nccl.broadcast(w.pose[0], w.pose[0], 30 * (size_t)w.n_cam, kNcclDouble, 0, h->comm, s);
      if (r != 0) { h->err = std::string("ncclBroadcast(camera state): ") + (g_nccl.errstr ? g_nccl.errstr(r) : "?"); return BA_ERR_NCCL; }
    }
    prof_mark(h, TAG_BACK, s);
    launches += launch_backsub(w, h->sm_count, s);
    if (h->world > 1) {
      prof_mark(h, TAG_COMM, s);
      int r = g_nccl.allreduce(w.cost_cand, w.cost_cand, 2, kNcclDouble, kNcclSum, h->comm, s);
      if (r != 0) { h->err = std::string("ncclAllReduce(cost): ") + (g_nccl.errstr ? g_nccl.errstr(r) : "?"); return BA_ERR_NCCL; }
      launches += launch_accept(w, s);
      // every rank has passed its assembly by now (it fed the all-reduce after it): nobody reads this buffer any more
      if (peers) CK(cudaMemsetAsync(w.S, 0, sizeof(double) * nred, s));
    }
  }
  return launches;
}

static void dump_dbg(ba_handle* h) {
  if (!h->dbg.p) return;
  std::vector<long long> hb(4096);
  cudaMemcpy(hb.data(), h->dbg.p, sizeof(long long) * 4096, cudaMemcpyDeviceToHost);
  fprintf(stderr, "[ba dbg]");
  std::vector<std::pair<long long, long long>> ev;      // (clock, tag) in time order
  for (int i = 0; i < 2000; ++i)
    if (hb[2 + 2 * i] != 0) ev.push_back({hb[2 + 2 * i], hb[1 + 2 * i]});
  std::sort(ev.begin(), ev.end());
  long long prev = 0;
  for (auto& e : ev) {
    fprintf(stderr, " %lld:%lld", e.second, prev ? e.first - prev : 0LL);
    prev = e.first;
  }
  fprintf(stderr, "\n");
}

static void fill_report(const ba_handle* h, const LmCtl* c, ba_report* rep, int status, float ms,
                        int launches) {
  memset(rep, 0, sizeof(*rep));
  rep->status = status;
  rep->n_iter = c->iter;
  rep->n_accepted = c->n_accepted;
  rep->solver_used = h->last_solver;
  rep->band_width = h->last_band_w;
  rep->s_blocks = h->last_s_blocks;
  for (int i = 0; i <= BA_MAX_ITER; ++i) rep->cost_hist[i] = c->cost_hist[i];
  for (int i = 0; i < BA_MAX_ITER; ++i) {
    rep->cand_cost_hist[i] = c->cand_hist[i];
    rep->lambda_hist[i] = c->lambda_hist[i];
    rep->accepted[i] = c->accepted[i];
    rep->solve_fail[i] = c->solve_fail_hist[i];
    rep->pcg_iters[i] = c->pcg_iters_hist[i];
  }
  rep->gpu_ms = ms;
  rep->n_launches = launches;
  rep->ms_linearize = h->prof_ms[TAG_LIN] + h->prof_ms[TAG_LIN_PB] + h->prof_ms[TAG_LIN_CAM];
  rep->ms_point_blocks = h->prof_ms[TAG_LIN_PB];
  rep->ms_cam_blocks = h->prof_ms[TAG_LIN_CAM];
  rep->peer_exchange = (h->world > 1 && h->peer.open && h->last_solver == BA_SOLVER_BAND) ? 1 : 0;
  rep->ms_solve = h->prof_ms[TAG_SOLVE];
  rep->ms_backsub = h->prof_ms[TAG_BACK];
  rep->ms_other = h->prof_ms[TAG_OTHER];
  rep->ms_comm = h->prof_ms[TAG_COMM];
}

extern "C" int ba_solve_local(ba_handle* h, int32_t n_cam, int32_t n_pt, int32_t n_obs,
                              const double intr[4], double* cam_R, double* cam_t,
                              const uint8_t* cam_fixed, double* X, const int32_t* obs_cam,
                              const int32_t* obs_pt, const float* obs_uv, const int32_t* pt_offsets,
                              uint8_t* flags, void* stream, ba_report* report) {
  if (!h) return BA_ERR_INVALID;
  if (n_cam <= 0 || n_pt < 0 || n_obs < 0 || !intr || !cam_R || !cam_t || !cam_fixed ||
      (n_pt > 0 && (!X || !pt_offsets)) || (n_obs > 0 && (!obs_cam || !obs_pt || !obs_uv)))
    return BA_ERR_INVALID;
  DeviceGuard dev_guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  int mode = MODE_DENSE, bw = -1;
  int rc = pick_mode(h, n_cam, n_pt, obs_cam, pt_offsets, s, &mode, &bw);
  if (rc) return rc;
  const bool sparse = mode != MODE_DENSE;
  Work w;
  rc = make_work(h, n_cam, n_pt, n_obs, intr, cam_fixed, obs_cam, obs_pt, obs_uv, pt_offsets, w, mode, s, bw);
  if (rc) return rc;
  h->last_solver = mode == MODE_DENSE ? BA_SOLVER_CHOLESKY : (mode == MODE_BAND ? BA_SOLVER_BAND : BA_SOLVER_PCG);
  h->last_band_w = bw;
  h->last_s_blocks = (int)(w.s_len / 36);
  if (h->world > 1) {
    // peer-memory exchange for the cyclic-reduction band solver; (re)mapped whenever the reduction buffer moved
    const bool want = peer_wanted(h) && mode == MODE_BAND && w.bcr_ws != nullptr;
    if (h->peer.open && (!want || h->peer.red_exported != h->red.p) && (rc = peer_close(h, s))) return rc;
    if (want && !h->peer.open && (rc = peer_open(h, s))) return rc;
    h->peer.solves++;
  }
  int schur = pick_schur(h, w, /*ignore_alignment=*/true);
  LmInit li{h->prm.lambda0, h->prm.huber_delta, h->prm.chi2_thr, h->prm.lambda_up,
            h->prm.lambda_down, h->prm.lambda_min, h->prm.lambda_max, h->prm.accept_rtol, 0};
  li.epoch_base = (long long)(64ull * h->peer.solves);

  int launches = 0;
  h->prof_n = 0;
  if (w.dbg) cudaMemsetAsync(w.dbg, 0, sizeof(long long) * 4096, s);
  CK(cudaEventRecord(h->ev0, s));
  prof_mark(h, TAG_OTHER, s);
  {
    // the shard's camera range: everything, unless the class-order pass below measures it
    if ((rc = ensure(h, h->ro_range, 64))) return rc;
    const int full_range[2] = {0, n_cam - 1};
    CK(cudaMemcpyAsync(h->ro_range.p, full_range, sizeof(full_range), cudaMemcpyHostToDevice, s));
  }
  if (schur == BA_SCHUR_PANEL && h->prm.max_iter > 0 && n_pt > 0 && n_obs > 0 && panel_reorder_enabled()) {
    // the panel path works in its own point order (class order, ba_panel.cu): a permuted copy of the observations
    // and offsets, built on the device once per solve; k_init / k_finalize translate points and flags
    if ((rc = ensure(h, h->ro_cnt, sizeof(unsigned long long) * (panel_classes(n_cam) + 2)))) return rc;
    if ((rc = ensure(h, h->ro_perm, sizeof(int) * (size_t)(n_pt + 1)))) return rc;
    if ((rc = ensure(h, h->ro_off, sizeof(int) * (size_t)(n_pt + 1)))) return rc;
    if ((rc = ensure(h, h->ro_cam, sizeof(int) * (size_t)(n_obs + 4)))) return rc;
    if ((rc = ensure(h, h->ro_pt, sizeof(int) * (size_t)(n_obs + 4)))) return rc;
    if ((rc = ensure(h, h->ro_uv, sizeof(float) * 2 * (size_t)(n_obs + 4)))) return rc;
    PanelOrder po{(int*)h->ro_range.p, (unsigned long long*)h->ro_cnt.p, (int*)h->ro_perm.p, (int*)h->ro_off.p,
                  (int*)h->ro_cam.p, (int*)h->ro_pt.p, (float*)h->ro_uv.p};
    launches += launch_panel_reorder(w, po, s);
    w.perm = po.perm;
    w.pt_off_in = pt_offsets;
    w.obs_cam = po.cam; w.obs_pt = po.pt; w.obs_uv = po.uv; w.pt_off = po.off;
  }
  if (schur == BA_SCHUR_PANEL && !panel_supported(w)) schur = BA_SCHUR_MMA;
  launches += launch_init(w, cam_R, cam_t, X, li, s);
  if (schur == BA_SCHUR_PANEL && h->prm.max_iter > 0) launches += launch_panel_setup(w, h->sm_count, s);
  if (h->prm.max_iter == 0) {
    // nothing will linearise: the report's initial cost comes from a plain evaluation of the objective
    launches += launch_eval_cost(w, w.pose[0], w.X[0], w.cost_cand, h->prm.huber_delta, s);
    if (h->world > 1) {
      int r = g_nccl.allreduce(w.cost_cand, w.cost_cand, 1, kNcclDouble, kNcclSum, h->comm, s);
      if (r != 0) { h->err = std::string("ncclAllReduce(cost): ") + (g_nccl.errstr ? g_nccl.errstr(r) : "?"); return BA_ERR_NCCL; }
    }
    launches += launch_cost0(w, s);
  }

  bool use_graph = h->prm.use_graph != 0 && h->prm.max_iter > 0 && !h->prm.profile && !sparse;
  if (use_graph) {
    ba_handle::GraphKey key;
    key.n_cam = n_cam; key.n_pt = n_pt; key.n_obs = n_obs; key.max_iter = h->prm.max_iter;
    key.schur = schur; key.world = 2 * h->world + (h->peer.open ? 1 : 0);
    const void* ptrs[10] = {cam_fixed, w.obs_cam, w.obs_pt, w.obs_uv, w.pt_off, w.S, w.pose[0], w.X[0], w.ctl, w.dp};
    memcpy(key.ptrs, ptrs, sizeof(ptrs));
    key.scal[0] = intr[0]; key.scal[1] = intr[1]; key.scal[2] = intr[2]; key.scal[3] = intr[3];
    if (!(h->graph && key == h->graph_key)) {
      if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
      cudaGraph_t g = nullptr;
      cudaStream_t cs = h->stream;
      CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      int n = enqueue_lm_loop(h, w, schur, cs);
      cudaError_t ce = cudaStreamEndCapture(cs, &g);
      if (n < 0) { if (g) cudaGraphDestroy(g); return n; }
      if (ce != cudaSuccess) { h->err = std::string("graph capture: ") + cudaGetErrorString(ce); return BA_ERR_CUDA; }
      ce = cudaGraphInstantiate(&h->graph, g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) { h->graph = nullptr; h->err = std::string("graph instantiate: ") + cudaGetErrorString(ce); return BA_ERR_CUDA; }
      h->graph_key = key;
      h->graph_launches = n;
    }
    CK(cudaGraphLaunch(h->graph, s));
    launches += h->graph_launches;
  } else {
    int n = enqueue_lm_loop(h, w, schur, s);
    if (n < 0) return n;
    launches += n;
  }
  prof_mark(h, TAG_OTHER, s);
  launches += launch_finalize(w, cam_R, cam_t, X, flags, s);
  prof_mark(h, TAG_OTHER, s);
  CK(cudaEventRecord(h->ev1, s));
  CK(cudaGetLastError());
  if (report) {
    CK(cudaMemcpyAsync(h->h_ctl, w.ctl, sizeof(LmCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    if (h->prm.profile) prof_collect(h);
    fill_report(h, h->h_ctl, report, BA_OK, ms, launches);
    dump_dbg(h);
  }
  return BA_OK;
}

extern "C" int ba_solve_local_host(ba_handle* h, int32_t n_cam, int32_t n_pt, int32_t n_obs,
                                   const double intr[4], double* cam_R, double* cam_t,
                                   const uint8_t* cam_fixed, double* X, const int32_t* obs_cam,
                                   const int32_t* obs_pt, const float* obs_uv,
                                   const int32_t* pt_offsets, uint8_t* flags, ba_report* report) {
  (void)obs_pt;                    // expanded from pt_offsets on the device
  if (!h) return BA_ERR_INVALID;
  if (n_cam <= 0 || n_pt < 0 || n_obs < 0 || !intr || !cam_R || !cam_t || !cam_fixed ||
      (n_pt > 0 && (!X || !pt_offsets)) || (n_obs > 0 && (!obs_cam || !obs_uv)))
    return BA_ERR_INVALID;
  // the kernels index with the offsets and camera indices unchecked: the ends are looked at here, everything in between
  // by a kernel once the arrays are on the device (k_validate; the host loops cost 2-3 ms of a C4 call)
  if (n_pt > 0) {
    if (pt_offsets[0] != 0 || pt_offsets[n_pt] != n_obs) { h->err = "pt_offsets must start at 0 and end at n_obs"; return BA_ERR_INVALID; }
  } else if (n_obs != 0) {
    h->err = "observations without points";
    return BA_ERR_INVALID;
  }
  DeviceGuard dev_guard(h->device);
  cudaStream_t s = h->stream;
  int rc;
  // R | t | fixed in one pinned staging record: one copy up, one copy (R | t) down
  const size_t nR = 9 * (size_t)n_cam, nT = 3 * (size_t)n_cam;
  const size_t small_bytes = sizeof(double) * (nR + nT) + (size_t)n_cam;
  if (small_bytes > h->h_small_cap) {
    if (h->h_small) { cudaFreeHost(h->h_small); h->h_small = nullptr; h->h_small_cap = 0; }
    const size_t cap = small_bytes + small_bytes / 4 + 256;
    CK(cudaMallocHost((void**)&h->h_small, cap));
    h->h_small_cap = cap;
  }
  if ((rc = ensure(h, h->d_small, small_bytes))) return rc;
  if ((rc = ensure(h, h->d_X, sizeof(double) * 3 * (size_t)(n_pt + 1)))) return rc;
  if ((rc = ensure(h, h->d_ocam, sizeof(int32_t) * (size_t)(n_obs + 1)))) return rc;
  if ((rc = ensure(h, h->d_opt, sizeof(int32_t) * (size_t)(n_obs + 1)))) return rc;
  if ((rc = ensure(h, h->d_ouv, sizeof(float) * 2 * (size_t)(n_obs + 1)))) return rc;
  if ((rc = ensure(h, h->d_off, sizeof(int32_t) * (size_t)(n_pt + 1)))) return rc;
  if ((rc = ensure(h, h->d_flags, (size_t)(n_obs + 1)))) return rc;
  memcpy(h->h_small, cam_R, sizeof(double) * nR);
  memcpy(h->h_small + nR, cam_t, sizeof(double) * nT);
  memcpy(h->h_small + nR + nT, cam_fixed, (size_t)n_cam);
  double* d_R = (double*)h->d_small.p;
  double* d_t = d_R + nR;
  const uint8_t* d_fixed = (const uint8_t*)(d_t + nT);
  CK(cudaMemcpyAsync(h->d_small.p, h->h_small, small_bytes, cudaMemcpyHostToDevice, s));
  // indices first, their check behind them; the host looks at its verdict while the big arrays are still on their way
  if ((rc = ensure(h, h->d_bad, 64))) return rc;
  if (n_pt > 0) CK(cudaMemcpyAsync(h->d_off.p, pt_offsets, sizeof(int32_t) * (size_t)(n_pt + 1), cudaMemcpyHostToDevice, s));
  if (n_obs > 0) CK(cudaMemcpyAsync(h->d_ocam.p, obs_cam, sizeof(int32_t) * (size_t)n_obs, cudaMemcpyHostToDevice, s));
  launch_validate(n_cam, n_pt, n_obs, (const int*)h->d_ocam.p, (const int*)h->d_off.p, (int*)h->d_bad.p, s);
  CK(cudaMemcpyAsync(h->h_nnzb, h->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(h->ev_val, s));
  if (n_pt > 0) CK(cudaMemcpyAsync(h->d_X.p, X, sizeof(double) * 3 * (size_t)n_pt, cudaMemcpyHostToDevice, s));
  if (n_obs > 0) CK(cudaMemcpyAsync(h->d_ouv.p, obs_uv, sizeof(float) * 2 * (size_t)n_obs, cudaMemcpyHostToDevice, s));
  CK(cudaEventSynchronize(h->ev_val));
  if (*h->h_nnzb != 0) {
    cudaStreamSynchronize(s);
    h->err = (*h->h_nnzb & 1) ? "pt_offsets must be non-decreasing" : "obs_cam out of range";
    return BA_ERR_INVALID;
  }
  if (n_pt > 0 && n_obs > 0) launch_expand_obs_pt((const int*)h->d_off.p, n_pt, (int*)h->d_opt.p, s);
  rc = ba_solve_local(h, n_cam, n_pt, n_obs, intr, d_R, d_t, d_fixed, (double*)h->d_X.p,
                      (const int32_t*)h->d_ocam.p, (const int32_t*)h->d_opt.p, (const float*)h->d_ouv.p,
                      (const int32_t*)h->d_off.p, flags ? (uint8_t*)h->d_flags.p : nullptr, s, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_small, h->d_small.p, sizeof(double) * (nR + nT), cudaMemcpyDeviceToHost, s));
  if (n_pt > 0) CK(cudaMemcpyAsync(X, h->d_X.p, sizeof(double) * 3 * (size_t)n_pt, cudaMemcpyDeviceToHost, s));
  if (flags && n_obs > 0) CK(cudaMemcpyAsync(flags, h->d_flags.p, (size_t)n_obs, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h->h_ctl, h->ctl.p, sizeof(LmCtl), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  memcpy(cam_R, h->h_small, sizeof(double) * nR);
  memcpy(cam_t, h->h_small + nR, sizeof(double) * nT);
  dump_dbg(h);
  if (report) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    int launches = 2 + (h->graph_launches > 0 ? h->graph_launches : h->prm.max_iter * (h->world > 1 ? 4 : 3));
    if (h->prm.profile) prof_collect(h);
    fill_report(h, h->h_ctl, report, BA_OK, ms, launches);
  }
  return BA_OK;
}

extern "C" int ba_eval_cost(ba_handle* h, int32_t n_cam, int32_t n_pt, int32_t n_obs,
                            const double intr[4], const double* cam_R, const double* cam_t,
                            const double* X, const int32_t* obs_cam, const int32_t* obs_pt,
                            const float* obs_uv, void* stream, double* cost_out) {
  if (!h || !cost_out || n_cam <= 0) return BA_ERR_INVALID;
  DeviceGuard dev_guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  Work w;
  int rc = make_work(h, n_cam, n_pt, n_obs, intr, nullptr, obs_cam, obs_pt, obs_uv, nullptr, w, MODE_NONE);
  if (rc) return rc;
  LmInit li{h->prm.lambda0, h->prm.huber_delta, h->prm.chi2_thr, h->prm.lambda_up,
            h->prm.lambda_down, h->prm.lambda_min, h->prm.lambda_max, h->prm.accept_rtol};
  launch_init(w, cam_R, cam_t, X, li, s);
  launch_eval_cost(w, w.pose[0], w.X[0], w.cost_cand, h->prm.huber_delta, s);
  CK(cudaMemcpyAsync(cost_out, w.cost_cand, sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaMemsetAsync(w.cost_cand, 0, sizeof(double), s));
  return BA_OK;
}

extern "C" int ba_debug_linearize(ba_handle* h, int32_t n_cam, int32_t n_pt, int32_t n_obs,
                                  const double intr[4], const double* cam_R, const double* cam_t,
                                  const uint8_t* cam_fixed, const double* X, const int32_t* obs_cam,
                                  const int32_t* obs_pt, const float* obs_uv,
                                  const int32_t* pt_offsets, double lambda, void* stream,
                                  double* S_host, double* b_host, double* diagU_host,
                                  double* cost_host) {
  if (!h || n_cam <= 0) return BA_ERR_INVALID;
  DeviceGuard dev_guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (6 * n_cam > rcs_cholesky_max_n()) { h->err = "ba_debug_linearize: dense layout only"; return BA_ERR_TOO_LARGE; }
  Work w;
  int rc = make_work(h, n_cam, n_pt, n_obs, intr, cam_fixed, obs_cam, obs_pt, obs_uv, pt_offsets, w);
  if (rc) return rc;
  LmInit li{lambda, h->prm.huber_delta, h->prm.chi2_thr, h->prm.lambda_up,
            h->prm.lambda_down, h->prm.lambda_min, h->prm.lambda_max, h->prm.accept_rtol};
  launch_init(w, cam_R, cam_t, X, li, s);
  const int dbg_schur = pick_schur(h, w);
  if (dbg_schur == BA_SCHUR_PANEL) launch_panel_setup(w, h->sm_count, s);
  launch_linearize(w, dbg_schur, h->sm_count, s);
  const size_t n = (size_t)w.n;
  if (S_host) CK(cudaMemcpyAsync(S_host, w.S, sizeof(double) * s_doubles(w.n_cam), cudaMemcpyDeviceToHost, s));
  if (b_host) CK(cudaMemcpyAsync(b_host, w.b, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (diagU_host) CK(cudaMemcpyAsync(diagU_host, w.diagU, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  if (cost_host) CK(cudaMemcpyAsync(cost_host, w.cost_lin, sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return BA_OK;
}

// ------------------------------------------------------------------------------------------
// pose-only
// ------------------------------------------------------------------------------------------
extern "C" int ba_solve_pose_only(ba_handle* h, int32_t n_frames, int32_t n_match,
                                  const double intr[4], double* R, double* t, const float* X,
                                  const float* uv, const int32_t* frame_offsets, uint8_t* flags,
                                  double* cost_hist, uint8_t* accepted, void* stream) {
  if (!h || n_frames < 0 || n_match < 0 || !intr) return BA_ERR_INVALID;
  if (n_frames == 0) return BA_OK;
  if (!R || !t || !frame_offsets || (n_match > 0 && (!X || !uv))) return BA_ERR_INVALID;
  DeviceGuard dev_guard(h->device);
  PoseWork w;
  w.n_frames = n_frames;
  w.K = Intr{intr[0], intr[1], intr[2], intr[3]};
  w.R = R; w.t = t; w.X = X; w.uv = uv; w.frame_off = frame_offsets;
  w.flags = flags; w.cost_hist = cost_hist; w.accepted = accepted;
  w.max_iter = h->prm.max_iter;
  w.huber_delta = h->prm.huber_delta; w.chi2_thr = h->prm.chi2_thr; w.lambda0 = h->prm.lambda0;
  w.lambda_up = h->prm.lambda_up; w.lambda_down = h->prm.lambda_down;
  w.lambda_min = h->prm.lambda_min; w.lambda_max = h->prm.lambda_max;
  w.accept_rtol = h->prm.accept_rtol;
  launch_pose_only(w, (cudaStream_t)stream);
  CK(cudaGetLastError());
  return BA_OK;
}

extern "C" int ba_solve_pose_only_host(ba_handle* h, int32_t n_frames, int32_t n_match,
                                       const double intr[4], double* R, double* t, const float* X,
                                       const float* uv, const int32_t* frame_offsets,
                                       uint8_t* flags, double* cost_hist, uint8_t* accepted,
                                       float* gpu_ms) {
  if (!h || n_frames < 0 || n_match < 0 || !intr) return BA_ERR_INVALID;
  if (n_frames == 0) return BA_OK;
  if (!R || !t || !frame_offsets || (n_match > 0 && (!X || !uv))) return BA_ERR_INVALID;
  DeviceGuard dev_guard(h->device);
  cudaStream_t s = h->stream;
  const int mi = h->prm.max_iter;
  int rc;
  if ((rc = ensure(h, h->d_R, sizeof(double) * 9 * (size_t)n_frames))) return rc;
  if ((rc = ensure(h, h->d_t, sizeof(double) * 3 * (size_t)n_frames))) return rc;
  if ((rc = ensure(h, h->d_X, sizeof(float) * 3 * (size_t)(n_match + 1)))) return rc;
  if ((rc = ensure(h, h->d_ouv, sizeof(float) * 2 * (size_t)(n_match + 1)))) return rc;
  if ((rc = ensure(h, h->d_foff, sizeof(int32_t) * (size_t)(n_frames + 1)))) return rc;
  if ((rc = ensure(h, h->d_flags, (size_t)(n_match + 1)))) return rc;
  if ((rc = ensure(h, h->d_hist, sizeof(double) * (size_t)n_frames * (mi + 1)))) return rc;
  if ((rc = ensure(h, h->d_acc, (size_t)n_frames * (mi + 1)))) return rc;
  CK(cudaMemcpyAsync(h->d_R.p, R, sizeof(double) * 9 * (size_t)n_frames, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->d_t.p, t, sizeof(double) * 3 * (size_t)n_frames, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->d_foff.p, frame_offsets, sizeof(int32_t) * (size_t)(n_frames + 1), cudaMemcpyHostToDevice, s));
  if (n_match > 0) {
    CK(cudaMemcpyAsync(h->d_X.p, X, sizeof(float) * 3 * (size_t)n_match, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->d_ouv.p, uv, sizeof(float) * 2 * (size_t)n_match, cudaMemcpyHostToDevice, s));
  }
  CK(cudaEventRecord(h->ev0, s));
  rc = ba_solve_pose_only(h, n_frames, n_match, intr, (double*)h->d_R.p, (double*)h->d_t.p,
                          (const float*)h->d_X.p, (const float*)h->d_ouv.p,
                          (const int32_t*)h->d_foff.p, flags ? (uint8_t*)h->d_flags.p : nullptr,
                          cost_hist ? (double*)h->d_hist.p : nullptr,
                          accepted ? (uint8_t*)h->d_acc.p : nullptr, s);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev1, s));
  CK(cudaMemcpyAsync(R, h->d_R.p, sizeof(double) * 9 * (size_t)n_frames, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(t, h->d_t.p, sizeof(double) * 3 * (size_t)n_frames, cudaMemcpyDeviceToHost, s));
  if (flags && n_match > 0) CK(cudaMemcpyAsync(flags, h->d_flags.p, (size_t)n_match, cudaMemcpyDeviceToHost, s));
  if (cost_hist) CK(cudaMemcpyAsync(cost_hist, h->d_hist.p, sizeof(double) * (size_t)n_frames * (mi + 1), cudaMemcpyDeviceToHost, s));
  if (accepted && mi > 0) CK(cudaMemcpyAsync(accepted, h->d_acc.p, (size_t)n_frames * mi, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (gpu_ms) cudaEventElapsedTime(gpu_ms, h->ev0, h->ev1);
  return BA_OK;
}

extern "C" int ba_pose_ransac_host(ba_handle* h, int32_t n_match, const double intr[4], const float* X, const float* uv,
                                   int32_t n_hyp, double thr_px, uint32_t seed, double* R_out, double* t_out,
                                   int32_t* n_inliers) {
  if (!h || !intr || !R_out || !t_out || !n_inliers || n_hyp <= 0 || n_hyp > 65535 || !(thr_px > 0)) return BA_ERR_INVALID;
  *n_inliers = 0;
  if (n_match < 4 || !X || !uv) return BA_ERR_INVALID;
  DeviceGuard dev_guard(h->device);
  cudaStream_t s = h->stream;
  int rc;
  const int od = ransac_out_doubles();
  if ((rc = ensure(h, h->d_X, sizeof(float) * 3 * (size_t)(n_match + 1)))) return rc;
  if ((rc = ensure(h, h->d_ouv, sizeof(float) * 2 * (size_t)(n_match + 1)))) return rc;
  if ((rc = ensure(h, h->d_hist, sizeof(double) * (size_t)n_hyp * od))) return rc;
  CK(cudaMemcpyAsync(h->d_X.p, X, sizeof(float) * 3 * (size_t)n_match, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(h->d_ouv.p, uv, sizeof(float) * 2 * (size_t)n_match, cudaMemcpyHostToDevice, s));
  RansacWork w;
  w.n_match = n_match; w.n_hyp = n_hyp;
  w.K = Intr{intr[0], intr[1], intr[2], intr[3]};
  w.X = (const float*)h->d_X.p; w.uv = (const float*)h->d_ouv.p;
  w.thr2 = thr_px * thr_px; w.seed = seed; w.out = (double*)h->d_hist.p;
  launch_pose_ransac(w, s);
  CK(cudaGetLastError());
  std::vector<double> out((size_t)n_hyp * od);
  CK(cudaMemcpyAsync(out.data(), h->d_hist.p, sizeof(double) * out.size(), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  int best = -1;
  double best_cnt = 0.0;
  for (int i = 0; i < n_hyp; ++i)
    if (out[(size_t)i * od] > best_cnt) { best_cnt = out[(size_t)i * od]; best = i; }
  if (best < 0) return BA_OK;                       // no admissible hypothesis: n_inliers stays 0
  for (int q = 0; q < 9; ++q) R_out[q] = out[(size_t)best * od + 1 + q];
  for (int q = 0; q < 3; ++q) t_out[q] = out[(size_t)best * od + 10 + q];
  *n_inliers = (int32_t)best_cnt;
  return BA_OK;
}

// ------------------------------------------------------------------------------------------
// multi-GPU plumbing
// ------------------------------------------------------------------------------------------
extern "C" int ba_nccl_unique_id(void* out_id) {
  if (!out_id) return BA_ERR_INVALID;
  if (!load_nccl()) return BA_ERR_NCCL;
  nccl_uid id;
  if (g_nccl.get_uid(&id) != 0) return BA_ERR_NCCL;
  memcpy(out_id, &id, sizeof(id));
  return BA_OK;
}

extern "C" int ba_dist_init(ba_handle* h, const void* id, int32_t rank, int32_t world) {
  if (!h || !id || world < 1 || rank < 0 || rank >= world) return BA_ERR_INVALID;
  if (world == 1) { h->rank = 0; h->world = 1; return BA_OK; }
  if (!load_nccl()) { h->err = "libnccl.so.2 not found"; return BA_ERR_NCCL; }
  DeviceGuard dev_guard(h->device);
  nccl_uid uid;
  memcpy(&uid, id, sizeof(uid));
  int r = g_nccl.init_rank(&h->comm, world, uid, rank);
  if (r != 0) { h->err = std::string("ncclCommInitRank: ") + (g_nccl.errstr ? g_nccl.errstr(r) : "?"); return BA_ERR_NCCL; }
  h->rank = rank;
  h->world = world;
  if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
  return BA_OK;
}

extern "C" int ba_dist_shutdown(ba_handle* h) {
  if (!h) return BA_ERR_INVALID;
  peer_unmap(h);
  if (h->comm && g_nccl.destroy) { g_nccl.destroy(h->comm); h->comm = nullptr; }
  h->world = 1; h->rank = 0;
  if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
  return BA_OK;
}
