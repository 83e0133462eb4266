/*
 * b200ba.h -- C ABI of the B200-native local bundle-adjustment solver.
 *
 * This is the drop-in boundary.  The reference
 * (AntonSHBK/visual-slam-core) is pure Python and has no FFI of its own; the interface
 * these entry points replace is the body of its optimizer plugin:
 *
 *   ba_solve_local[_host]      <- TorchOptimizer.optimize_local, the Adam/autograd hot loop
 *                                 visual_slam/optimization/torch_optimizer.py:103-184
 *                                 (inputs = the tensors it packs at :43-95, outputs = what it
 *                                 writes back at :188-203, plus per-observation outlier flags)
 *   ba_solve_pose_only[_host]  <- estimate_motion_from_2d3d (cv2.solvePnPRansac refinement)
 *                                 visual_slam/utils/motion_estimation.py:57-132,
 *                                 called from visual_slam/tracking.py:350-356
 *   ba_pose_ransac_host        <- the RANSAC consensus inside cv2.solvePnPRansac at the same call site (no prior passed)
 *   ba_dist_*                  <- (no reference counterpart; shard-by-point multi-GPU, 8e-E1)
 *
 * Conventions
 *   - every function returns an int status: BA_OK (0) or a negative BA_ERR_* code; nothing
 *     throws across the boundary.  ba_error_string() names a code; ba_last_error() returns the
 *     handle's last CUDA/NCCL message.
 *   - poses are world->camera, R row-major 3x3 double, t double[3]; points are double[3] (xyz);
 *     observations are sorted by point: obs of point j are [pt_offsets[j], pt_offsets[j+1]).
 *   - the *_host variants take host pointers and do the host<->device copies themselves
 *     (pinned staging owned by the handle); the plain variants take DEVICE pointers and
 *     enqueue on `stream` (a cudaStream_t passed as void*).
 *   - arithmetic and the optimised state (poses, points) are float64; observations are float32
 *     (cv2.KeyPoint.pt, torch_optimizer.py:94,174).  The packer rounds the initial poses and points
 *     through float32 as the reference does when it builds its tensors (:51-52,65-68).
 *   - a handle is not re-entrant: one solve at a time per handle (the Python plugin locks).
 */
#ifndef B200BA_H
#define B200BA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BA_MAX_ITER 64
#define BA_NCCL_ID_BYTES 128

enum {
  BA_OK = 0,
  BA_ERR_INVALID = -1,      /* bad argument (null pointer, negative size, max_iter > BA_MAX_ITER) */
  BA_ERR_CUDA = -2,         /* a CUDA runtime call failed; see ba_last_error() */
  BA_ERR_TOO_LARGE = -3,    /* problem does not fit the selected solver */
  BA_ERR_NCCL = -4,         /* NCCL could not be loaded / a collective failed */
  BA_ERR_NO_DEVICE = -5     /* no CUDA device: the solver has no CPU fallback */
};

/* solver_kind */
enum {
  BA_SOLVER_AUTO = 0,      /* dense Cholesky if the window fits shared memory, else banded, else PCG */
  BA_SOLVER_CHOLESKY = 1,  /* dense in-shared-memory blocked Cholesky (windows up to 38 keyframes) */
  BA_SOLVER_PCG = 2,       /* block-sparse S + block-Jacobi PCG (any co-visibility structure) */
  BA_SOLVER_BAND = 3,      /* block-banded S + direct solve (co-visibility bandwidth <= 32 keyframes): block cyclic
                              reduction over 15-keyframe super-blocks when the bandwidth is <= 15 and the window has
                              >= 60 keyframes, else the sequential banded Cholesky */
  BA_SOLVER_BAND_SEQ = 4   /* block-banded S + the sequential banded Cholesky (diagnostics, cross-check) */
};
/* schur_kind */
enum {
  BA_SCHUR_AUTO = 0,     /* BA_SCHUR_PANEL for windows of <= 20 keyframes and for banded (sequential) windows, else BA_SCHUR_STREAM */
  BA_SCHUR_SCATTER = 1,  /* warp per point, one float64 RED per entry (no ordering requirement) */
  BA_SCHUR_TILED = 2,    /* round-1 v2 kernel: thread-owned blocks over a slot table */
  BA_SCHUR_STREAM = 3,   /* compact block lists, Q lanes per block, rows sorted by camera */
  BA_SCHUR_MMA = 4,      /* round-1 kernel: lane-per-keyframe staging + FP64 tensor-core (mma m8n8k4) panel products */
  BA_SCHUR_PANEL = 5     /* round-2 default: point-block pass + camera-major pass + lane-per-observation staging fed by
                            bulk copies, FP64 tensor-core products restricted to each point's camera interval */
};

typedef struct ba_params {
  double huber_delta;   /* config.optimization.huber_delta, torch_optimizer.py:108 */
  double chi2_thr;      /* outlier flag: r_u^2 + r_v^2 > chi2_thr (5.991) */
  double lambda0;       /* initial LM damping (1e-4) */
  double lambda_up;     /* on reject  (10)  */
  double lambda_down;   /* on accept  (0.1) */
  double lambda_min;    /* 1e-6: keeps rank-deficient point blocks (single in-window observation) conditioned */
  double lambda_max;    /* 1e8  */
  double accept_rtol;   /* accept iff cand < cur * (1 - accept_rtol)  (1e-8) */
  int32_t max_iter;     /* LM iterations, rejected ones count (10) */
  int32_t solver_kind;  /* BA_SOLVER_* */
  int32_t schur_kind;   /* BA_SCHUR_*  */
  int32_t pcg_max_iter; /* block-Jacobi PCG cap */
  double pcg_tol;       /* relative residual */
  int32_t use_graph;    /* capture the LM loop in a CUDA graph (1) */
  int32_t profile;      /* 1: no graph, CUDA events around every kernel -> ba_report.ms_* (0) */
} ba_params;

typedef struct ba_report {
  int32_t status;                       /* BA_OK or error */
  int32_t n_iter;                       /* LM iterations executed (= max_iter) */
  int32_t n_accepted;
  int32_t solver_used;                  /* BA_SOLVER_CHOLESKY / BA_SOLVER_PCG / BA_SOLVER_BAND */
  double cost_hist[BA_MAX_ITER + 1];    /* [0] initial cost, [i+1] cost after iteration i */
  double cand_cost_hist[BA_MAX_ITER];   /* candidate cost evaluated by iteration i */
  double lambda_hist[BA_MAX_ITER];      /* damping used by iteration i */
  uint8_t accepted[BA_MAX_ITER];        /* 1 = step accepted */
  int32_t solve_fail[BA_MAX_ITER];      /* 1 = reduced system not SPD / PCG breakdown */
  int32_t pcg_iters[BA_MAX_ITER];
  float gpu_ms;                         /* device time of the whole solve (CUDA events) */
  int32_t n_launches;                   /* kernels launched by this solve */
  float ms_linearize;                   /* profile=1 only: summed device time per kernel class */
  float ms_solve;
  float ms_backsub;
  float ms_other;                       /* init + finalize */
  int32_t band_width;                   /* BA_SOLVER_BAND: co-visibility bandwidth in keyframes, else -1 */
  int32_t s_blocks;                     /* 6x6 blocks of the reduced camera matrix that are stored / all-reduced */
  float ms_comm;                        /* profile=1, sharded: NCCL collectives (candidate cost + solve-failed flag; the reduced camera
                                           system and the camera-state broadcast where the peer exchange is off) */
  float ms_point_blocks;                /* profile=1, panel path: the point-block pass inside ms_linearize */
  float ms_cam_blocks;                  /* profile=1, panel path: the camera-major pass inside ms_linearize */
  int32_t peer_exchange;                /* sharded: 1 = the solver's first kernel summed the ranks' reduced camera systems over
                                           NVLink peer memory (its time is inside ms_solve), 0 = ncclAllReduce in front of it */
} ba_report;

typedef struct ba_handle ba_handle;

int ba_version(void);
const char* ba_error_string(int code);
const char* ba_last_error(const ba_handle* h);
void ba_default_params(ba_params* p);

int ba_create(const ba_params* params, int device, ba_handle** out);
int ba_destroy(ba_handle* h);
int ba_set_params(ba_handle* h, const ba_params* params);

/* Local BA, device pointers.  cam_R/cam_t/X are refined in place; flags (may be NULL) gets one
 * uint8 per observation.  report (host, may be NULL) is filled after a stream synchronise; with
 * report == NULL the call only enqueues work. */
int ba_solve_local(ba_handle* h, int32_t n_cam, int32_t n_pt, int32_t n_obs,
                   const double intr[4],
                   double* cam_R, double* cam_t, const uint8_t* cam_fixed,
                   double* X,
                   const int32_t* obs_cam, const int32_t* obs_pt, const float* obs_uv,
                   const int32_t* pt_offsets,
                   uint8_t* flags, void* stream, ba_report* report);

/* Same, HOST pointers; H2D and D2H copies are inside the call.  obs_pt is NOT uploaded: the device copy is
 * expanded from pt_offsets (the two describe the same thing; the argument may be NULL here).  The small pose
 * arrays travel as one copy each way through a pinned staging buffer of the handle. */
int ba_solve_local_host(ba_handle* h, int32_t n_cam, int32_t n_pt, int32_t n_obs,
                        const double intr[4],
                        double* cam_R, double* cam_t, const uint8_t* cam_fixed,
                        double* X,
                        const int32_t* obs_cam, const int32_t* obs_pt, const float* obs_uv,
                        const int32_t* pt_offsets,
                        uint8_t* flags, ba_report* report);

/* Batched pose-only LM (one free pose per frame, fixed points).  R/t refined in place.
 * cost_hist (may be NULL): n_frames x (max_iter+1) doubles; accepted (may be NULL):
 * n_frames x max_iter bytes. */
int ba_solve_pose_only(ba_handle* h, int32_t n_frames, int32_t n_match,
                       const double intr[4],
                       double* R, double* t,
                       const float* X, const float* uv, const int32_t* frame_offsets,
                       uint8_t* flags, double* cost_hist, uint8_t* accepted, void* stream);

int ba_solve_pose_only_host(ba_handle* h, int32_t n_frames, int32_t n_match,
                            const double intr[4],
                            double* R, double* t,
                            const float* X, const float* uv, const int32_t* frame_offsets,
                            uint8_t* flags, double* cost_hist, uint8_t* accepted, float* gpu_ms);

/* Pose of ONE frame without a prior, HOST pointers: RANSAC over `n_hyp` minimal three-point (P3P) samples on the device,
 * every admissible root scored by the number of the frame's matches that reproject within `thr_px`; the best pose comes
 * back in R_out (9, row-major, world->camera) / t_out (3) with its inlier count (0: no admissible hypothesis).  This is the
 * bootstrap cv2.solvePnPRansac provides at the reference's call site, which passes no initial guess
 * (visual_slam/utils/motion_estimation.py:107-116, visual_slam/tracking.py:350-356); ba_solve_pose_only refines it. */
int ba_pose_ransac_host(ba_handle* h, int32_t n_match, const double intr[4], const float* X, const float* uv,
                        int32_t n_hyp, double thr_px, uint32_t seed, double* R_out, double* t_out, int32_t* n_inliers);

/* Objective only (reference cost, torch_optimizer.py:152-177) at the given state; device pointers. */
int ba_eval_cost(ba_handle* h, int32_t n_cam, int32_t n_pt, int32_t n_obs, const double intr[4],
                 const double* cam_R, const double* cam_t, const double* X,
                 const int32_t* obs_cam, const int32_t* obs_pt, const float* obs_uv,
                 void* stream, double* cost_out /* host */);

/* Multi-GPU: shard by map point; the reduced camera system and the costs are all-reduced. */
int ba_nccl_unique_id(void* out_id /* BA_NCCL_ID_BYTES */);
int ba_dist_init(ba_handle* h, const void* id /* BA_NCCL_ID_BYTES */, int32_t rank, int32_t world);
int ba_dist_shutdown(ba_handle* h);

/* Debug/inspection for tests: copy the reduced camera system of the LAST linearisation
 * to host buffers: S block-packed lower triangle (6x6 block (I,J), J <= I, at ((I(I+1)/2+J)*36,
 * row-major inside; n_cam(n_cam+1)/2*36 doubles), b and diagU (6 n_cam each). Any may be NULL. */
int ba_debug_linearize(ba_handle* h, int32_t n_cam, int32_t n_pt, int32_t n_obs,
                       const double intr[4], const double* cam_R, const double* cam_t,
                       const uint8_t* cam_fixed, const double* X,
                       const int32_t* obs_cam, const int32_t* obs_pt, const float* obs_uv,
                       const int32_t* pt_offsets, double lambda, void* stream,
                       double* S_host, double* b_host, double* diagU_host, double* cost_host);

#ifdef __cplusplus
}
#endif
#endif /* B200BA_H */
