// Kernel-facing view of one local-BA solve and the launch wrappers (defined in ba_local.cu,
// ba_pose.cu) used by the C ABI in ba_api.cu.
#pragma once
#include "ba_device.cuh"

namespace ba {

// Plain-old-data view handed to every kernel by value.
struct Work {
  int n_cam, n_pt, n_obs, n;          // n = 6 * n_cam (dense reduced camera system order)
  Intr K;
  const int* obs_cam;                 // [n_obs]
  const int* obs_pt;                  // [n_obs] non-decreasing
  const float* obs_uv;                // [2 n_obs]
  const int* pt_off;                  // [n_pt + 1]
  const unsigned char* fixed;         // [n_cam]
  double* pose[2];                    // [12 n_cam] each: R row-major (9) then t (3); double-buffered
  double* X[2];                       // [3 n_pt] each; double-buffered
  double* S;                          // reduced camera matrix, block-packed lower triangle: 6x6 block
                                      // (I,J), J <= I, at s_block_off(I,J), row-major inside (reduction buffer)
  double* b;                          // [n]
  double* diagU;                      // [n]   diag of the un-damped pose blocks
  double* cost_lin;                   // [1]   cost at the linearisation point (same buffer as S)
  double* cost_cand;                  // [2]   candidate cost; [1] = solve-failed flag of this rank (sharded runs)
  double* dp;                         // [n]   camera step
  LmCtl* ctl;
  int world;                          // > 1: accept/reject runs as its own kernel after the all-reduce
  long long* dbg;                     // optional clock64() trace (BA_DBG_TIMES=1), null otherwise
  int* part;                          // [part_n + 1] first point of every CTA of the mma linearise kernel (equal
  int part_n;                         // observation shares), written once per solve by k_init; null / 0: none
  // panel linearise path (ba_panel.cu): per-point records and the camera-sorted copy of the observations
  double* rec;                        // [10 (n_pt + 1)] C (6) | q (3) | cmin, cmax | flags
  int* cs_cnt;                        // [n_cam] observation counts, then fill cursors
  int* cs_blk_off;                    // [n_cam + 1] first 32-entry block of every camera
  int* cs_blk_cam;                    // [cs_nblk_max] camera of every block
  int* cs_pt;                         // [32 cs_nblk_max] point of every entry (-1: pad)
  float2* cs_uv;                      // [32 cs_nblk_max]
  int* cs_nblk;                       // [1] blocks in use
  int cs_nblk_max;
  // class order of the panel path (ba_panel.cu, launch_panel_reorder): when perm != null, obs_* / pt_off / X / rec are
  // in the solver's own point order (points sorted by first camera and camera span) and position p holds the caller's
  // point perm[p]; k_init / k_finalize translate, nothing else knows
  const int* perm;                    // [n_pt]
  const int* pt_off_in;               // [n_pt + 1] the caller's offsets (flags go back to the caller's observation order)
  // block-sparse reduced camera matrix (large windows); all null / 0 for the dense packed layout
  const unsigned* bsr_bits;           // [n_cam * bsr_W] structure bitmap (symmetric, diagonal set)
  const int* bsr_wpre;                // [n_cam * bsr_W] set bits before each word of its row
  const int* bsr_row_ptr;             // [n_cam + 1]
  const int* bsr_col;                 // [nnzb]
  const int* bsr_tpos;                // [nnzb] position of the transposed block
  int bsr_W, nnzb;
  // block-banded reduced camera matrix (sequential windows): S_band[I][d] = block (I, I-d), d = 0..band_w;
  // band_w < 0 when the layout is not banded
  int band_w;
  double* band_L;                     // [n_cam][band_w+1][36] columns of the Cholesky factor
  double* band_y;                     // [n]  L^-1 b
  double* band_dinv;                  // [n]  1 / diag(L)
  double* bcr_ws;                     // workspace of the cyclic-reduction band solver (null: sequential band solver)
  size_t s_len;                       // doubles in S (dense packed: n_cam(n_cam+1)/2*36, sparse: nnzb*36)
  // block-Jacobi PCG workspace (sparse path)
  double* pcg_r; double* pcg_z; double* pcg_p[2]; double* pcg_Ap; double* pcg_Minv; double* pcg_acc;
  int pcg_max_iter;
  double pcg_tol;
};

// phase timestamps of CTA 0 / thread 0 for latency hunting (store only: no load on the critical path)
__device__ __forceinline__ void dbg_mark(const Work& w, int slot, int tag) {
  if (w.dbg && blockIdx.x == 0 && threadIdx.x == 0 && slot < 2000) {
    w.dbg[1 + 2 * slot] = tag;
    w.dbg[2 + 2 * slot] = clock64();
  }
}

__host__ __device__ __forceinline__ size_t s_block_off(int I, int J) {
  return ((size_t)I * (I + 1) / 2 + J) * 36;
}
__host__ __device__ __forceinline__ size_t s_doubles(int n_cam) {
  return (size_t)n_cam * (n_cam + 1) / 2 * 36;
}

// position of block (a, b) in the block-sparse layout (the block must exist)
__device__ __forceinline__ int bsr_find(const Work& w, int a, int b) {
  const int word = b >> 5;
  const size_t at = (size_t)a * w.bsr_W + word;
  const unsigned m = __ldg(w.bsr_bits + at);
  return __ldg(w.bsr_row_ptr + a) + __ldg(w.bsr_wpre + at) + __popc(m & ((1u << (b & 31)) - 1u));
}
// where block (a, b), a >= b, of the reduced camera matrix is accumulated
__device__ __forceinline__ double* s_block_ptr(const Work& w, int a, int b) {
  if (w.band_w >= 0) return w.S + ((size_t)a * (w.band_w + 1) + (a - b)) * 36;
  if (w.bsr_bits == nullptr) return w.S + s_block_off(a, b);
  return w.S + (size_t)bsr_find(w, a, b) * 36;
}

// launch wrappers; every one returns the number of kernels it launched
struct LmInit {
  double lambda0, huber_delta, chi2_thr, lambda_up, lambda_down, lambda_min, lambda_max, accept_rtol;
  long long epoch_base;   // 0 unless the peer-memory exchange is on
};
int launch_init(const Work& w, const double* R_in, const double* t_in, const double* X_in,
                const LmInit& p, cudaStream_t s);
int launch_linearize_scatter(const Work& w, int sm_count, cudaStream_t s);
int launch_linearize_tiled(const Work& w, int sm_count, cudaStream_t s);
int launch_linearize_stream(const Work& w, int sm_count, cudaStream_t s);
int configure_stream();
int launch_linearize_mma(const Work& w, int sm_count, cudaStream_t s);
int launch_linearize_panel(const Work& w, int sm_count, cudaStream_t s, cudaEvent_t* marks = nullptr);   // point blocks + camera blocks + Schur panels
int launch_panel_setup(const Work& w, int sm_count, cudaStream_t s);       // once per solve: camera-sorted observations
int configure_panel();
int panel_window();
size_t panel_rec_doubles(int n_pt);
int panel_cs_blocks_max(int n_cam, int n_obs);
bool panel_supported(const Work& w);
// class order (once per solve, before launch_init): points sorted by (first camera, camera span), observations copied
// along.  Buffers are the caller's workspace: cnt [panel_classes + 1] 64-bit, perm [n_pt], off [n_pt + 1],
// cam / pt [n_obs], uv [2 n_obs].  `w` keeps the caller's arrays; the caller switches the pointers afterwards.
struct PanelOrder {
  int* range;                               // [2] first / last camera of the shard (min / max over the points)
  unsigned long long* cnt;
  int* perm;
  int* off;
  int* cam;
  int* pt;
  float* uv;
};
size_t panel_classes(int n_cam);
int launch_panel_reorder(const Work& w, const PanelOrder& o, cudaStream_t s);
int configure_mma();
int mma_grid(int n_cam, int n_pt, int sm_count);   // CTAs launch_linearize_mma will use
int mma_window();
int launch_rcs_cholesky(const Work& w, cudaStream_t s);
int launch_symbolic(const Work& w, unsigned* bits, int* wpre, int* row_nnz, int* row_ptr, cudaStream_t s);
int launch_symbolic_mark(const Work& w, unsigned* bits, cudaStream_t s);
int launch_symbolic_or(const Work& w, const unsigned* all, int world, unsigned* bits, cudaStream_t s);
int launch_symbolic_count(const Work& w, unsigned* bits, int* wpre, int* row_nnz, int* row_ptr, cudaStream_t s);
int launch_symbolic_fill(const Work& w, int* col_idx, int* tpos, cudaStream_t s);
int pcg_grid_size(int n_cam, int sm_count);
int launch_rcs_pcg(const Work& w, int grid, cudaStream_t s);
int launch_band_width(const Work& w, int* out, cudaStream_t s);
int launch_rcs_band(const Work& w, cudaStream_t s);
bool bcr_supported(int n_cam, int band_w);
size_t bcr_workspace_doubles(int n_cam);
int configure_bcr();
// Sharded runs, banded path: instead of an NCCL all-reduce of S | b | diagU | cost in front of the solver, the solver's
// first kernel (k_bcr_assemble) sums the ranks' reduction buffers itself over NVLink peer memory (CUDA IPC mappings)
// while it builds level 0 -- one kernel for the exchange and the assembly.  red[q] = rank q's reduction buffer (own
// at [rank]); flags: one 64-bit epoch per (owner, poster); a rank posts "my linearisation of epoch e is complete" into
// every peer's flag array and waits until every peer has posted e into its own (e = LmCtl::epoch_base + LM iteration + 1,
// from device memory: the LM loop is a replayed CUDA graph).  world <= 1: no exchange.
constexpr int kMaxPeers = 8;
struct PeerView {
  int world, rank;
  const double* red[kMaxPeers];
  long long* flag_remote[kMaxPeers];        // this rank's slot in rank q's flag array (4 words: epoch | first camera | last camera | -)
  const long long* flag_local;              // this rank's flag array, one slot per poster
  const int* cam_range;                     // [2] first / last camera this rank's shard observes
};
int launch_rcs_bcr(const Work& w, double* ws, cudaStream_t s, int sm_count = 0, const PeerView* pv = nullptr, int it = 0);
int configure_band();
int band_max_w();
int launch_backsub(const Work& w, int sm_count, cudaStream_t s);
int launch_accept(const Work& w, cudaStream_t s);
int launch_cost0(const Work& w, cudaStream_t s);   // max_iter == 0: cost_cand -> cost_hist[0]
int launch_finalize(const Work& w, double* R_out, double* t_out, double* X_out,
                    unsigned char* flags, cudaStream_t s);
int launch_validate(int n_cam, int n_pt, int n_obs, const int* obs_cam, const int* pt_off, int* bad, cudaStream_t s);
int launch_expand_obs_pt(const int* pt_off, int n_pt, int* obs_pt, cudaStream_t s);   // CSR offsets -> point of every observation
int launch_eval_cost(const Work& w, const double* pose, const double* X, double* out,
                     double delta, cudaStream_t s);
int configure_tiled();
int configure_kernels();   // per-device function attributes; call once after cudaSetDevice
size_t rcs_cholesky_smem_bytes(int n);
int rcs_cholesky_max_n();
bool tiled_supported(int n_cam);

struct PoseWork {
  int n_frames;
  Intr K;
  double* R;                 // [9 n_frames] in/out
  double* t;                 // [3 n_frames] in/out
  const float* X;            // [3 n_match]
  const float* uv;           // [2 n_match]
  const int* frame_off;      // [n_frames+1]
  unsigned char* flags;      // [n_match] or null
  double* cost_hist;         // [n_frames*(max_iter+1)] or null
  unsigned char* accepted;   // [n_frames*max_iter] or null
  int max_iter;
  double huber_delta, chi2_thr, lambda0, lambda_up, lambda_down, lambda_min, lambda_max, accept_rtol;
};
int launch_pose_only(const PoseWork& w, cudaStream_t s);

// pose without a prior: RANSAC over minimal P3P samples (ba_ransac.cu)
struct RansacWork {
  int n_match, n_hyp;
  Intr K;
  const float* X;            // [3 n_match] world points
  const float* uv;           // [2 n_match] pixels
  double thr2;               // squared reprojection threshold (px^2)
  unsigned seed;
  double* out;               // [n_hyp][16]: inliers | R (9) | t (3)
};
int ransac_out_doubles();
int launch_pose_ransac(const RansacWork& w, cudaStream_t s);

}  // namespace ba
