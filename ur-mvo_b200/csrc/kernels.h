// csrc/kernels.h — host-side launchers exported by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ba_types.h"

namespace urmvo {

// ---- ba_kernels.cu
// Dynamic shared memory of the BA kernels: work_stride doubles + ints_per_warp ints per warp.
size_t ba_smem_bytes(int threads, int work_stride, int ints_per_warp);
int ba_stage_doubles(int kmax, int stereo = 0);  // staging fields of the one-point-per-warp modes (stereo: 3-row edges)
int ba_pack_doubles();           // staging fields of the packed modes
int pcg_dense_doubles(int Ncf, int threads);  // shared memory of the dense in-smem PCG
int ba_tile_doubles();           // atomic mode: per-warp transposition tile for coalesced REDs
// Batched windows: grid = n_clusters * cluster_size CTAs, one cluster per window.
cudaError_t launch_ba_cluster(const BAWin* wins_dev, const BARun& run, int mode, int kmax, int work_stride,
                              int ints_per_warp, int n_clusters, int cluster_size, int threads,
                              cudaStream_t stream);
// One (or a few) large problems on a cooperative grid. grid_blocks <= co-resident capacity.
cudaError_t launch_ba_grid(const BAWin* wins_dev, const BARun& run, int kmax, int grid_blocks,
                           int threads, cudaStream_t stream, int stereo = 0);
// Max co-resident CTAs of the grid kernel on the current device for this configuration.
int ba_grid_capacity(int threads, int kmax, int stereo = 0);
cudaError_t ba_timing_read(unsigned long long* out, bool reset);
// point-sharded BA: phase kernels on a cooperative grid, NCCL all-reduces in between (capi.cu)
size_t shard_state_bytes();
int shard_grid_capacity(int threads, int kmax);
cudaError_t launch_sh_init(const BAWin* w, void* stt, int grid, int threads, cudaStream_t s);
cudaError_t launch_sh_begin_pass(void* stt, int robust, cudaStream_t s);
cudaError_t launch_sh_lin(const BAWin* w, const BARun& run, void* stt, double* scal, int kmax, int diag,
                          int grid, int threads, cudaStream_t s);
cudaError_t launch_sh_lambda(const BAWin* w, void* stt, const double* scal, cudaStream_t s);
cudaError_t launch_sh_solve(const BAWin* w, const BARun& run, void* stt, double* scal, int kmax, int rank,
                            int grid, int threads, cudaStream_t s);
cudaError_t launch_sh_decide(void* stt, const double* scal, void* host_copy, cudaStream_t s);
cudaError_t launch_sh_classify(const BAWin* w, const BARun& run, void* stt, int pass, int grid, int threads,
                               cudaStream_t s);
cudaError_t launch_sh_finish(const BAWin* w, void* stt, int grid, int threads, cudaStream_t s);
void shard_flags(const void* host_copy, int* cont_trials, int* terminate);
// large problems in tile mode (ba_large.cu): plain launches on the context stream, no grid barriers
size_t lg_state_bytes();
cudaError_t lg_timing_read(unsigned long long* out, bool reset);
int lg_band_max_m();
size_t band_smem_bytes(int M, int Ncf);
cudaError_t lg_prepare(int M, int Ncf);
cudaError_t launch_lg_init(const BAWin* w, void* stt, int grid, cudaStream_t s);
cudaError_t launch_lg_begin_pass(void* stt, int robust, int n_iter, cudaStream_t s);
cudaError_t launch_lg_pack(const BAWin* w, int grid, cudaStream_t s);
cudaError_t launch_lg_lin(const BAWin* w, const BARun& run, void* stt, double* scal, int diag, int rank, int grid,
                          cudaStream_t s);
cudaError_t launch_lg_lambda(const BAWin* w, void* stt, const double* scal, int world, cudaStream_t s);
cudaError_t launch_lg_solve(const BAWin* w, void* stt, int M, int Ncf, cudaStream_t s);
cudaError_t launch_lg_backsub(const BAWin* w, const BARun& run, void* stt, double* scal, int grid, cudaStream_t s);
cudaError_t launch_lg_decide(void* stt, const double* scal, void* host_copy, cudaStream_t s);
cudaError_t launch_lg_classify(const BAWin* w, const BARun& run, void* stt, int pass, int grid, cudaStream_t s);  // 2 launches
cudaError_t launch_lg_finish(const BAWin* w, void* stt, int grid, cudaStream_t s);
void lg_flags(const void* host_copy, int* active, int* it);
// long block-banded systems: block cyclic reduction instead of the sequential band solve (ba_bcr.cu)
constexpr int kBcrMaxMb = 108;      // scalars per super-block (18 cameras)
constexpr int kBcrMaxLevels = 12;
struct BcrShape {
  int n, m, mb, K, L, M;            // free cameras, cameras / scalars per super-block, super-blocks, levels, band row stride
  size_t coff[kBcrMaxLevels + 1];   // first coupling block of each level
  size_t off_D, off_C, off_rhs, off_x, off_invd, off_fail, total;  // doubles, into the workspace
};
bool bcr_shape(int n, int bw, int force, BcrShape* out);
cudaError_t bcr_prepare(const BcrShape& sh);
cudaError_t launch_bcr_solve(const BAWin* w, void* stt, const BcrShape& sh, double* work, cudaStream_t s, int* n_launch);
constexpr int kBAPartWidth = 8;  // doubles per CTA and buffer in BAWin::part (+8 flag doubles)

// ---- pose_kernels.cu
cudaError_t launch_pose_only(int B, const int* obs_off, const double* pose_in, const double* uv,
                             const double* Xw, const double* intr, double chi2_thr, double delta,
                             int rounds, int its_per_round, uint8_t* inlier, uint8_t* level,
                             double* pose_out, int* n_inlier, int* lm_iters, cudaStream_t stream,
                             const double* ur = nullptr, const uint8_t* kind = nullptr, const double* tab = nullptr,
                             double chi2_thr_s = 0.0, double delta_s = 0.0);

// ---- twoview_kernels.cu
struct TVMotionOut {
  int used_H;
  int n_motion;
  int n_inl;
  int n_good[8];
  float cos_kth[8];
  float R[8][9];
  float t[8][3];
};
struct TVBuffers {
  int n1, n2, N, n_hyp, words;
  const float *keys1, *keys2;
  const int *m1, *m2, *sets;
  const float* K;
  float *pn1, *pn2, *T1, *T2;
  float4 *uv, *pnm;
  float* models;      // [2][n_hyp][18]
  float* scores;      // [2][n_hyp]
  uint32_t* masks;    // [2][n_hyp][words]
  int* best_idx;      // [2]
  float* best_score;  // [2]
  float* P3D;         // [8][n1*3]
  uint8_t* good;      // [8][n1]
  float* cosbuf;      // [8][N]
  TVMotionOut* motion;
};
// normalise + gather + fit + score + arg-max: 5 launches. Returns launches made in *n_launch.
cudaError_t launch_tv_ransac(const TVBuffers& b, float sigma, int score_mode, int n_sm, cudaStream_t stream, int* n_launch);
cudaError_t launch_tv_motion(const TVBuffers& b, float th2, cudaStream_t stream);

// ---- fm_kernels.cu (per-frame fundamental-matrix RANSAC; pts = (x0,y0,x1,y1) per match).
// One RANSAC round evaluates n_hyp hypotheses; hyp_ids[i] = problem * max_iters + iteration addresses
// the persistent model store models_all [B*max_iters][27]; sets / n_models / counts are per round.  Problems with
// fewer than 15 matches are OpenCV's LMedS branch: their counts are the bits of the median float error, and
// thr_b (per problem, may be NULL) carries their own squared threshold (negative: flag every match).
cudaError_t launch_fm_solve(int n_hyp, const int* hyp_ids, const int* sets, const float4* pts, double* models_all,
                            int* n_models, cudaStream_t s);
cudaError_t launch_fm_score(int n_hyp, const int* hyp_ids, int max_iters, const int* off, const float4* pts,
                            const double* models_all, const int* n_models, float thr2, int* counts, int n_sm,
                            cudaStream_t s);
cudaError_t launch_fm_mask(int B, int max_n, const int* off, const float4* pts, const double* models_all,
                           const int* win_id, float thr2, const float* thr_b, uint8_t* mask, double* win_F,
                           cudaStream_t s);

// ---- map_kernels.cu (device-resident map: slot-addressed keyframe / mappoint / observation arrays)
cudaError_t launch_map_gather(double* pose_in, double* pts_in, double* uv, const double* d_kf, const double* d_pt,
                              const double* d_uv, const int* kf_slot, const int* pt_slot, const int* obs_slot, int Nc,
                              int Np, int No, cudaStream_t s);
cudaError_t launch_map_scatter(double* d_kf, double* d_pt, const double* pose_out, const double* pts_out, const int* cam_free,
                               const int* kf_slot, const int* pt_slot, int Nc, int Np, cudaStream_t s);
cudaError_t launch_map_set(double* dst, const double* vals, const int* slot, int n, int W, cudaStream_t s);
cudaError_t launch_map_get(double* vals, const double* src, const int* slot, int n, int W, cudaStream_t s);

// ---- pnp_kernels.cu (SolvePnPWithCV: EPnP RANSAC hypotheses, inlier counting, refinement over the inliers).
// H = B * max_iters hypotheses; sets [H][5] indices local to the problem; models [H][12] = R | t (T_cw);
// masks [H][words_max]; counts[h] = inliers or -1 (no model).
cudaError_t launch_pnp_hypotheses(int H, int max_iters, int words_max, const int* off, const int* sets, const float* obj,
                                  const float* img, const double* K4, double* models, int* valid, float thr2,
                                  unsigned* masks, int* counts, cudaStream_t s);
cudaError_t launch_pnp_refine(int B, int max_iters, int words_max, const int* off, const float* obj, const float* img,
                              const double* K4, const double* models, const unsigned* masks, const int* best,
                              double* out_Rt, uint8_t* inlier, cudaStream_t s);

}  // namespace urmvo
