/* include/urmvo_b200.h — C ABI of the B200-native geometric back-end for UR-MVO.
 *
 * Drop-in boundary for the reference's hot path:
 *   LocalmapOptimization   /root/reference/include/g2o_optimization.h:13-15  (src/g2o_optimization.cc:20-177)
 *   FrameOptimization      /root/reference/include/g2o_optimization.h:17-19  (src/g2o_optimization.cc:179-321)
 *   EpipolarGeometry::reconstruct  /root/reference/include/epipolar_geometry.h:31-35 (src/epipolar_geometry.cc:18-98)
 *   cv::findFundamentalMat(FM_RANSAC) call of PointMatching::MatchingPoints  (src/point_matching.cc:50-60)
 *   Mapping::TriangulateMappoint, batched  (src/mapping.cc:151-205)
 * The C++ adapters in ur-mvo_b200/adapter/ keep those signatures and flatten the reference's
 * MapOfPoses / MapOfPoints3d / constraint vectors / cv::KeyPoint into the SoA arrays below.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a HOST pointer unless the name says _dev.
 *  - every function returns 0 on success, <0 (urmvo_status) on error; nothing throws or aborts;
 *    urmvo_last_error() describes the last failure on the calling thread.
 *  - a context owns one CUDA device, one non-blocking stream and its device workspaces.  Calls on
 *    one context must not overlap; different contexts are independent (re-entrant per handle).
 *    The legacy default stream is never used and cudaDeviceSynchronize is never called.
 *  - there is NO CPU fallback: without a usable sm_100 device every entry point fails loudly.
 *  - poses are T_wc as (qx,qy,qz,qw,px,py,pz) doubles — the reference's Pose3d (include/types.h:18-31).
 */
#ifndef URMVO_B200_H_
#define URMVO_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct urmvo_ctx urmvo_ctx;
typedef struct urmvo_ba_plan urmvo_ba_plan;       /* device-resident batch of BA windows */
typedef struct urmvo_pose_plan urmvo_pose_plan;   /* device-resident batch of pose-only frames */
typedef struct urmvo_tv_plan urmvo_tv_plan;       /* device-resident two-view problem */
typedef struct urmvo_fm_plan urmvo_fm_plan;       /* device-resident batch of fundamental-matrix RANSAC problems */

typedef enum {
  URMVO_OK = 0,
  URMVO_ERR_NO_DEVICE = -1,   /* no CUDA device / not sm_100 / driver failure */
  URMVO_ERR_CUDA = -2,        /* a CUDA runtime call failed */
  URMVO_ERR_ARG = -3,         /* invalid argument */
  URMVO_ERR_NCCL = -4,        /* NCCL unavailable or a collective failed */
  URMVO_ERR_UNSUPPORTED = -5
} urmvo_status;

int urmvo_version(void);
const char* urmvo_last_error(void);

int urmvo_create(urmvo_ctx** ctx, int device);
void urmvo_destroy(urmvo_ctx* ctx);
/* The context's cudaStream_t (as void*), so that callers can record CUDA events around plan runs. */
void* urmvo_stream(urmvo_ctx* ctx);
int urmvo_sync(urmvo_ctx* ctx);
/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t urmvo_launch_count(urmvo_ctx* ctx);

/* ------------------------------------------------------------------ local BA (B1-B6) */

typedef struct {
  double pcg_tol;        /* relative tolerance on sqrt(r^T M^-1 r); 0 -> 1e-10 */
  int32_t pcg_max_iter;  /* 0 -> max(60, 2*6*Ncf) capped at 1000 */
  int32_t cluster_size;  /* CTAs cooperating on one window in batch mode (1,2,4,8,16); 0 -> auto */
  int32_t threads;       /* threads per CTA; 0 -> auto */
  int32_t force_atomic;  /* 1: always accumulate S with global fp64 atomics (default: shared-memory copies when they fit) */
  int32_t dense_solver;  /* windows whose reduced camera system is solved in shared memory (<= 16 free cameras):
                            0 -> direct tiled Cholesky (default: what g2o's LinearSolverEigen does, 3x faster than the
                                 other two; falls back to 1 when the tiles do not fit the CTA),
                            1 -> direct LDL^T with one barrier per pivot,
                            2 -> block-Jacobi PCG (tolerance pcg_tol; the solver BASELINE.json's north_star names) */
  int32_t large_mode;    /* one large window (>= 100k observations) and the point-sharded solve:
                            0 -> tile mode when it fits (points renumbered along the trajectory, S as a block band,
                                 direct block-banded Cholesky — csrc/ba_large.cu), else the atomic / PCG kernels,
                            1 -> always the round-1 path (global fp64 atomics + block-Jacobi PCG),
                            2 -> tile mode or URMVO_ERR_UNSUPPORTED */
  int32_t band_solver;   /* tile mode, direct solve of the block-banded reduced camera system:
                            0 -> block cyclic reduction for long trajectories (>= 160 free cameras: log-depth
                                 elimination on many SMs, csrc/ba_bcr.cu), else the sequential block-banded Cholesky,
                            1 -> always the sequential band Cholesky (one SM, one block column after the other),
                            2 -> cyclic reduction whenever the shape allows it (>= 2 super-blocks) */
} urmvo_ba_options;

typedef struct {
  int32_t iters[2];          /* outer LM iterations run by optimize(it0) / optimize(it1) */
  int32_t trials[2];         /* damped solves (trial steps) in each call */
  int32_t pcg_iters[2];      /* PCG iterations summed over the trials of each call */
  int32_t n_level1;          /* observations excluded from the second optimisation */
  double chi2_initial;       /* robust chi2 at the input estimate */
  double chi2_final[2];      /* currentChi after each optimize() call */
  double lambda_final[2];
} urmvo_ba_stats;

/* One-shot LocalmapOptimization on host buffers (mono edges).
 *  poses  Nc*7 in/out (T_wc), fixed Nc bytes, pts Np*3 in/out,
 *  uv No*2, cam / pt: No dense indices (into poses / pts).  Observations may be in any order.
 *  intr = fx,fy,cx,cy.  chi2_thr = cfg.mono_point (Huber delta = (double)(float)sqrt(chi2_thr)).
 *  it0 / it1 = 10 / 5 in the reference.  inlier: No bytes out.  stats may be NULL. */
int urmvo_local_ba(urmvo_ctx* ctx, int Nc, double* poses, const uint8_t* fixed, int Np, double* pts,
                   int No, const double* uv, const int32_t* cam, const int32_t* pt,
                   const double* intr, double chi2_thr, int it0, int it1, uint8_t* inlier,
                   urmvo_ba_stats* stats, const urmvo_ba_options* opts);

/* Batch of B independent windows, concatenated: window w owns cams [cam_off[w],cam_off[w+1]),
 * points [pt_off[w],pt_off[w+1]) and observations [obs_off[w],obs_off[w+1]); cam/pt indices are
 * LOCAL to the window.  stats: B entries or NULL. */
int urmvo_local_ba_batch(urmvo_ctx* ctx, int B, const int32_t* cam_off, const int32_t* pt_off,
                         const int32_t* obs_off, double* poses, const uint8_t* fixed, double* pts,
                         const double* uv, const int32_t* cam, const int32_t* pt, const double* intr,
                         double chi2_thr, int it0, int it1, uint8_t* inlier, urmvo_ba_stats* stats,
                         const urmvo_ba_options* opts);

/* The same two calls for a STEREO camera (reference src/g2o_optimization.cc:96-118: when the camera type is
 * STEREO the graph holds the mono edges AND one EdgeStereoSE3ProjectXYZ per stereo constraint).
 *  uv3 No*3 = (u_left, v_left, u_right); kind No bytes: 1 = stereo edge (3-row residual, Omega = I3, Huber delta
 *  (float)sqrt(chi2_thr_stereo), outlier threshold chi2_thr_stereo = cfg.stereo_point), 0 = mono edge (u_right
 *  ignored, chi2_thr_mono = cfg.mono_point); intr5 = fx, fy, cx, cy, bf.  Mono and stereo constraints of the
 *  reference's two vectors are concatenated by the caller; inlier comes back in the same order. */
int urmvo_local_ba_stereo(urmvo_ctx* ctx, int Nc, double* poses, const uint8_t* fixed, int Np, double* pts,
                          int No, const double* uv3, const uint8_t* kind, const int32_t* cam, const int32_t* pt,
                          const double* intr5, double chi2_thr_mono, double chi2_thr_stereo, int it0, int it1,
                          uint8_t* inlier, urmvo_ba_stats* stats, const urmvo_ba_options* opts);
int urmvo_local_ba_batch_stereo(urmvo_ctx* ctx, int B, const int32_t* cam_off, const int32_t* pt_off,
                                const int32_t* obs_off, double* poses, const uint8_t* fixed, double* pts,
                                const double* uv3, const uint8_t* kind, const int32_t* cam, const int32_t* pt,
                                const double* intr5, double chi2_thr_mono, double chi2_thr_stereo, int it0, int it1,
                                uint8_t* inlier, urmvo_ba_stats* stats, const urmvo_ba_options* opts);

/* The same two calls with SEVERAL camera models in one graph.  The reference reads fx, fy, cx, cy (and BF) per
 * constraint from camera_list[mpc->id_camera] (src/g2o_optimization.cc:86-89, :106-113); the calls above take one
 * intrinsics set, which is what every configuration of the reference has (one camera).  Here intr5_tab holds
 * n_models (1..128) rows of (fx, fy, cx, cy, bf) and kind_model[o] = (1 if stereo edge, else 0) | (id_camera << 1);
 * everything else as in the stereo calls (a mono rig passes bf = 0 and no stereo bit).  These windows run on the
 * one-point-per-warp accumulation modes (the packed / tile modes are single-camera). */
int urmvo_local_ba_multicam(urmvo_ctx* ctx, int Nc, double* poses, const uint8_t* fixed, int Np, double* pts,
                            int No, const double* uv3, const uint8_t* kind_model, const int32_t* cam,
                            const int32_t* pt, int n_models, const double* intr5_tab, double chi2_thr_mono,
                            double chi2_thr_stereo, int it0, int it1, uint8_t* inlier, urmvo_ba_stats* stats,
                            const urmvo_ba_options* opts);
int urmvo_local_ba_batch_multicam(urmvo_ctx* ctx, int B, const int32_t* cam_off, const int32_t* pt_off,
                                  const int32_t* obs_off, double* poses, const uint8_t* fixed, double* pts,
                                  const double* uv3, const uint8_t* kind_model, const int32_t* cam,
                                  const int32_t* pt, int n_models, const double* intr5_tab, double chi2_thr_mono,
                                  double chi2_thr_stereo, int it0, int it1, uint8_t* inlier, urmvo_ba_stats* stats,
                                  const urmvo_ba_options* opts);

/* Plan API: upload once, run many times from HBM-resident inputs (each run restarts from the
 * uploaded initial estimate), download when wanted. */
int urmvo_ba_plan_create(urmvo_ctx* ctx, urmvo_ba_plan** plan, int B, const int32_t* cam_off,
                         const int32_t* pt_off, const int32_t* obs_off, const double* poses,
                         const uint8_t* fixed, const double* pts, const double* uv,
                         const int32_t* cam, const int32_t* pt, const double* intr,
                         double chi2_thr, int it0, int it1, const urmvo_ba_options* opts);
int urmvo_ba_plan_run(urmvo_ba_plan* plan);            /* asynchronous on the context stream */
int urmvo_ba_plan_download(urmvo_ba_plan* plan, double* poses, double* pts, uint8_t* inlier,
                           urmvo_ba_stats* stats);    /* synchronises the stream */
void urmvo_ba_plan_destroy(urmvo_ba_plan* plan);

/* ---- point-sharded large BA over NCCL (one process per GPU) ----
 * Every rank passes ALL cameras (replicated) and ITS OWN contiguous range of points with their
 * observations (pt indices local to the range).  Per damped trial the ranks all-reduce the reduced
 * camera system [S | b_s | b_p | chi2] and the trial cost over NCCL/NVLink; everything else is local.
 *   rank 0: urmvo_nccl_unique_id(id) -> broadcast the 128 bytes (torch.distributed, MPI, ...) ->
 *   every rank: urmvo_comm_init(ctx, rank, world, id).  Without a communicator the plan runs alone.
 *   urmvo_ba_covisibility fills the upper-triangular free-camera co-visibility of the local
 *   observations (returns Ncf); OR the matrices of all ranks and pass the result as `covis` so that
 *   every rank builds the same block structure of S (NULL: the local structure, world size 1 only).
 *   Run with urmvo_ba_plan_run (synchronous here), read back with urmvo_ba_plan_download
 *   (poses: all cameras, identical on every rank; pts / inlier: the rank's own). */
int urmvo_nccl_unique_id(uint8_t* id128);
int urmvo_comm_init(urmvo_ctx* ctx, int rank, int world, const uint8_t* id128);
int urmvo_ba_covisibility(int Nc, const uint8_t* fixed, int Np, int No, const int32_t* cam, const int32_t* pt,
                          uint8_t* upper /* Ncf*Ncf */);
int urmvo_sharded_ba_create(urmvo_ctx* ctx, urmvo_ba_plan** plan, int Nc, const double* poses,
                            const uint8_t* fixed, int Np, const double* pts, int No, const double* uv,
                            const int32_t* cam, const int32_t* pt, const double* intr, double chi2_thr,
                            int it0, int it1, const uint8_t* covis, const urmvo_ba_options* opts);
int urmvo_sharded_ba_run(urmvo_ba_plan* plan);

/* Phase times of the last run of a large / sharded plan in tile mode, from CUDA events around the first
 * trial of every enqueued batch: ms[0] linearise, ms[1] all-reduce of the reduced camera system,
 * ms[2] direct band solve, ms[3] back-substitution + trial cost + decision.  info[0] = 1 if the plan
 * runs in tile mode, info[1] = block half-bandwidth of S, info[2] = trials enqueued, info[3] = host
 * synchronisations, info[4] = doubles per all-reduce of the reduced system.  Returns URMVO_OK. */
int urmvo_ba_plan_phase_info(urmvo_ba_plan* plan, float* ms4, int32_t* info5);

/* Development aid: SM cycles spent per phase by window 0 of the BA launches since the last reset
 * (0 LIN diag, 1 LIN, 2 reduce, 3 PCG, 4 camera update, 5 BACKSUB, 6 reduce, 7 unused). */
int urmvo_debug_ba_timing(uint64_t* cycles8, int reset);
/* Same for the direct band solve of the tile mode (thread 0 of its CTA): 0 diagonal factorisation,
 * 1 panel, 2 trailing update, 3 back substitution, 4 whole kernel, 5 tail, 6 block steps. */
int urmvo_debug_lg_timing(uint64_t* cycles8, int reset);

/* ------------------------------------------------------------------ pose-only (B7-B8) */

/* Batched FrameOptimization: frame f owns observations [obs_off[f], obs_off[f+1]).
 * poses B*7 in/out (T_wc); uv No*2; Xw No*3 (world points, constants); inlier No bytes in/out
 * (the reference reads the incoming flag, src/g2o_optimization.cc:273); n_inlier B ints out
 * (= #constraints - #outliers of the last round, :319-320). rounds / its = 4 / 10. */
int urmvo_pose_only_batch(urmvo_ctx* ctx, int B, const int32_t* obs_off, double* poses,
                          const double* uv, const double* Xw, const double* intr, double chi2_thr,
                          int rounds, int its_per_round, uint8_t* inlier, int32_t* n_inlier);

/* The same with stereo edges (camera type STEREO, reference src/g2o_optimization.cc:235-258,
 * EdgeStereoSE3ProjectXYZOnlyPose): uv3 No*3 = (u_left, v_left, u_right), kind No bytes (1 = stereo edge:
 * 3-row residual, Omega = I3, Huber delta (float)sqrt(chi2_thr_stereo), threshold chi2_thr_stereo = cfg.stereo_point;
 * 0 = mono edge, u_right ignored), intr5 = fx, fy, cx, cy, bf. */
int urmvo_pose_only_batch_stereo(urmvo_ctx* ctx, int B, const int32_t* obs_off, double* poses,
                                 const double* uv3, const uint8_t* kind, const double* Xw, const double* intr5,
                                 double chi2_thr_mono, double chi2_thr_stereo, int rounds, int its_per_round,
                                 uint8_t* inlier, int32_t* n_inlier);

/* Per-constraint camera models (camera_list[mpc->id_camera], src/g2o_optimization.cc:221-224, :243-250): intr5_tab
 * and kind_model as in urmvo_local_ba_multicam. */
int urmvo_pose_only_batch_multicam(urmvo_ctx* ctx, int B, const int32_t* obs_off, double* poses,
                                   const double* uv3, const uint8_t* kind_model, const double* Xw, int n_models,
                                   const double* intr5_tab, double chi2_thr_mono, double chi2_thr_stereo,
                                   int rounds, int its_per_round, uint8_t* inlier, int32_t* n_inlier);

int urmvo_pose_plan_create(urmvo_ctx* ctx, urmvo_pose_plan** plan, int B, const int32_t* obs_off,
                           const double* poses, const double* uv, const double* Xw,
                           const double* intr, double chi2_thr, int rounds, int its_per_round,
                           const uint8_t* inlier);
int urmvo_pose_plan_run(urmvo_pose_plan* plan);
int urmvo_pose_plan_download(urmvo_pose_plan* plan, double* poses, uint8_t* inlier, int32_t* n_inlier,
                             int32_t* lm_iters /* B, total outer iterations, may be NULL */);
void urmvo_pose_plan_destroy(urmvo_pose_plan* plan);

/* ------------------------------------------------------------------ two-view RANSAC (R1-R8) */

typedef struct {
  float SH, SF;            /* best homography / fundamental scores */
  int32_t best_H, best_F;  /* hypothesis index of the best model, -1 if no score > 0 */
  float H21[9], F21[9];    /* best models, row-major */
  int32_t used_H;          /* 1: reconstructed from H, 0: from F, -1: SH+SF == 0 */
  int32_t n_good[8];       /* nGood per motion hypothesis (4 for F, 8 for H) */
  float parallax[8];       /* degrees */
  int32_t best_motion;     /* accepted motion hypothesis or -1 */
} urmvo_tv_stats;

/* EpipolarGeometry::reconstruct on host buffers.
 *  keys1 n1*2, keys2 n2*2 pixel coordinates; matches12 n1 ints (index into keys2 or -1);
 *  K 3x3 row-major; sets n_hyp*8 indices into the list of valid matches (host-generated with the
 *  reference's Random::RandomInt so that "same seeds" means the same array).
 *  Outputs: T21 4x4 row-major, P3D n1*3, triangulated n1 bytes; mask_H / mask_F: N bytes each
 *  (N = number of valid matches) or NULL.  *success = 1 if the reference would return true. */
int urmvo_two_view(urmvo_ctx* ctx, int n1, const float* keys1, int n2, const float* keys2,
                   const int32_t* matches12, const float* K, float sigma, int n_hyp,
                   const int32_t* sets, float* T21, float* P3D, uint8_t* triangulated,
                   uint8_t* mask_H, uint8_t* mask_F, urmvo_tv_stats* stats, int* success);

/* Scoring rule of the FUNDAMENTAL hypotheses (the homography rule is always the reference's):
 *  URMVO_TV_SCORE_REFERENCE  what EpipolarGeometry::_check_F computes (src/epipolar_geometry.cc:372-449): the two
 *                            squared point-to-epipolar-line distances / sigma^2, each tested against 3.841, each
 *                            passing one adding 5.991 - chi2 — the default, and the only mode with a reference to match;
 *  URMVO_TV_SCORE_SAMPSON    Sampson error (x2^T F x1)^2 / (|F x1|_12^2 + |F^T x2|_12^2) / sigma^2, ONE test against
 *                            3.841 per match, score += 5.991 - chi2 (the mode BASELINE.json's north_star names;
 *                            bit-exact against the oracle's restatement of the same rule). */
enum { URMVO_TV_SCORE_REFERENCE = 0, URMVO_TV_SCORE_SAMPSON = 1 };
int urmvo_two_view_scored(urmvo_ctx* ctx, int n1, const float* keys1, int n2, const float* keys2,
                          const int32_t* matches12, const float* K, float sigma, int n_hyp,
                          const int32_t* sets, int score_mode, float* T21, float* P3D, uint8_t* triangulated,
                          uint8_t* mask_H, uint8_t* mask_F, urmvo_tv_stats* stats, int* success);

int urmvo_tv_plan_create(urmvo_ctx* ctx, urmvo_tv_plan** plan, int n1, const float* keys1, int n2,
                         const float* keys2, const int32_t* matches12, const float* K, float sigma,
                         int n_hyp, const int32_t* sets);
/* Selects the scoring rule used by the next run_ransac (default URMVO_TV_SCORE_REFERENCE). */
int urmvo_tv_plan_set_score_mode(urmvo_tv_plan* plan, int score_mode);
/* Fit + score + arg-max of all hypotheses of both models (asynchronous). */
int urmvo_tv_plan_run_ransac(urmvo_tv_plan* plan);
/* Per-hypothesis results for parity tests: model 0 = F, 1 = H.
 * scores n_hyp floats, masks n_hyp*ceil(N/32) words, models n_hyp*9 floats (any may be NULL). */
int urmvo_tv_plan_download_hyps(urmvo_tv_plan* plan, int model, float* scores, uint32_t* masks,
                                float* models);
/* Model selection + motion recovery + triangulation; synchronises and fills the outputs. */
int urmvo_tv_plan_reconstruct(urmvo_tv_plan* plan, float* T21, float* P3D, uint8_t* triangulated,
                              uint8_t* mask_H, uint8_t* mask_F, urmvo_tv_stats* stats, int* success);
void urmvo_tv_plan_destroy(urmvo_tv_plan* plan);

/* ---- per-frame outlier rejection of the matcher -----------------------
 * Replaces the OpenCV call of reference src/point_matching.cc:53
 *     cv::findFundamentalMat(points0, points1, cv::FM_RANSAC, 3, 0.99, inliers);
 * (default maxIters 1000): same cv::RNG subset sequence, same 7-point models, same error and threshold rule,
 * same "strictly more inliers wins, then shrink the iteration budget" replay, hence the same inlier flags as
 * OpenCV.  Below 15 points OpenCV leaves the RANSAC branch and so does this call: N == 7 solves the seven
 * points directly and flags every match; 8 <= N <= 14 is LMedS (median of the float errors, sigma = 2.5 * 1.4826
 * * (1 + 5 / (N - 7)) * sqrt(median)).  N == 7 and N == 14 reproduce the real OpenCV bit for bit; for 8 <= N <= 13
 * the median is the rounding noise of an exactly-fitted sample point (~1e-27), so the result is a valid LMedS
 * answer but no two implementations (two OpenCV builds included) agree on it.  N < 7: URMVO_ERR_UNSUPPORTED
 * (OpenCV returns an empty matrix and leaves the mask untouched). */
typedef struct {
  int32_t found;      /* 1: a model with >= 7 inliers exists (cv: non-empty F) */
  int32_t iters;      /* RANSAC / LMedS iterations the sequential loop would have run */
  int32_t n_inliers;
  int32_t n_models;   /* 7-point models produced by the evaluated iterations */
  double F[9];        /* row-major, F33 = 1 (p1^T F p0 = 0) */
} urmvo_fm_stats;

/* B independent frame pairs in one call.  off: B+1 prefix offsets into pts0/pts1 (N_b*2 floats each,
 * pixel coordinates of the matched keypoints in image 0 / image 1).  inlier: sum(N_b) bytes out. */
int urmvo_fm_ransac_batch(urmvo_ctx* ctx, int B, const int32_t* off, const float* pts0, const float* pts1,
                          double thresh, double confidence, int max_iters, uint8_t* inlier,
                          urmvo_fm_stats* stats);
/* One frame pair (the call shape of point_matching.cc:53). */
int urmvo_fm_ransac(urmvo_ctx* ctx, int N, const float* pts0, const float* pts1, double thresh,
                    double confidence, int max_iters, uint8_t* inlier, urmvo_fm_stats* stats);
/* Device-resident form: create uploads the correspondences and the host-drawn subsets, run launches
 * the solve + score kernels (asynchronous, repeatable), finish reads the inlier counts back, replays
 * OpenCV's sequential selection on the host and launches the mask kernel. */
int urmvo_fm_plan_create(urmvo_ctx* ctx, urmvo_fm_plan** plan, int B, const int32_t* off, const float* pts0,
                         const float* pts1, double thresh, double confidence, int max_iters);
int urmvo_fm_plan_run(urmvo_fm_plan* plan);
int urmvo_fm_plan_finish(urmvo_fm_plan* plan, uint8_t* inlier, urmvo_fm_stats* stats);
int urmvo_fm_plan_hypotheses(const urmvo_fm_plan* plan); /* (iteration, problem) pairs evaluated per run */
void urmvo_fm_plan_destroy(urmvo_fm_plan* plan);

/* ------------------------------------------------------------------ device-resident map
 * Mapping::LocalMapOptimization (reference src/mapping.cc:335-535) rebuilds the bundle-adjustment problem from
 * shared_ptr graphs on every keyframe and copies every pose, point and keypoint into fresh containers
 * (:353-469) before it calls LocalmapOptimization (:471).  With a urmvo_map the keyframe poses, mappoint positions
 * and observations stay in HBM across keyframes, addressed by the caller's ids (frame ids, mappoint ids); a keyframe
 * uploads only what is new, a window is selected by id lists, and the optimised values are written back into the map
 * on the device.  The library keeps the index structure (id -> slot, per-point observer lists) on the host.
 *  poses: T_wc as (qx,qy,qz,qw,px,py,pz) like Pose3d; intr = fx,fy,cx,cy (mono edges).
 *  set_*: adds new ids or overwrites existing ones.  remove_observations: unknown pairs are ignored.
 *  urmvo_map_local_ba: the window is every stored observation (kf, pt) with kf in kf_ids and pt in pt_ids, point by
 *  point in pt_ids order and in insertion order inside a point; a point with fewer than two observations in the window
 *  is left out (:463-465).  kf_fixed[i] = 1 keeps keyframe i fixed (:355-356).  Cameras are indexed in kf_ids order, so
 *  a caller that lists ids in ascending order gets the vertex order of the reference's std::map.  On return the map
 *  holds the optimised free poses and points (read them with urmvo_map_get_*), and the observations used are listed
 *  with their inlier flags (n_obs entries of obs_kf / obs_pt / inlier; max_obs = capacity of those arrays) so that
 *  the caller can erase the outliers (:474-500) with urmvo_map_remove_observations.  Same arithmetic, same results
 *  as urmvo_local_ba on the same window (tests/test_gpu_map.py). */
typedef struct urmvo_map urmvo_map;
int urmvo_map_create(urmvo_ctx* ctx, urmvo_map** map, const double* intr);
void urmvo_map_destroy(urmvo_map* map);
int urmvo_map_set_keyframes(urmvo_map* map, int n, const int32_t* ids, const double* poses);
int urmvo_map_set_points(urmvo_map* map, int n, const int32_t* ids, const double* xyz);
int urmvo_map_add_observations(urmvo_map* map, int n, const int32_t* kf_ids, const int32_t* pt_ids, const double* uv);
int urmvo_map_remove_observations(urmvo_map* map, int n, const int32_t* kf_ids, const int32_t* pt_ids);
int urmvo_map_get_keyframes(urmvo_map* map, int n, const int32_t* ids, double* poses);
int urmvo_map_get_points(urmvo_map* map, int n, const int32_t* ids, double* xyz);
int urmvo_map_local_ba(urmvo_map* map, int n_kf, const int32_t* kf_ids, const uint8_t* kf_fixed, int n_pt,
                       const int32_t* pt_ids, double chi2_thr, int it0, int it1, const urmvo_ba_options* opts,
                       int32_t max_obs, int32_t* n_obs, int32_t* obs_kf, int32_t* obs_pt, uint8_t* inlier,
                       urmvo_ba_stats* stats);

/* ------------------------------------------------------------------ SolvePnPWithCV (B9)
 * Replaces the OpenCV call of SolvePnPWithCV, reference src/g2o_optimization.cc:353-355:
 *     cv::solvePnPRansac(object_points, image_points, camera_matrix, dist_coeffs (zero), rvec, tvec,
 *                        false, 100, 20.0, 0.99, cv_inliers);
 * run for every tracked frame before FrameOptimization (src/tracking.cc:799).  Same procedure as OpenCV's: 5-point
 * subsets drawn with cv::RNG(-1), EPnP per subset, float reprojection error <= reproj_thr^2, strictly-more-inliers
 * wins, RANSACUpdateNumIters(confidence), refinement of the winner over its inliers.  All hypotheses of the budget
 * are evaluated in parallel and OpenCV's sequential bookkeeping is replayed on the host, so the result does not
 * depend on the parallel evaluation.
 *  obj N*3 floats (the reference gathers cv::Point3f, :344-346), img N*2 floats, intr = fx,fy,cx,cy.
 *  max_iters / reproj_thr / confidence = 100 / 20.0 / 0.99 in the reference (<= 0 selects these).
 *  inlier: N bytes out (1 = in cv_inliers).  stats->R, t: T_cw (the reference converts (rvec, tvec) to
 *  T_wc = [R^T, -R^T t], :357-366).  Problems with fewer than 6 points return URMVO_ERR_UNSUPPORTED (the reference
 *  itself returns 0 below 8 points, :349-350, without calling OpenCV). */
typedef struct {
  int32_t found;      /* 1: some minimal sample had more than 4 inliers */
  int32_t iters;      /* iterations OpenCV's loop runs before its adaptive budget ends */
  int32_t n_inliers;  /* inliers of the best model = cv_inliers.rows */
  int32_t n_models;   /* minimal samples among those iterations that produced a model */
  double R[9];        /* T_cw rotation, row-major, after the refinement over the inliers */
  double t[3];
} urmvo_pnp_stats;
int urmvo_pnp_ransac(urmvo_ctx* ctx, int N, const float* obj, const float* img, const double* intr, int max_iters,
                     double reproj_thr, double confidence, uint8_t* inlier, urmvo_pnp_stats* stats);
/* B frames at once: frame b owns points [off[b], off[b+1]) of obj / img / inlier; stats: B entries. */
int urmvo_pnp_ransac_batch(urmvo_ctx* ctx, int B, const int32_t* off, const float* obj, const float* img,
                           const double* intr, int max_iters, double reproj_thr, double confidence, uint8_t* inlier,
                           urmvo_pnp_stats* stats);

/* ---- batched mappoint triangulation ----------------------------------
 * Mapping::TriangulateMappoint (reference src/mapping.cc:151-205) for n_pts mappoints in one launch:
 * multi-view midpoint from the observing keyframes, Eigen::ColPivHouseholderQR rank test (1e-5).
 * obs_off: n_pts+1 offsets; per observer the index of its keyframe in poses_Rp and the keypoint (u,v);
 * poses_Rp: n_poses x 12 doubles, the keyframe pose T_wc as R (row-major) | p (Frame::GetPose()).
 * ok[l] = 1 and pts[3l..] written on success; ok[l] = 0 (< 2 observers or rank < 3) leaves pts[3l..]
 * untouched, like the reference's early return. */
int urmvo_triangulate_batch(urmvo_ctx* ctx, int n_pts, const int32_t* obs_off, const int32_t* obs_pose,
                            const double* obs_uv, int n_poses, const double* poses_Rp, const double* intr,
                            double* pts, uint8_t* ok);

#ifdef __cplusplus
}
#endif
#endif /* URMVO_B200_H_ */
