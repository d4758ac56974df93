/* oracle/oracle.h — C API of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (ur-mvo_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * PARITY: the fundamental-matrix RANSAC part (fm_oracle.cpp) is PINNED bit-for-bit against outputs
 * of the real cv2.findFundamentalMat (tests/golden/golden_fm_r01.npz).  The BA / pose-only /
 * two-view reconstruction parts are PARITY UNPINNED: the reference (be2rlab/UR-MVO) ships no tests, golden vectors
 * or fixtures for this path and its arithmetic lives in g2o / Eigen / OpenCV,
 * none of which is vendored or installable here.  This oracle is
 * a restatement of src/g2o_optimization.cc and src/epipolar_geometry.cc plus the
 * published g2o Levenberg-Marquardt semantics; it is cross-checked in tests/
 * against scipy / numpy / cv2 and against the glibc rand() known answers.
 */
#ifndef URMVO_ORACLE_H_
#define URMVO_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define URMVO_ORACLE_TRACE_MAX 64

/* One row per outer LM iteration (one g2o OptimizationAlgorithmLevenberg::solve call). */
typedef struct {
  double chi2_before;   /* activeRobustChi2 at the start of the iteration */
  double chi2_after;    /* currentChi when the iteration returns */
  double lambda_after;  /* _currentLambda when the iteration returns */
  int32_t trials;       /* damped solves tried (qmax) */
  int32_t accepted;     /* 1 if the last trial was accepted */
} urmvo_oracle_trace_row;

typedef struct {
  int32_t n_rows;              /* rows used in trace */
  int32_t iters[4];            /* outer iterations run per optimize() call (BA: 2 calls, pose-only: 4 rounds) */
  double chi2_final[4];        /* currentChi after each optimize() call */
  double lambda_final[4];
  urmvo_oracle_trace_row trace[URMVO_ORACLE_TRACE_MAX];
} urmvo_oracle_stats;

/* LocalmapOptimization (reference src/g2o_optimization.cc:20-177), mono edges only.
 * poses: Nc*7 doubles, T_wc as (qx,qy,qz,qw,px,py,pz), updated in place.
 * fixed: Nc bytes.  pts: Np*3 doubles, updated in place.
 * uv: No*2, cam/pt: dense indices into poses/pts.  intr = fx,fy,cx,cy.
 * chi2_thr = cfg.mono_point; huber delta = (double)(float)sqrt(chi2_thr).
 * it0/it1 = 10/5 in the reference.  inlier: No bytes out.  Returns 0. */
int urmvo_oracle_local_ba(int Nc, double* poses, const uint8_t* fixed, int Np, double* pts,
                          int No, const double* uv, const int32_t* cam, const int32_t* pt,
                          const double* intr, double chi2_thr, int it0, int it1,
                          uint8_t* inlier, urmvo_oracle_stats* stats);

/* FrameOptimization (reference src/g2o_optimization.cc:179-321), mono edges only, one frame.
 * pose: 7 doubles T_wc in/out.  Xw: No*3 world points, uv: No*2.
 * inlier: No bytes in/out (the reference reads the incoming flag at :273).
 * Returns No - num_outlier of the last round run. */
int urmvo_oracle_pose_only(double* pose, int No, const double* uv, const double* Xw,
                           const double* intr, double chi2_thr, int rounds, int its_per_round,
                           uint8_t* inlier, urmvo_oracle_stats* stats);

/* Batched wrapper: B independent frames, obs_offset has B+1 entries. n_threads<=1: serial. */
int urmvo_oracle_pose_only_batch(int B, const int32_t* obs_offset, double* poses,
                                 const double* uv, const double* Xw, const double* intr,
                                 double chi2_thr, int rounds, int its_per_round,
                                 uint8_t* inlier, int32_t* n_inlier, int n_threads);

/* Stereo camera (reference src/g2o_optimization.cc:96-118, 235-258): mono edges and EdgeStereoSE3ProjectXYZ
 * [OnlyPose] edges in one graph.  uv3 No*3 = (u, v, u_right), kind[o] = 1 marks a stereo edge (3 rows, Omega = I3,
 * Huber delta (float)sqrt(chi2_thr_stereo), threshold chi2_thr_stereo), intr5 = fx fy cx cy bf. */
int urmvo_oracle_local_ba_stereo(int Nc, double* poses, const uint8_t* fixed, int Np, double* pts, int No,
                                 const double* uv3, const uint8_t* kind, const int32_t* cam, const int32_t* pt,
                                 const double* intr5, double chi2_thr_mono, double chi2_thr_stereo, int it0,
                                 int it1, uint8_t* inlier, urmvo_oracle_stats* stats);
int urmvo_oracle_pose_only_stereo(double* pose, int No, const double* uv3, const uint8_t* kind, const double* Xw,
                                  const double* intr5, double chi2_thr_mono, double chi2_thr_stereo, int rounds,
                                  int its_per_round, uint8_t* inlier, urmvo_oracle_stats* stats);
/* Several camera models in one graph: the reference reads the intrinsics per constraint from
 * camera_list[mpc->id_camera] (src/g2o_optimization.cc:86-89, :106-113, :221-224, :243-250).  intr5_tab = n_models
 * rows of (fx fy cx cy bf), kind_model[o] = (1 if stereo edge) | (camera model index << 1), n_models <= 128.
 * Return -1 on a bad model index. */
int urmvo_oracle_local_ba_multicam(int Nc, double* poses, const uint8_t* fixed, int Np, double* pts, int No,
                                   const double* uv3, const uint8_t* kind_model, const int32_t* cam,
                                   const int32_t* pt, int n_models, const double* intr5_tab, double chi2_thr_mono,
                                   double chi2_thr_stereo, int it0, int it1, uint8_t* inlier,
                                   urmvo_oracle_stats* stats);
int urmvo_oracle_pose_only_multicam(double* pose, int No, const double* uv3, const uint8_t* kind_model,
                                    const double* Xw, int n_models, const double* intr5_tab, double chi2_thr_mono,
                                    double chi2_thr_stereo, int rounds, int its_per_round, uint8_t* inlier,
                                    urmvo_oracle_stats* stats);
/* unit-test hook: EdgeStereoSE3ProjectXYZ error (3), Jacobians 3x6 / 3x3; returns isDepthPositive */
int urmvo_oracle_edge_stereo(const double* Tcw, const double* X, const double* uv3, const double* intr5,
                             double* e, double* Jpose, double* Jpoint);

/* Pieces exposed for unit tests. */
/* Residual e(2), J_pose(2x6 row-major, rotation first), J_point(2x3) of EdgeSE3ProjectXYZ.
 * T_cw given as (qx,qy,qz,qw,tx,ty,tz). Returns 1 if depth > 0. */
int urmvo_oracle_edge(const double* Tcw, const double* X, const double* uv, const double* intr,
                      double* e, double* Jpose, double* Jpoint);
/* rho[3] of g2o RobustKernelHuber::robustify(e2) with the given delta. */
void urmvo_oracle_huber(double e2, double delta, double* rho);
/* T <- exp(update) * T (VertexSE3Expmap::oplusImpl). update = (omega, upsilon). */
void urmvo_oracle_se3_oplus(double* Tcw, const double* update);
/* T_wc <-> T_cw (SE3Quat(q,t).inverse(), normalising q and forcing w>=0). */
void urmvo_oracle_se3_inverse(const double* Tin, double* Tout);

/* ------------------------------------------------------------------ two-view (fp32) */

typedef struct {
  float SH, SF;            /* best scores */
  int32_t best_H, best_F;  /* hypothesis index of the best model (-1: none scored > 0) */
  float H21[9], F21[9];    /* row-major best models (de-normalised) */
  int32_t used_H;          /* 1: reconstructed from H, 0: from F, -1: SH+SF==0 */
  int32_t n_good[8];       /* nGood of each motion hypothesis tested (4 for F, 8 for H) */
  float parallax[8];
  int32_t best_motion;     /* index of the accepted motion hypothesis or -1 */
} urmvo_oracle_tv_stats;

/* EpipolarGeometry::reconstruct (reference src/epipolar_geometry.cc:18-98) with the
 * 8-point sets supplied by the caller (n_hyp*8 indices into the match list).
 * keys1/keys2: n1*2 / n2*2 pixel coordinates, matches12: n1 ints (-1 = unmatched).
 * mask_H/mask_F: N bytes (N = number of valid matches, in order) or NULL.
 * Returns 1 on success (T21 row-major 4x4, P3D n1*3, triangulated n1 bytes). */
int urmvo_oracle_two_view(int n1, const float* keys1, int n2, const float* keys2,
                          const int32_t* matches12, const float* K, float sigma, int n_hyp,
                          const int32_t* sets, float* T21, float* P3D, uint8_t* triangulated,
                          uint8_t* mask_H, uint8_t* mask_F, urmvo_oracle_tv_stats* stats);

/* Score every hypothesis (no arg-max): scores n_hyp floats, masks n_hyp*ceil(N/32) words,
 * models n_hyp*9 floats.  model: 0 = F, 1 = H.  For parity tests of the RANSAC kernel. */
int urmvo_oracle_score_all(int n1, const float* keys1, int n2, const float* keys2,
                           const int32_t* matches12, float sigma, int n_hyp, const int32_t* sets,
                           int model, float* scores, uint32_t* masks, float* models);
/* Same two entry points with the scoring rule of the fundamental model selectable: score_mode 0 = the
 * reference's symmetric point-line chi2 (src/epipolar_geometry.cc:372-449), 1 = Sampson error (the extra
 * mode BASELINE.json's north_star names; not computed by the reference). */
int urmvo_oracle_two_view_mode(int n1, const float* keys1, int n2, const float* keys2,
                               const int32_t* matches12, const float* K, float sigma, int n_hyp,
                               const int32_t* sets, int score_mode, float* T21, float* P3D,
                               uint8_t* triangulated, uint8_t* mask_H, uint8_t* mask_F,
                               urmvo_oracle_tv_stats* stats);
int urmvo_oracle_score_all_mode(int n1, const float* keys1, int n2, const float* keys2,
                                const int32_t* matches12, float sigma, int n_hyp, const int32_t* sets,
                                int model, int score_mode, float* scores, uint32_t* masks, float* models);

/* Draw n_hyp x 8 index sets exactly as reconstruct() does (:53-71) with glibc rand().
 * reseed != 0 calls srand(seed) first (Random::seed_rand). */
void urmvo_oracle_draw_sets(int N, int n_hyp, int reseed, int seed, int32_t* sets);

/* One-sided Jacobi SVD used by the oracle (fp32). A is m x n row-major (m<=16, n<=9).
 * Outputs: sigma[n] (descending), V n x n row-major (columns are right singular vectors),
 * U m x n row-major (may be NULL). */
void urmvo_oracle_svd(int m, int n, const float* A, float* sigma, float* U, float* V);

/* ---- per-frame fundamental-matrix RANSAC (reference src/point_matching.cc:44-58 ->
 * cv::findFundamentalMat(points0, points1, cv::FM_RANSAC, 3, 0.99, inliers)); fm_oracle.cpp.
 * PARITY PINNED against the real cv2.findFundamentalMat (tests/golden/golden_fm_r01.npz). */

/* N >= 15 matches, p0/p1: N*2 floats.  mask: N bytes out, F9: best model (row-major, F33 = 1),
 * stats3 = {iterations run, inliers of the best model, models scored}.
 * Returns 1 if a model was found, 0 if none, -1 if N < 15 (OpenCV's 7-point / LMedS branches). */
int urmvo_oracle_fm_ransac(int N, const float* p0, const float* p1, double thresh, double confidence,
                           int max_iters, uint8_t* mask, double* F9, int32_t* stats3);
/* The whole cv::findFundamentalMat(..., FM_RANSAC, thresh, confidence, mask) dispatch for any N: N >= 15 as above,
 * N == 7 the direct 7-point solution with every mask byte 1, 8 <= N <= 14 LMeDSPointSetRegistrator::run (pinned
 * against the real cv2 for N == 7 and N == 14; for 8..13 the median error is a rounding-noise value of an
 * exactly-fitted sample point and no two builds agree — see fm_oracle.cpp).  Returns 1 / 0 like OpenCV returns
 * a matrix / an empty one, -1 for N < 7. */
int urmvo_oracle_find_fundamental(int N, const float* p0, const float* p1, double thresh, double confidence,
                                  int max_iters, uint8_t* mask, double* F9, int32_t* stats3);
/* The 7-index subsets OpenCV's RANSAC draws for iterations 0..max_iters-1 (cv::RNG state -1,
 * collinearity re-draws included).  idx: max_iters*7.  Returns the number of subsets generated. */
int urmvo_oracle_fm_subsets(int N, const float* p0, const float* p1, int max_iters, int32_t* idx);
/* run7Point on 7 correspondences: up to 3 models (27 doubles).  Returns the number of models. */
int urmvo_oracle_fm_run7(const float* p0, const float* p1, double* F27);
/* FMEstimatorCallback::computeError: err[i] = (float)max(d1^2 s1, d2^2 s2). */
void urmvo_oracle_fm_errors(int N, const float* p0, const float* p1, const double* F, float* err);

/* ---- SolvePnPWithCV (reference src/g2o_optimization.cc:323-377): cv::solvePnPRansac(..., false, 100, 20.0, 0.99),
 * pnp_oracle.cpp.  PARITY PARTLY PINNED (statistically against cv2.solvePnPRansac, see the file header).
 * obj N*3 floats, img N*2 floats, K4 = fx fy cx cy.  R9 / t3: T_cw after the final refinement over the inliers,
 * mask: N bytes (inliers of the best RANSAC model), stats4 = {iterations run, inliers, models scored, 0},
 * counts (may be NULL): inlier count of every iteration run (-1: no model).  Returns 1 found, 0 none, -1 N < 6. */
int urmvo_oracle_pnp_ransac(int N, const float* obj, const float* img, const double* K4, int max_iters,
                            double reproj_thr, double confidence, double* R9, double* t3, uint8_t* mask,
                            int32_t* stats4, int32_t* counts);
/* The 5-index subsets of iterations 0..max_iters-1 (cv::RNG state -1, no checkSubset).  idx: max_iters*5. */
int urmvo_oracle_pnp_subsets(int N, int max_iters, int32_t* idx);
/* EPnP on the 5 correspondences idx5 (the RANSAC kernel): T_cw.  Returns 1 on success. */
int urmvo_oracle_pnp_epnp5(const float* obj, const float* img, const double* K4, const int32_t* idx5, double* R9,
                           double* t3);

/* ---- Mapping::TriangulateMappoint (reference src/mapping.cc:151-205), tri_oracle.cpp.  PARITY UNPINNED.
 * n_obs observers of one mappoint: Rp = (R row-major | p) of the keyframe pose T_wc (12 doubles each),
 * uv = keypoint position.  Returns 1 and writes X (3) on success, 0 for < 2 observers or rank < 3. */
int urmvo_oracle_triangulate(int n_obs, const double* Rp, const double* uv, const double* intr, double* X);

#ifdef __cplusplus
}
#endif
#endif
