// oracle/ba_oracle.cpp — CPU restatement of the reference's g2o BA / pose-only path.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED: g2o is not vendored by the
// reference (docker/Dockerfile:145-146 clones RainerKuemmerle/g2o HEAD) and cannot be built
// here; the semantics below restate the published g2o sources:
//   g2o/core/optimization_algorithm_levenberg.cpp  (solve, computeLambdaInit, computeScale)
//   g2o/core/sparse_optimizer.cpp                   (optimize, activeRobustChi2)
//   g2o/core/block_solver.hpp                       (buildSystem, setLambda, Schur solve)
//   g2o/core/robust_kernel_impl.cpp                 (RobustKernelHuber::robustify)
//   g2o/core/base_binary_edge.hpp                   (constructQuadraticForm)
//   g2o/types/sba/types_six_dof_expmap.{h,cpp}      (EdgeSE3ProjectXYZ[OnlyPose], VertexSE3Expmap)
//   g2o/types/slam3d/se3quat.h                      (SE3Quat exp / inverse / operator*)
// and the call sequence of the reference's src/g2o_optimization.cc (cited per function).
//
// Everything is fp64 like g2o's number_t.  Single-threaded like the reference's use of g2o.

#include <algorithm>
#include <atomic>
#include <thread>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

// ---------------------------------------------------------------- SE3 (g2o se3quat.h)

struct SE3 {
  double q[4];  // x y z w
  double t[3];
};

void quat_normalize_w(double* q) {  // SE3Quat::normalizeRotation
  if (q[3] < 0) {
    for (int i = 0; i < 4; i++) q[i] = -q[i];
  }
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= n;
}

void quat_mul(const double* a, const double* b, double* r) {  // Eigen quaternion product
  double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[3] = aw * bw - ax * bx - ay * by - az * bz;
  r[0] = aw * bx + ax * bw + ay * bz - az * by;
  r[1] = aw * by + ay * bw + az * bx - ax * bz;
  r[2] = aw * bz + az * bw + ax * by - ay * bx;
}

void quat_to_R(const double* q, double* R) {  // Eigen toRotationMatrix, row-major
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w;
  double txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

void R_to_quat(const double* m, double* q) {  // Eigen quaternion from rotation matrix
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 3 + i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
    q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
  }
}

void rot_vec(const double* R, const double* v, double* r) {
  for (int i = 0; i < 3; i++) r[i] = R[i * 3] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2];
}

SE3 se3_from(const double* q, const double* t) {  // SE3Quat(q, t)
  SE3 T;
  std::memcpy(T.q, q, sizeof(T.q));
  std::memcpy(T.t, t, sizeof(T.t));
  quat_normalize_w(T.q);
  return T;
}

SE3 se3_inverse(const SE3& T) {  // SE3Quat::inverse
  SE3 r;
  r.q[0] = -T.q[0]; r.q[1] = -T.q[1]; r.q[2] = -T.q[2]; r.q[3] = T.q[3];
  double R[9], nt[3] = {-T.t[0], -T.t[1], -T.t[2]};
  quat_to_R(r.q, R);
  rot_vec(R, nt, r.t);
  return r;
}

SE3 se3_exp(const double* u) {  // SE3Quat::exp, update = (omega, upsilon)
  const double* om = u;
  const double* up = u + 3;
  double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double O2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += O[i * 3 + k] * O[k * 3 + j];
      O2[i * 3 + j] = s;
    }
  double a, b, c, d;  // R = I + a*O + b*O2 ; V = I + c*O + d*O2
  if (theta < 0.00001) {
    a = 1.0; b = 0.5; c = 0.5; d = 1.0 / 6.0;
  } else {
    a = std::sin(theta) / theta;
    b = (1 - std::cos(theta)) / (theta * theta);
    c = b;
    d = (theta - std::sin(theta)) / (theta * theta * theta);
  }
  double R[9], V[9];
  for (int i = 0; i < 9; i++) {
    double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = I + a * O[i] + b * O2[i];
    V[i] = I + c * O[i] + d * O2[i];
  }
  SE3 T;
  R_to_quat(R, T.q);
  quat_normalize_w(T.q);
  rot_vec(V, up, T.t);
  return T;
}

SE3 se3_mul(const SE3& A, const SE3& B) {  // SE3Quat::operator*
  SE3 r;
  quat_mul(A.q, B.q, r.q);
  double R[9], Rt[3];
  quat_to_R(A.q, R);
  rot_vec(R, B.t, Rt);
  for (int i = 0; i < 3; i++) r.t[i] = A.t[i] + Rt[i];
  quat_normalize_w(r.q);
  return r;
}

// ------------------------------------------------ edge (types_six_dof_expmap.cpp)

struct Intr { double fx, fy, cx, cy; };

// Camera-frame point. SE3Quat::map = r*xyz + t.
inline void map_point(const double* R, const double* t, const double* X, double* pc) {
  rot_vec(R, X, pc);
  pc[0] += t[0]; pc[1] += t[1]; pc[2] += t[2];
}

// EdgeSE3ProjectXYZ::computeError: e = obs - cam_project(T.map(X))
inline void edge_error(const double* pc, const double* uv, const Intr& K, double* e) {
  e[0] = uv[0] - (pc[0] / pc[2] * K.fx + K.cx);
  e[1] = uv[1] - (pc[1] / pc[2] * K.fy + K.cy);
}

// EdgeSE3ProjectXYZ::linearizeOplus. Jp: 2x6 (rotation first), Jx: 2x3 (may be null).
inline void edge_jac(const double* R, const double* pc, const Intr& K, double* Jp, double* Jx) {
  double x = pc[0], y = pc[1], z = pc[2], z2 = z * z;
  Jp[0] = x * y / z2 * K.fx;
  Jp[1] = -(1 + (x * x / z2)) * K.fx;
  Jp[2] = y / z * K.fx;
  Jp[3] = -1. / z * K.fx;
  Jp[4] = 0;
  Jp[5] = x / z2 * K.fx;
  Jp[6] = (1 + y * y / z2) * K.fy;
  Jp[7] = -x * y / z2 * K.fy;
  Jp[8] = -x / z * K.fy;
  Jp[9] = 0;
  Jp[10] = -1. / z * K.fy;
  Jp[11] = y / z2 * K.fy;
  if (Jx) {
    double tmp[6] = {K.fx, 0, -x / z * K.fx, 0, K.fy, -y / z * K.fy};
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += tmp[r * 3 + k] * R[k * 3 + c];
        Jx[r * 3 + c] = -1. / z * s;
      }
  }
}

// EdgeStereoSE3ProjectXYZ (types_six_dof_expmap.cpp; used at src/g2o_optimization.cc:96-118, 235-258):
// measurement (u_left, v_left, u_right), cam_project = (x/z fx + cx, y/z fy + cy, u - bf/z).
// The generalised edge has THREE rows; a mono edge (kind 0) leaves the third row zero, so every
// sum below adds exact zeros for it and a mono-only problem gives the bits of the 2-row code.
inline void edge_error3(const double* pc, const double* uv3, int kind, const Intr& K, double bf, double* e) {
  edge_error(pc, uv3, K, e);
  if (kind) {
    const double invz = 1.0 / pc[2];
    const double ul = pc[0] * invz * K.fx + K.cx;
    e[0] = uv3[0] - ul;
    e[1] = uv3[1] - (pc[1] * invz * K.fy + K.cy);
    e[2] = uv3[2] - (ul - bf * invz);
  } else {
    e[2] = 0.0;
  }
}

// Jp: 3x6 (rotation first), Jx: 3x3 (may be null).
inline void edge_jac3(const double* R, const double* pc, int kind, const Intr& K, double bf, double* Jp, double* Jx) {
  edge_jac(R, pc, K, Jp, Jx);
  const double x = pc[0], y = pc[1], z = pc[2], z2 = z * z;
  if (!kind) {
    for (int a = 0; a < 6; a++) Jp[12 + a] = 0.0;
    if (Jx) for (int a = 0; a < 3; a++) Jx[6 + a] = 0.0;
    return;
  }
  Jp[12] = Jp[0] - bf * y / z2;
  Jp[13] = Jp[1] + bf * x / z2;
  Jp[14] = Jp[2];
  Jp[15] = Jp[3];
  Jp[16] = 0;
  Jp[17] = Jp[5] - bf / z2;
  if (Jx) {
    for (int c = 0; c < 3; c++) {
      Jx[c] = -K.fx * R[c] / z + K.fx * x * R[6 + c] / z2;
      Jx[3 + c] = -K.fy * R[3 + c] / z + K.fy * y * R[6 + c] / z2;
      Jx[6 + c] = Jx[c] - bf * R[6 + c] / z2;
    }
  }
}

// RobustKernelHuber::robustify
inline void huber(double e2, double delta, double* rho) {
  double dsqr = delta * delta;
  if (e2 <= dsqr) {
    rho[0] = e2; rho[1] = 1.; rho[2] = 0.;
  } else {
    double sqrte = std::sqrt(e2);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e2;
  }
}

// ---------------------------------------------------- small dense helpers

// 3x3 symmetric inverse via cofactors (Eigen fixed-size inverse), in/out full row-major.
inline bool inv3(const double* a, double* r) {
  double c00 = a[4] * a[8] - a[5] * a[7];
  double c01 = a[5] * a[6] - a[3] * a[8];
  double c02 = a[3] * a[7] - a[4] * a[6];
  double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  double id = 1.0 / det;
  r[0] = c00 * id;
  r[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  r[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  r[3] = c01 * id;
  r[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  r[6] = c02 * id;
  r[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  return std::isfinite(id);
}

// Skyline (envelope) Cholesky of a symmetric positive definite matrix stored dense-lower in a
// row-pointer layout: row i holds columns first[i]..i.  Exact direct solve, the role of
// LinearSolverEigen (sparse Cholesky) in the reference (src/g2o_optimization.cc:27-35).
struct Skyline {
  int n = 0;
  std::vector<int> first;
  std::vector<size_t> rowptr;  // offset of column first[i] of row i
  std::vector<double> a;
  void init(int n_, const std::vector<int>& first_) {
    n = n_;
    first = first_;
    rowptr.assign(n + 1, 0);
    for (int i = 0; i < n; i++) rowptr[i + 1] = rowptr[i] + (size_t)(i - first[i] + 1);
    a.assign(rowptr[n], 0.0);
  }
  inline double& at(int i, int j) { return a[rowptr[i] + (size_t)(j - first[i])]; }  // j<=i, j>=first[i]
  void zero() { std::fill(a.begin(), a.end(), 0.0); }
  bool factor() {
    for (int i = 0; i < n; i++) {
      double* Li = &a[rowptr[i]] - first[i];
      for (int j = first[i]; j <= i; j++) {
        double* Lj = &a[rowptr[j]] - first[j];
        double s = Li[j];
        int k0 = std::max(first[i], first[j]);
        for (int k = k0; k < j; k++) s -= Li[k] * Lj[k];
        if (j < i) {
          Li[j] = s / Lj[j];
        } else {
          if (!(s > 0.0)) return false;
          Li[i] = std::sqrt(s);
        }
      }
    }
    return true;
  }
  void solve(double* x) const {  // in place, x = A^-1 x
    for (int i = 0; i < n; i++) {
      const double* Li = &a[rowptr[i]] - first[i];
      double s = x[i];
      for (int k = first[i]; k < i; k++) s -= Li[k] * x[k];
      x[i] = s / Li[i];
    }
    for (int i = n - 1; i >= 0; i--) {
      const double* Li = &a[rowptr[i]] - first[i];
      x[i] /= Li[i];
      double xi = x[i];
      for (int k = first[i]; k < i; k++) x[k] -= Li[k] * xi;
    }
  }
};

enum SolveResult { kOK = 0, kTerminate = 1 };

// ------------------------------------------------------------------ local BA

struct BAProblem {
  int Nc, Np, No;
  Intr K;
  // several camera models in one graph (camera_list[mpc->id_camera], src/g2o_optimization.cc:86-89): per-edge index
  // into (Ks, bfs); empty = every edge uses (K, bf)
  std::vector<Intr> Ks;
  std::vector<double> bfs;
  std::vector<uint8_t> model;
  const Intr& Kof(int o) const { return model.empty() ? K : Ks[model[o]]; }
  double bfof(int o) const { return model.empty() ? bf : bfs[model[o]]; }
  std::vector<SE3> cams;         // T_cw
  std::vector<uint8_t> fixed;
  std::vector<int> cam_free_idx;  // dense index among non-fixed cams or -1
  int Ncf = 0;
  std::vector<double> pts;
  std::vector<double> uv3;       // (u, v, u_right) per edge; u_right unused for a mono edge
  std::vector<uint8_t> kind;     // 0 mono (EdgeSE3ProjectXYZ), 1 stereo (EdgeStereoSE3ProjectXYZ)
  double bf = 0;
  const int32_t* ocam;
  const int32_t* opt;
  std::vector<uint8_t> level;  // 0 active, 1 outlier
  bool robust = true;
  double delta = 0, delta_s = 0;  // Huber delta of the mono / stereo edges
  // point-major CSR over observations (observation order inside a point = input order)
  std::vector<int> pt_start, pt_obs;
  // cached per-edge error (g2o Edge::_error), survives pop() — the stale-error quirk
  std::vector<double> err;
  // LM state (OptimizationAlgorithmLevenberg members)
  double lambda = 0, ni = 2;
  // linear system
  std::vector<double> Hpp, bp, Hll, bl, W;  // W per obs 6x3
  std::vector<double> x;                    // [6*Ncf + 3*Np]
  Skyline S;
  std::vector<double> Rcache;  // 9 per cam
};

void ba_refresh_R(BAProblem& P) {
  P.Rcache.resize((size_t)P.Nc * 9);
  for (int c = 0; c < P.Nc; c++) quat_to_R(P.cams[c].q, &P.Rcache[(size_t)c * 9]);
}

// SparseOptimizer::computeActiveErrors + activeRobustChi2
double ba_compute_errors(BAProblem& P) {
  ba_refresh_R(P);
  double chi = 0;
  for (int o = 0; o < P.No; o++) {
    if (P.level[o]) continue;
    double pc[3];
    int c = P.ocam[o];
    map_point(&P.Rcache[(size_t)c * 9], P.cams[c].t, &P.pts[(size_t)P.opt[o] * 3], pc);
    double* e = &P.err[(size_t)o * 3];
    edge_error3(pc, &P.uv3[(size_t)o * 3], P.kind[o], P.Kof(o), P.bfof(o), e);
    double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    if (P.robust) {
      double rho[3];
      huber(e2, P.kind[o] ? P.delta_s : P.delta, rho);
      chi += rho[0];
    } else {
      chi += e2;
    }
  }
  return chi;
}

// BlockSolver::buildSystem: linearizeOplus + constructQuadraticForm per active edge.
void ba_build_system(BAProblem& P) {
  std::fill(P.Hpp.begin(), P.Hpp.end(), 0.0);
  std::fill(P.bp.begin(), P.bp.end(), 0.0);
  std::fill(P.Hll.begin(), P.Hll.end(), 0.0);
  std::fill(P.bl.begin(), P.bl.end(), 0.0);
  for (int o = 0; o < P.No; o++) {
    if (P.level[o]) continue;
    int c = P.ocam[o], l = P.opt[o];
    const double* R = &P.Rcache[(size_t)c * 9];
    double pc[3], Jp[18], Jx[9];
    map_point(R, P.cams[c].t, &P.pts[(size_t)l * 3], pc);
    edge_jac3(R, pc, P.kind[o], P.Kof(o), P.bfof(o), Jp, Jx);
    const double* e = &P.err[(size_t)o * 3];
    double w = 1.0;
    if (P.robust) {
      double rho[3];
      huber(e[0] * e[0] + e[1] * e[1] + e[2] * e[2], P.kind[o] ? P.delta_s : P.delta, rho);
      w = rho[1];
    }
    double r0 = -w * e[0], r1 = -w * e[1], r2 = -w * e[2];  // omega_r = -rho1 * Omega * e
    // point block (vertex 0, "from")
    double* Hl = &P.Hll[(size_t)l * 9];
    double* b_l = &P.bl[(size_t)l * 3];
    for (int a = 0; a < 3; a++) {
      b_l[a] += Jx[a] * r0 + Jx[3 + a] * r1 + Jx[6 + a] * r2;
      for (int b = 0; b < 3; b++) Hl[a * 3 + b] += w * (Jx[a] * Jx[b] + Jx[3 + a] * Jx[3 + b] + Jx[6 + a] * Jx[6 + b]);
    }
    int cf = P.cam_free_idx[c];
    if (cf >= 0) {
      double* Hp = &P.Hpp[(size_t)cf * 36];
      double* b_p = &P.bp[(size_t)cf * 6];
      for (int a = 0; a < 6; a++) {
        b_p[a] += Jp[a] * r0 + Jp[6 + a] * r1 + Jp[12 + a] * r2;
        for (int b = 0; b < 6; b++) Hp[a * 6 + b] += w * (Jp[a] * Jp[b] + Jp[6 + a] * Jp[6 + b] + Jp[12 + a] * Jp[12 + b]);
      }
      double* Wo = &P.W[(size_t)o * 18];  // Hpl block (pose rows, point cols) = Jp^T w Jx
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 3; b++) Wo[a * 3 + b] = w * (Jp[a] * Jx[b] + Jp[6 + a] * Jx[3 + b] + Jp[12 + a] * Jx[6 + b]);
    }
  }
}

// OptimizationAlgorithmLevenberg::computeLambdaInit: tau * max |H_jj| over all active vertices.
double ba_lambda_init(const BAProblem& P, const std::vector<uint8_t>& pt_active,
                      const std::vector<uint8_t>& cam_active) {
  double m = 0;
  for (int c = 0; c < P.Nc; c++) {
    int cf = P.cam_free_idx[c];
    if (cf < 0 || !cam_active[c]) continue;
    for (int j = 0; j < 6; j++) m = std::max(std::fabs(P.Hpp[(size_t)cf * 36 + j * 7]), m);
  }
  for (int l = 0; l < P.Np; l++) {
    if (!pt_active[l]) continue;
    for (int j = 0; j < 3; j++) m = std::max(std::fabs(P.Hll[(size_t)l * 9 + j * 4]), m);
  }
  return 1e-5 * m;
}

// BlockSolver::solve with Schur complement, lambda added to every diagonal block.
bool ba_solve(BAProblem& P, double lambda) {
  const int np6 = 6 * P.Ncf;
  P.S.zero();
  std::vector<double> bs(P.bp);
  for (int cf = 0; cf < P.Ncf; cf++)
    for (int a = 0; a < 6; a++)
      for (int b = 0; b <= a; b++) {
        double v = P.Hpp[(size_t)cf * 36 + a * 6 + b];
        if (a == b) v += lambda;
        P.S.at(cf * 6 + a, cf * 6 + b) = v;
      }
  std::vector<double> Dinv((size_t)P.Np * 9);
  double Y[18];
  for (int l = 0; l < P.Np; l++) {
    double D[9];
    std::memcpy(D, &P.Hll[(size_t)l * 9], sizeof(D));
    D[0] += lambda; D[4] += lambda; D[8] += lambda;
    double* Di = &Dinv[(size_t)l * 9];
    inv3(D, Di);
    const double* b_l = &P.bl[(size_t)l * 3];
    for (int s = P.pt_start[l]; s < P.pt_start[l + 1]; s++) {
      int oi = P.pt_obs[s];
      if (P.level[oi]) continue;
      int ci = P.cam_free_idx[P.ocam[oi]];
      if (ci < 0) continue;
      const double* Wi = &P.W[(size_t)oi * 18];
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 3; b++)
          Y[a * 3 + b] = Wi[a * 3] * Di[b] + Wi[a * 3 + 1] * Di[3 + b] + Wi[a * 3 + 2] * Di[6 + b];
      for (int a = 0; a < 6; a++)
        bs[(size_t)ci * 6 + a] -= Y[a * 3] * b_l[0] + Y[a * 3 + 1] * b_l[1] + Y[a * 3 + 2] * b_l[2];
      for (int s2 = P.pt_start[l]; s2 < P.pt_start[l + 1]; s2++) {
        int oj = P.pt_obs[s2];
        if (P.level[oj]) continue;
        int cj = P.cam_free_idx[P.ocam[oj]];
        if (cj < 0 || cj > ci) continue;  // lower triangle (ci >= cj)
        const double* Wj = &P.W[(size_t)oj * 18];
        for (int a = 0; a < 6; a++)
          for (int b = 0; b < 6; b++) {
            if (ci == cj && b > a) continue;
            P.S.at(ci * 6 + a, cj * 6 + b) -=
                Y[a * 3] * Wj[b * 3] + Y[a * 3 + 1] * Wj[b * 3 + 1] + Y[a * 3 + 2] * Wj[b * 3 + 2];
          }
      }
    }
  }
  if (np6 > 0) {
    if (!P.S.factor()) return false;
    P.S.solve(bs.data());
  }
  std::copy(bs.begin(), bs.end(), P.x.begin());
  // landmarks: x_l = Dinv (b_l - W^T x_p)
  for (int l = 0; l < P.Np; l++) {
    double cp[3] = {P.bl[(size_t)l * 3], P.bl[(size_t)l * 3 + 1], P.bl[(size_t)l * 3 + 2]};
    for (int s = P.pt_start[l]; s < P.pt_start[l + 1]; s++) {
      int oi = P.pt_obs[s];
      if (P.level[oi]) continue;
      int ci = P.cam_free_idx[P.ocam[oi]];
      if (ci < 0) continue;
      const double* Wi = &P.W[(size_t)oi * 18];
      const double* xp = &P.x[(size_t)ci * 6];
      for (int b = 0; b < 3; b++)
        for (int a = 0; a < 6; a++) cp[b] -= Wi[a * 3 + b] * xp[a];
    }
    const double* Di = &Dinv[(size_t)l * 9];
    double* xl = &P.x[(size_t)np6 + (size_t)l * 3];
    for (int a = 0; a < 3; a++) xl[a] = Di[a * 3] * cp[0] + Di[a * 3 + 1] * cp[1] + Di[a * 3 + 2] * cp[2];
  }
  return true;
}

// SparseOptimizer::update: oplus on every active vertex.
void ba_update(BAProblem& P, const std::vector<uint8_t>& pt_active) {
  for (int c = 0; c < P.Nc; c++) {
    int cf = P.cam_free_idx[c];
    if (cf < 0) continue;
    P.cams[c] = se3_mul(se3_exp(&P.x[(size_t)cf * 6]), P.cams[c]);
  }
  const int np6 = 6 * P.Ncf;
  for (int l = 0; l < P.Np; l++) {
    if (!pt_active[l]) continue;
    for (int a = 0; a < 3; a++) P.pts[(size_t)l * 3 + a] += P.x[(size_t)np6 + (size_t)l * 3 + a];
  }
}

struct LMTrace {
  urmvo_oracle_stats* st;
  void row(double before, double after, double lambda, int trials, int accepted) {
    if (!st || st->n_rows >= URMVO_ORACLE_TRACE_MAX) return;
    urmvo_oracle_trace_row& r = st->trace[st->n_rows++];
    r.chi2_before = before; r.chi2_after = after; r.lambda_after = lambda;
    r.trials = trials; r.accepted = accepted;
  }
};

// SparseOptimizer::initializeOptimization(level 0) + optimize(n)
// Returns the number of outer iterations run; *chi_out = currentChi of the last solve().
int ba_optimize(BAProblem& P, int n_iter, urmvo_oracle_stats* st, double* chi_out) {
  // active vertices: those with at least one level-0 edge
  std::vector<uint8_t> pt_active(P.Np, 0), cam_active(P.Nc, 0);
  for (int o = 0; o < P.No; o++)
    if (!P.level[o]) { pt_active[P.opt[o]] = 1; cam_active[P.ocam[o]] = 1; }
  // Schur structure: envelope over co-visible free cameras through active edges
  {
    std::vector<int> first_blk(P.Ncf);
    for (int i = 0; i < P.Ncf; i++) first_blk[i] = i;
    for (int l = 0; l < P.Np; l++) {
      int mn = P.Ncf;
      for (int s = P.pt_start[l]; s < P.pt_start[l + 1]; s++) {
        int o = P.pt_obs[s];
        if (P.level[o]) continue;
        int cf = P.cam_free_idx[P.ocam[o]];
        if (cf >= 0) mn = std::min(mn, cf);
      }
      for (int s = P.pt_start[l]; s < P.pt_start[l + 1]; s++) {
        int o = P.pt_obs[s];
        if (P.level[o]) continue;
        int cf = P.cam_free_idx[P.ocam[o]];
        if (cf >= 0) first_blk[cf] = std::min(first_blk[cf], mn);
      }
    }
    std::vector<int> first(6 * P.Ncf);
    for (int i = 0; i < 6 * P.Ncf; i++) first[i] = first_blk[i / 6] * 6;
    P.S.init(6 * P.Ncf, first);
  }
  LMTrace tr{st};
  double currentChi = 0;
  int it = 0;
  for (; it < n_iter; it++) {
    // ---- OptimizationAlgorithmLevenberg::solve(it)
    currentChi = ba_compute_errors(P);
    double chi_before = currentChi;
    double tempChi = currentChi;
    ba_build_system(P);
    if (it == 0) {
      P.lambda = ba_lambda_init(P, pt_active, cam_active);
      P.ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    int accepted = 0;
    bool lambda_bad = false;
    do {
      std::vector<SE3> cams_backup = P.cams;  // push()
      std::vector<double> pts_backup = P.pts;
      bool ok2 = ba_solve(P, P.lambda);
      if (ok2) ba_update(P, pt_active);
      tempChi = ba_compute_errors(P);
      if (!ok2) tempChi = DBL_MAX;
      rho = currentChi - tempChi;
      double scale = 0;  // computeScale
      {
        const int np6 = 6 * P.Ncf;
        for (int j = 0; j < np6; j++) scale += P.x[j] * (P.lambda * P.x[j] + P.bp[j]);
        for (int l = 0; l < P.Np; l++) {
          if (!pt_active[l]) continue;
          for (int a = 0; a < 3; a++) {
            double xj = P.x[(size_t)np6 + (size_t)l * 3 + a];
            scale += xj * (P.lambda * xj + P.bl[(size_t)l * 3 + a]);
          }
        }
      }
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        double scaleFactor = std::max(1. / 3., alpha);
        P.lambda *= scaleFactor;
        P.ni = 2;
        currentChi = tempChi;
        accepted = 1;  // discardTop()
      } else {
        P.lambda *= P.ni;
        P.ni *= 2;
        P.cams = cams_backup;  // pop(); cached errors are NOT recomputed (stale-error quirk)
        P.pts = pts_backup;
        accepted = 0;
        if (!std::isfinite(P.lambda)) { lambda_bad = true; break; }
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    tr.row(chi_before, currentChi, P.lambda, qmax, accepted);
    if (qmax == 10 || rho == 0 || lambda_bad || !std::isfinite(P.lambda)) { it++; break; }
  }
  if (chi_out) *chi_out = currentChi;
  return it;
}

// uv_stride 2: mono measurements (u, v); 3: (u, v, u_right) with `kind` selecting the edge type.
void ba_setup(BAProblem& P, int Nc, const double* poses, const uint8_t* fixed, int Np,
              const double* pts, int No, const double* uv, int uv_stride, const uint8_t* kind,
              const int32_t* cam, const int32_t* pt, const double* intr, double bf, double chi2_thr,
              double chi2_thr_stereo, int n_models = 0) {
  P.Nc = Nc; P.Np = Np; P.No = No;
  P.K = Intr{intr[0], intr[1], intr[2], intr[3]};
  P.bf = bf;
  P.uv3.assign((size_t)No * 3, 0.0);
  P.kind.assign(No, 0);
  if (n_models > 0) {  // intr = n_models rows of (fx fy cx cy bf); kind[o] = stereo bit | model << 1
    for (int m = 0; m < n_models; m++) {
      P.Ks.push_back(Intr{intr[m * 5], intr[m * 5 + 1], intr[m * 5 + 2], intr[m * 5 + 3]});
      P.bfs.push_back(intr[m * 5 + 4]);
    }
    P.model.assign(No, 0);
  }
  for (int o = 0; o < No; o++) {
    P.uv3[(size_t)o * 3] = uv[(size_t)o * uv_stride];
    P.uv3[(size_t)o * 3 + 1] = uv[(size_t)o * uv_stride + 1];
    const bool st = uv_stride == 3 && kind && (n_models > 0 ? (kind[o] & 1) : kind[o]);
    if (st) { P.uv3[(size_t)o * 3 + 2] = uv[(size_t)o * 3 + 2]; P.kind[o] = 1; }
    if (n_models > 0) P.model[o] = kind ? (uint8_t)(kind[o] >> 1) : 0;
  }
  P.cams.resize(Nc);
  P.fixed.assign(fixed, fixed + Nc);
  P.cam_free_idx.assign(Nc, -1);
  P.Ncf = 0;
  for (int c = 0; c < Nc; c++) {
    // src/g2o_optimization.cc:45 — setEstimate(SE3Quat(q, p).inverse())
    P.cams[c] = se3_inverse(se3_from(poses + (size_t)c * 7, poses + (size_t)c * 7 + 4));
    if (!fixed[c]) P.cam_free_idx[c] = P.Ncf++;
  }
  P.pts.assign(pts, pts + (size_t)Np * 3);
  P.ocam = cam; P.opt = pt;
  P.level.assign(No, 0);
  P.err.assign((size_t)No * 3, 0.0);
  P.delta = (double)(float)std::sqrt(chi2_thr);           // :71 const float thHuberMonoPoint = sqrt(cfg.mono_point)
  P.delta_s = (double)(float)std::sqrt(chi2_thr_stereo);  // :72 const float thHuberStereoPoint = sqrt(cfg.stereo_point)
  P.pt_start.assign(Np + 1, 0);
  for (int o = 0; o < No; o++) P.pt_start[pt[o] + 1]++;
  for (int l = 0; l < Np; l++) P.pt_start[l + 1] += P.pt_start[l];
  P.pt_obs.resize(No);
  {
    std::vector<int> fill(P.pt_start.begin(), P.pt_start.end() - 1);
    for (int o = 0; o < No; o++) P.pt_obs[fill[pt[o]]++] = o;
  }
  P.Hpp.assign((size_t)P.Ncf * 36, 0.0);
  P.bp.assign((size_t)P.Ncf * 6, 0.0);
  P.Hll.assign((size_t)Np * 9, 0.0);
  P.bl.assign((size_t)Np * 3, 0.0);
  P.W.assign((size_t)No * 18, 0.0);
  P.x.assign((size_t)P.Ncf * 6 + (size_t)Np * 3, 0.0);
}

inline bool ba_depth_positive(const BAProblem& P, int o) {  // isDepthPositive at the CURRENT estimate
  double R[9], pc[3];
  int c = P.ocam[o];
  quat_to_R(P.cams[c].q, R);
  map_point(R, P.cams[c].t, &P.pts[(size_t)P.opt[o] * 3], pc);
  return pc[2] > 0.0;
}

}  // namespace

static int local_ba_impl(int Nc, double* poses, const uint8_t* fixed, int Np, double* pts, int No,
                         const double* uv, int uv_stride, const uint8_t* kind, const int32_t* cam,
                         const int32_t* pt, const double* intr, double bf, double chi2_thr,
                         double chi2_thr_stereo, int it0, int it1, uint8_t* inlier,
                         urmvo_oracle_stats* stats, int n_models = 0);

extern "C" int urmvo_oracle_local_ba(int Nc, double* poses, const uint8_t* fixed, int Np,
                                     double* pts, int No, const double* uv, const int32_t* cam,
                                     const int32_t* pt, const double* intr, double chi2_thr,
                                     int it0, int it1, uint8_t* inlier,
                                     urmvo_oracle_stats* stats) {
  return local_ba_impl(Nc, poses, fixed, Np, pts, No, uv, 2, nullptr, cam, pt, intr, 0.0, chi2_thr, chi2_thr, it0, it1,
                       inlier, stats);
}

// Mono + stereo edges in one graph (camera type STEREO, src/g2o_optimization.cc:96-118): uv3 = (u, v, u_right),
// kind[o] = 1 for a stereo edge, intr5 = fx fy cx cy bf, thresholds cfg.mono_point / cfg.stereo_point.
extern "C" int urmvo_oracle_local_ba_stereo(int Nc, double* poses, const uint8_t* fixed, int Np,
                                            double* pts, int No, const double* uv3, const uint8_t* kind,
                                            const int32_t* cam, const int32_t* pt, const double* intr5,
                                            double chi2_thr_mono, double chi2_thr_stereo, int it0, int it1,
                                            uint8_t* inlier, urmvo_oracle_stats* stats) {
  return local_ba_impl(Nc, poses, fixed, Np, pts, No, uv3, 3, kind, cam, pt, intr5, intr5[4], chi2_thr_mono,
                       chi2_thr_stereo, it0, it1, inlier, stats);
}

// Several camera models in one graph: the reference reads fx, fy, cx, cy (and BF) per constraint from
// camera_list[mpc->id_camera] (src/g2o_optimization.cc:86-89, :106-113).  intr5_tab = n_models rows of
// (fx fy cx cy bf); kind_model[o] = (1 if stereo edge) | (camera model index << 1).
extern "C" int urmvo_oracle_local_ba_multicam(int Nc, double* poses, const uint8_t* fixed, int Np, double* pts,
                                              int No, const double* uv3, const uint8_t* kind_model,
                                              const int32_t* cam, const int32_t* pt, int n_models,
                                              const double* intr5_tab, double chi2_thr_mono,
                                              double chi2_thr_stereo, int it0, int it1, uint8_t* inlier,
                                              urmvo_oracle_stats* stats) {
  if (n_models <= 0 || n_models > 128) return -1;
  for (int o = 0; o < No; o++)
    if ((kind_model[o] >> 1) >= n_models) return -1;
  return local_ba_impl(Nc, poses, fixed, Np, pts, No, uv3, 3, kind_model, cam, pt, intr5_tab, intr5_tab[4],
                       chi2_thr_mono, chi2_thr_stereo, it0, it1, inlier, stats, n_models);
}

static int local_ba_impl(int Nc, double* poses, const uint8_t* fixed, int Np, double* pts, int No,
                         const double* uv, int uv_stride, const uint8_t* kind, const int32_t* cam,
                         const int32_t* pt, const double* intr, double bf, double chi2_thr,
                         double chi2_thr_stereo, int it0, int it1, uint8_t* inlier,
                         urmvo_oracle_stats* stats, int n_models) {
  if (stats) std::memset(stats, 0, sizeof(*stats));
  BAProblem P;
  ba_setup(P, Nc, poses, fixed, Np, pts, No, uv, uv_stride, kind, cam, pt, intr, bf, chi2_thr, chi2_thr_stereo, n_models);
  // :125-126 initializeOptimization(); optimize(10)  — Huber on every edge
  P.robust = true;
  double chi = 0;
  int n0 = ba_optimize(P, it0, stats, &chi);
  if (stats) { stats->iters[0] = n0; stats->chi2_final[0] = chi; stats->lambda_final[0] = P.lambda; }
  // :129-135 chi2() reads the cached error; isDepthPositive() re-maps with the current estimate
  auto cached_chi2 = [&](int o) {
    const double* e = &P.err[(size_t)o * 3];
    return e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  };
  for (int o = 0; o < No; o++) {  // :129-143 mono edges against cfg.mono_point, stereo edges against cfg.stereo_point
    if (cached_chi2(o) > (P.kind[o] ? chi2_thr_stereo : chi2_thr) || !ba_depth_positive(P, o)) P.level[o] = 1;
  }
  P.robust = false;  // e->setRobustKernel(0)
  // :146-147 initializeOptimization(0); optimize(5)
  int n1 = ba_optimize(P, it1, stats, &chi);
  if (stats) { stats->iters[1] = n1; stats->chi2_final[1] = chi; stats->lambda_final[1] = P.lambda; }
  // :150-154 — level-1 edges keep the error cached by the first optimize()
  for (int o = 0; o < No; o++)
    inlier[o] = (cached_chi2(o) <= (P.kind[o] ? chi2_thr_stereo : chi2_thr) && ba_depth_positive(P, o)) ? 1 : 0;
  // :164-176 write back T_wc = estimate().inverse(), points
  for (int c = 0; c < Nc; c++) {
    SE3 Twc = se3_inverse(P.cams[c]);
    std::memcpy(poses + (size_t)c * 7, Twc.q, 4 * sizeof(double));
    std::memcpy(poses + (size_t)c * 7 + 4, Twc.t, 3 * sizeof(double));
  }
  std::memcpy(pts, P.pts.data(), (size_t)Np * 3 * sizeof(double));
  return 0;
}

// ------------------------------------------------------------------ pose only

namespace {

struct PoseProblem {
  int No;
  Intr K;
  std::vector<Intr> Ks;  // per-edge camera models, see BAProblem
  std::vector<double> bfs;
  std::vector<uint8_t> model;
  const Intr& Kof(int o) const { return model.empty() ? K : Ks[model[o]]; }
  double bfof(int o) const { return model.empty() ? bf : bfs[model[o]]; }
  SE3 T;  // T_cw
  std::vector<double> uv3;
  std::vector<uint8_t> kind;
  double bf = 0;
  const double* Xw;
  std::vector<uint8_t> level;
  std::vector<uint8_t> robust;  // per-edge kernel presence (removed at round index 2 for all)
  double delta, delta_s;
  std::vector<double> err;
  double lambda = 0, ni = 2;
  double H[36], b[6], x[6];
};

double po_compute_errors(PoseProblem& P) {
  double R[9];
  quat_to_R(P.T.q, R);
  double chi = 0;
  for (int o = 0; o < P.No; o++) {
    if (P.level[o]) continue;
    double pc[3];
    map_point(R, P.T.t, &P.Xw[(size_t)o * 3], pc);
    double* e = &P.err[(size_t)o * 3];
    edge_error3(pc, &P.uv3[(size_t)o * 3], P.kind[o], P.Kof(o), P.bfof(o), e);
    double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    if (P.robust[o]) {
      double rho[3];
      huber(e2, P.kind[o] ? P.delta_s : P.delta, rho);
      chi += rho[0];
    } else {
      chi += e2;
    }
  }
  return chi;
}

void po_build(PoseProblem& P) {
  double R[9];
  quat_to_R(P.T.q, R);
  std::fill(P.H, P.H + 36, 0.0);
  std::fill(P.b, P.b + 6, 0.0);
  for (int o = 0; o < P.No; o++) {
    if (P.level[o]) continue;
    double pc[3], Jp[18];
    map_point(R, P.T.t, &P.Xw[(size_t)o * 3], pc);
    edge_jac3(R, pc, P.kind[o], P.Kof(o), P.bfof(o), Jp, nullptr);
    const double* e = &P.err[(size_t)o * 3];
    double w = 1.0;
    if (P.robust[o]) {
      double rho[3];
      huber(e[0] * e[0] + e[1] * e[1] + e[2] * e[2], P.kind[o] ? P.delta_s : P.delta, rho);
      w = rho[1];
    }
    double r0 = -w * e[0], r1 = -w * e[1], r2 = -w * e[2];
    for (int a = 0; a < 6; a++) {
      P.b[a] += Jp[a] * r0 + Jp[6 + a] * r1 + Jp[12 + a] * r2;
      for (int c = 0; c < 6; c++) P.H[a * 6 + c] += w * (Jp[a] * Jp[c] + Jp[6 + a] * Jp[6 + c] + Jp[12 + a] * Jp[12 + c]);
    }
  }
}

bool po_solve(PoseProblem& P, double lambda) {  // dense Cholesky on (H + lambda I)
  double L[36];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j <= i; j++) {
      double s = P.H[i * 6 + j] + (i == j ? lambda : 0.0);
      for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k];
      if (j < i) {
        L[i * 6 + j] = s / L[j * 6 + j];
      } else {
        if (!(s > 0.0)) return false;
        L[i * 6 + i] = std::sqrt(s);
      }
    }
  for (int i = 0; i < 6; i++) {
    double s = P.b[i];
    for (int k = 0; k < i; k++) s -= L[i * 6 + k] * P.x[k];
    P.x[i] = s / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; i--) {
    double s = P.x[i];
    for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * P.x[k];
    P.x[i] = s / L[i * 6 + i];
  }
  return true;
}

int po_optimize(PoseProblem& P, int n_iter, urmvo_oracle_stats* st, double* chi_out) {
  int n_active = 0;
  for (int o = 0; o < P.No; o++) n_active += !P.level[o];
  if (n_active == 0) {  // g2o: "0 vertices to optimize" → optimize() returns -1 without touching anything
    if (chi_out) *chi_out = 0;
    return 0;
  }
  LMTrace tr{st};
  double currentChi = 0;
  int it = 0;
  for (; it < n_iter; it++) {
    currentChi = po_compute_errors(P);
    double chi_before = currentChi, tempChi = currentChi;
    po_build(P);
    if (it == 0) {
      double m = 0;
      for (int j = 0; j < 6; j++) m = std::max(std::fabs(P.H[j * 7]), m);
      P.lambda = 1e-5 * m;
      P.ni = 2;
    }
    double rho = 0;
    int qmax = 0, accepted = 0;
    bool lambda_bad = false;
    do {
      SE3 backup = P.T;
      bool ok2 = po_solve(P, P.lambda);
      if (ok2) P.T = se3_mul(se3_exp(P.x), P.T);
      tempChi = po_compute_errors(P);
      if (!ok2) tempChi = DBL_MAX;
      rho = currentChi - tempChi;
      double scale = 0;
      for (int j = 0; j < 6; j++) scale += P.x[j] * (P.lambda * P.x[j] + P.b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        P.lambda *= std::max(1. / 3., alpha);
        P.ni = 2;
        currentChi = tempChi;
        accepted = 1;
      } else {
        P.lambda *= P.ni;
        P.ni *= 2;
        P.T = backup;
        accepted = 0;
        if (!std::isfinite(P.lambda)) { lambda_bad = true; break; }
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    tr.row(chi_before, currentChi, P.lambda, qmax, accepted);
    if (qmax == 10 || rho == 0 || lambda_bad || !std::isfinite(P.lambda)) { it++; break; }
  }
  if (chi_out) *chi_out = currentChi;
  return it;
}

}  // namespace

static int pose_only_impl(double* pose, int No, const double* uv, int uv_stride, const uint8_t* kind,
                          const double* Xw, const double* intr, double bf, double chi2_thr,
                          double chi2_thr_stereo, int rounds, int its_per_round, uint8_t* inlier,
                          urmvo_oracle_stats* stats, int n_models = 0);

extern "C" int urmvo_oracle_pose_only(double* pose, int No, const double* uv, const double* Xw,
                                      const double* intr, double chi2_thr, int rounds,
                                      int its_per_round, uint8_t* inlier,
                                      urmvo_oracle_stats* stats) {
  return pose_only_impl(pose, No, uv, 2, nullptr, Xw, intr, 0.0, chi2_thr, chi2_thr, rounds, its_per_round, inlier, stats);
}

// Mono + stereo edges (src/g2o_optimization.cc:235-258): uv3 = (u, v, u_right), kind[o] = 1 for a stereo edge,
// intr5 = fx fy cx cy bf.  The edge order of the reference (all mono edges, then all stereo edges) only
// fixes the summation order; the caller's order is used here.
extern "C" int urmvo_oracle_pose_only_stereo(double* pose, int No, const double* uv3, const uint8_t* kind,
                                             const double* Xw, const double* intr5, double chi2_thr_mono,
                                             double chi2_thr_stereo, int rounds, int its_per_round,
                                             uint8_t* inlier, urmvo_oracle_stats* stats) {
  return pose_only_impl(pose, No, uv3, 3, kind, Xw, intr5, intr5[4], chi2_thr_mono, chi2_thr_stereo, rounds,
                        its_per_round, inlier, stats);
}

// Per-constraint camera models (camera_list[mpc->id_camera], src/g2o_optimization.cc:221-224, :243-250); same
// encoding as urmvo_oracle_local_ba_multicam.
extern "C" int urmvo_oracle_pose_only_multicam(double* pose, int No, const double* uv3, const uint8_t* kind_model,
                                               const double* Xw, int n_models, const double* intr5_tab,
                                               double chi2_thr_mono, double chi2_thr_stereo, int rounds,
                                               int its_per_round, uint8_t* inlier, urmvo_oracle_stats* stats) {
  if (n_models <= 0 || n_models > 128) return -1;
  for (int o = 0; o < No; o++)
    if ((kind_model[o] >> 1) >= n_models) return -1;
  return pose_only_impl(pose, No, uv3, 3, kind_model, Xw, intr5_tab, intr5_tab[4], chi2_thr_mono, chi2_thr_stereo,
                        rounds, its_per_round, inlier, stats, n_models);
}

static int pose_only_impl(double* pose, int No, const double* uv, int uv_stride, const uint8_t* kind,
                          const double* Xw, const double* intr, double bf, double chi2_thr,
                          double chi2_thr_stereo, int rounds, int its_per_round, uint8_t* inlier,
                          urmvo_oracle_stats* stats, int n_models) {
  if (stats) std::memset(stats, 0, sizeof(*stats));
  PoseProblem P;
  P.No = No;
  P.K = Intr{intr[0], intr[1], intr[2], intr[3]};
  P.bf = bf;
  P.uv3.assign((size_t)No * 3, 0.0);
  P.kind.assign(No, 0);
  if (n_models > 0) {
    for (int m = 0; m < n_models; m++) {
      P.Ks.push_back(Intr{intr[m * 5], intr[m * 5 + 1], intr[m * 5 + 2], intr[m * 5 + 3]});
      P.bfs.push_back(intr[m * 5 + 4]);
    }
    P.model.assign(No, 0);
  }
  for (int o = 0; o < No; o++) {
    P.uv3[(size_t)o * 3] = uv[(size_t)o * uv_stride];
    P.uv3[(size_t)o * 3 + 1] = uv[(size_t)o * uv_stride + 1];
    const bool st = uv_stride == 3 && kind && (n_models > 0 ? (kind[o] & 1) : kind[o]);
    if (st) { P.uv3[(size_t)o * 3 + 2] = uv[(size_t)o * 3 + 2]; P.kind[o] = 1; }
    if (n_models > 0) P.model[o] = kind ? (uint8_t)(kind[o] >> 1) : 0;
  }
  P.Xw = Xw;
  P.level.assign(No, 0);
  P.robust.assign(No, 1);
  P.err.assign((size_t)No * 3, 0.0);
  P.delta = (double)(float)std::sqrt(chi2_thr);           // :205
  P.delta_s = (double)(float)std::sqrt(chi2_thr_stereo);  // :206
  std::fill(P.x, P.x + 6, 0.0);
  const SE3 T0 = se3_inverse(se3_from(pose, pose + 4));  // :198-199
  P.T = T0;
  int num_outlier = 0;
  for (int iter = 0; iter < rounds; iter++) {
    P.T = T0;  // :265-266 reset to the INPUT pose every round
    double chi = 0;
    int n = po_optimize(P, its_per_round, stats, &chi);  // :267-268
    if (stats && iter < 4) { stats->iters[iter] = n; stats->chi2_final[iter] = chi; stats->lambda_final[iter] = P.lambda; }
    num_outlier = 0;
    double R[9];
    quat_to_R(P.T.q, R);
    for (int o = 0; o < No; o++) {
      if (!inlier[o]) {  // :273-275 e->computeError() at the current estimate
        double pc[3];
        map_point(R, P.T.t, &Xw[(size_t)o * 3], pc);
        edge_error3(pc, &P.uv3[(size_t)o * 3], P.kind[o], P.Kof(o), P.bfof(o), &P.err[(size_t)o * 3]);
      }
      const double* e = &P.err[(size_t)o * 3];
      const float chi2 = (float)(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);  // :277 / :297
      if (chi2 > (P.kind[o] ? chi2_thr_stereo : chi2_thr)) {
        inlier[o] = 0; P.level[o] = 1; num_outlier++;
      } else {
        inlier[o] = 1; P.level[o] = 0;
      }
      if (iter == 2) P.robust[o] = 0;  // :287-288
    }
    if (No < 10) break;  // :310-311 optimizer.edges().size() < 10
  }
  SE3 Twc = se3_inverse(P.T);  // :315-317
  std::memcpy(pose, Twc.q, 4 * sizeof(double));
  std::memcpy(pose + 4, Twc.t, 3 * sizeof(double));
  return No - num_outlier;  // :319-320
}

extern "C" int urmvo_oracle_pose_only_batch(int B, const int32_t* obs_offset, double* poses,
                                            const double* uv, const double* Xw,
                                            const double* intr, double chi2_thr, int rounds,
                                            int its_per_round, uint8_t* inlier,
                                            int32_t* n_inlier, int n_threads) {
  auto work = [&](int f) {
    int o0 = obs_offset[f], n = obs_offset[f + 1] - o0;
    int r = urmvo_oracle_pose_only(poses + (size_t)f * 7, n, uv + (size_t)o0 * 2,
                                   Xw + (size_t)o0 * 3, intr, chi2_thr, rounds, its_per_round,
                                   inlier + o0, nullptr);
    if (n_inlier) n_inlier[f] = r;
  };
  if (n_threads <= 1) {
    for (int f = 0; f < B; f++) work(f);
  } else {
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
      th.emplace_back([&] { for (int f = next++; f < B; f = next++) work(f); });
    for (auto& t : th) t.join();
  }
  return 0;
}

// ------------------------------------------------------------------ unit-test hooks

extern "C" int urmvo_oracle_edge(const double* Tcw, const double* X, const double* uv,
                                 const double* intr, double* e, double* Jpose, double* Jpoint) {
  Intr K{intr[0], intr[1], intr[2], intr[3]};
  double R[9], pc[3];
  quat_to_R(Tcw, R);
  map_point(R, Tcw + 4, X, pc);
  edge_error(pc, uv, K, e);
  edge_jac(R, pc, K, Jpose, Jpoint);
  return pc[2] > 0.0;
}

extern "C" void urmvo_oracle_huber(double e2, double delta, double* rho) { huber(e2, delta, rho); }

extern "C" void urmvo_oracle_se3_oplus(double* Tcw, const double* update) {
  SE3 T;
  std::memcpy(T.q, Tcw, 4 * sizeof(double));
  std::memcpy(T.t, Tcw + 4, 3 * sizeof(double));
  T = se3_mul(se3_exp(update), T);
  std::memcpy(Tcw, T.q, 4 * sizeof(double));
  std::memcpy(Tcw + 4, T.t, 3 * sizeof(double));
}

extern "C" void urmvo_oracle_se3_inverse(const double* Tin, double* Tout) {
  SE3 T = se3_inverse(se3_from(Tin, Tin + 4));
  std::memcpy(Tout, T.q, 4 * sizeof(double));
  std::memcpy(Tout + 4, T.t, 3 * sizeof(double));
}

// EdgeStereoSE3ProjectXYZ: e (3), Jpose 3x6, Jpoint 3x3; uv3 = (u, v, u_right), intr5 = fx fy cx cy bf.
extern "C" int urmvo_oracle_edge_stereo(const double* Tcw, const double* X, const double* uv3,
                                        const double* intr5, double* e, double* Jpose, double* Jpoint) {
  Intr K{intr5[0], intr5[1], intr5[2], intr5[3]};
  double R[9], pc[3];
  quat_to_R(Tcw, R);
  map_point(R, Tcw + 4, X, pc);
  edge_error3(pc, uv3, 1, K, intr5[4], e);
  edge_jac3(R, pc, 1, K, intr5[4], Jpose, Jpoint);
  return pc[2] > 0.0;
}
