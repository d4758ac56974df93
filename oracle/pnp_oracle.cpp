// oracle/pnp_oracle.cpp — CPU restatement of SolvePnPWithCV (reference src/g2o_optimization.cc:323-377):
//     cv::solvePnPRansac(object_points, image_points, camera_matrix, dist_coeffs (all zero, src/camera.cc:164-166),
//                        rotation_vector, translation_vector, false, 100, 20.0, 0.99, cv_inliers);
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// PARITY: PARTLY PINNED.  The algorithm lives in OpenCV 4.2 (calib3d: solvepnp.cpp / ptsetreg.cpp / epnp.cpp),
// which the reference links but does not vendor.  Restated from the published sources:
//   solvePnPRansac: float points, 5-point minimal sets, EPnP as the RANSAC kernel (SOLVEPNP_ITERATIVE default
//       -> model_points = 5, ransac_kernel_method = SOLVEPNP_EPNP), reprojection error in float
//       ((float)|ip - (float)proj|^2 <= (float)(20*20)), "strictly more inliers wins", RANSACUpdateNumIters,
//       final solvePnP(SOLVEPNP_ITERATIVE) on the inliers of the best model;
//   RANSACPointSetRegistrator::run / getSubset with cv::RNG(-1) (the same generator as fm_oracle.cpp; the
//       PnP callback has no checkSubset);
//   epnp::compute_pose: PCA control points, barycentric coordinates, the 2n x 12 system, its four smallest
//       right singular vectors, the three beta initialisations, 5 Gauss-Newton steps each (Householder QR),
//       absolute orientation of the camera-frame points, the candidate with the smallest reprojection error.
// What cannot be pinned bit for bit: OpenCV takes its SVDs from LAPACK.  With FIVE points the 10 x 12 EPnP
// system has a two-dimensional null space, so the "smallest" singular vectors are an arbitrary basis of it and
// about a quarter of the minimal-sample poses depend on that basis (measured with numpy against itself).
// This file fixes the basis by a fully specified cyclic Jacobi eigen-solver; the CUDA path follows the same
// specification and must agree with this file exactly (inlier masks, iteration counts).  Against the real
// library the restatement is pinned statistically: tests/golden/make_golden_pnp.py runs cv2.solvePnPRansac
// (cv2 4.13 — its EPnP differs in detail from 4.2's, the version the reference links) on seeded scenes and
// commits inlier sets and poses; tests/test_golden_pnp.py requires identical inlier sets on the well-posed
// scenes and the refined pose within 1e-5, because the final pose is the least-squares optimum over the
// inlier set and does not depend on which minimal sample found it.
// The final refinement is a Levenberg-Marquardt minimisation of the reprojection error over the inliers
// started from the best RANSAC model (OpenCV: DLT initialisation + CvLevMarq, 20 iterations, eps = FLT_EPSILON;
// both converge to the same minimum, OpenCV to ~1e-7 relative).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

struct CvRng {
  uint64_t state = 0xffffffffffffffffull;
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a + (int)(next() % (unsigned)(b - a)); }
};

void draw_subset(int count, CvRng& rng, int* idx) {  // getSubset, modelPoints = 5, no checkSubset
  for (int i = 0; i < 5; i++) {
    int v = rng.uniform(0, count);
    while (std::find(idx, idx + i, v) != idx + i) v = rng.uniform(0, count);
    idx[i] = v;
  }
}

int update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = std::min(std::max(p, 0.), 1.);
  ep = std::min(std::max(ep, 0.), 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::nearbyint(num / denom);
}

// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (row-major, destroyed).  V: eigenvectors as
// COLUMNS.  Pairs (p, q), p < q, in row order; a pair is rotated when a_pq != 0; at most 60 sweeps; stops
// when the off-diagonal sum of squares is <= 1e-32 x the squared Frobenius norm of the input.
void jacobi_eig(int n, double* A, double* V) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = i == j ? 1.0 : 0.0;
  double fro = 0;
  for (int i = 0; i < n * n; i++) fro += A[i] * A[i];
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
    if (off <= 1e-32 * fro) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {  // columns p, q
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {  // rows p, q
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
}

// The 12 x 12 problem uses the PARALLEL (round-robin) ordering: a sweep is 11 rounds of 6 disjoint pairs
// (pair 0 of round r is (r, 11); pair k = 1..5 is ((r + k) mod 11, (r - k) mod 11), smaller index first).  The six
// rotations of a round are computed from the matrix as it stands at the start of the round, then applied:
// first to the columns (and to V), then to the rows.  A pair with a_pq == 0 gets the identity.  Same stopping
// rule as above.  (The CUDA path computes the six rotations of a round on six lanes at once.)
void jacobi_eig12_rr(double* A, double* V) {
  const int n = 12;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = i == j ? 1.0 : 0.0;
  double fro = 0;
  for (int i = 0; i < n * n; i++) fro += A[i] * A[i];
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
    if (off <= 1e-32 * fro) break;
    for (int r = 0; r < 11; r++) {
      int P[6], Q[6];
      double C[6], S[6];
      for (int k = 0; k < 6; k++) {
        const int a = k == 0 ? r : (r + k) % 11, b = k == 0 ? 11 : (r - k + 11) % 11;
        P[k] = std::min(a, b);
        Q[k] = std::max(a, b);
        const double apq = A[P[k] * n + Q[k]];
        if (apq == 0.0) { C[k] = 1.0; S[k] = 0.0; continue; }
        const double theta = (A[Q[k] * n + Q[k]] - A[P[k] * n + P[k]]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        C[k] = 1.0 / std::sqrt(t * t + 1.0);
        S[k] = t * C[k];
      }
      for (int k = 0; k < 6; k++)
        for (int i = 0; i < n; i++) {
          const double akp = A[i * n + P[k]], akq = A[i * n + Q[k]];
          A[i * n + P[k]] = C[k] * akp - S[k] * akq;
          A[i * n + Q[k]] = S[k] * akp + C[k] * akq;
          const double vkp = V[i * n + P[k]], vkq = V[i * n + Q[k]];
          V[i * n + P[k]] = C[k] * vkp - S[k] * vkq;
          V[i * n + Q[k]] = S[k] * vkp + C[k] * vkq;
        }
      for (int k = 0; k < 6; k++)
        for (int i = 0; i < n; i++) {
          const double apk = A[P[k] * n + i], aqk = A[Q[k] * n + i];
          A[P[k] * n + i] = C[k] * apk - S[k] * aqk;
          A[Q[k] * n + i] = S[k] * apk + C[k] * aqk;
        }
    }
  }
}

// indices of the diagonal of A sorted by value (stable insertion sort), ascending or descending
void sort_diag(int n, const double* A, int* order, bool descending) {
  for (int i = 0; i < n; i++) order[i] = i;
  for (int i = 1; i < n; i++) {
    const int o = order[i];
    const double v = A[o * n + o];
    int j = i - 1;
    while (j >= 0 && (descending ? A[order[j] * n + order[j]] < v : A[order[j] * n + order[j]] > v)) {
      order[j + 1] = order[j];
      j--;
    }
    order[j + 1] = o;
  }
}

// epnp::qr_solve: least squares by Householder QR (A: nr x nc row-major, destroyed; b destroyed).
// Returns false for a zero column (singular).
bool qr_solve(int nr, int nc, double* A, double* b, double* X) {
  double A1[8], A2[8];
  for (int k = 0; k < nc; k++) {
    double eta = 0;
    for (int i = k; i < nr; i++) eta = std::max(eta, std::fabs(A[i * nc + k]));
    if (eta == 0) return false;
    const double inv_eta = 1.0 / eta;
    double sum2 = 0;
    for (int i = k; i < nr; i++) {
      A[i * nc + k] *= inv_eta;
      sum2 += A[i * nc + k] * A[i * nc + k];
    }
    double sigma = std::sqrt(sum2);
    if (A[k * nc + k] < 0) sigma = -sigma;
    A[k * nc + k] += sigma;
    A1[k] = sigma * A[k * nc + k];
    A2[k] = -eta * sigma;
    for (int j = k + 1; j < nc; j++) {
      double sum = 0;
      for (int i = k; i < nr; i++) sum += A[i * nc + k] * A[i * nc + j];
      const double tau = sum / A1[k];
      for (int i = k; i < nr; i++) A[i * nc + j] -= tau * A[i * nc + k];
    }
  }
  for (int j = 0; j < nc; j++) {
    double tau = 0;
    for (int i = j; i < nr; i++) tau += A[i * nc + j] * b[i];
    tau /= A1[j];
    for (int i = j; i < nr; i++) b[i] -= tau * A[i * nc + j];
  }
  X[nc - 1] = b[nc - 1] / A2[nc - 1];
  for (int i = nc - 2; i >= 0; i--) {
    double sum = 0;
    for (int j = i + 1; j < nc; j++) sum += A[i * nc + j] * X[j];
    X[i] = (b[i] - sum) / A2[i];
  }
  return true;
}

struct Epnp {
  double fu, fv, uc, vc;
  double pws[5][3], us[5][2], al[5][4], cws[4][3];
  double ut[4][12];  // the four eigenvectors of M^T M with the smallest eigenvalues: ut[0] the smallest
  double L[6][10], rho[6];

  bool prepare() {
    const int n = 5;
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int i = 0; i < n; i++) s += pws[i][j];
      cws[0][j] = s / n;
    }
    double A[9] = {0}, V[9];
    for (int i = 0; i < n; i++)
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) A[a * 3 + b] += (pws[i][a] - cws[0][a]) * (pws[i][b] - cws[0][b]);
    jacobi_eig(3, A, V);
    int ord[3];
    sort_diag(3, A, ord, true);
    for (int i = 1; i < 4; i++) {
      const int o = ord[i - 1];
      const double ev = A[o * 3 + o] > 0 ? A[o * 3 + o] : 0.0;
      const double k = std::sqrt(ev / n);
      for (int j = 0; j < 3; j++) cws[i][j] = cws[0][j] + k * V[j * 3 + o];
    }
    // barycentric coordinates: inverse of CC = [c1 - c0 | c2 - c0 | c3 - c0] by cofactors
    double cc[9];
    for (int i = 0; i < 3; i++)
      for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
    const double c00 = cc[4] * cc[8] - cc[5] * cc[7], c01 = cc[5] * cc[6] - cc[3] * cc[8], c02 = cc[3] * cc[7] - cc[4] * cc[6];
    const double det = cc[0] * c00 + cc[1] * c01 + cc[2] * c02;
    if (!(std::fabs(det) > 0)) return false;
    const double id = 1.0 / det;
    double ci[9];
    ci[0] = c00 * id; ci[1] = (cc[2] * cc[7] - cc[1] * cc[8]) * id; ci[2] = (cc[1] * cc[5] - cc[2] * cc[4]) * id;
    ci[3] = c01 * id; ci[4] = (cc[0] * cc[8] - cc[2] * cc[6]) * id; ci[5] = (cc[2] * cc[3] - cc[0] * cc[5]) * id;
    ci[6] = c02 * id; ci[7] = (cc[1] * cc[6] - cc[0] * cc[7]) * id; ci[8] = (cc[0] * cc[4] - cc[1] * cc[3]) * id;
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < 3; j++)
        al[i][1 + j] = ci[3 * j] * (pws[i][0] - cws[0][0]) + ci[3 * j + 1] * (pws[i][1] - cws[0][1]) +
                       ci[3 * j + 2] * (pws[i][2] - cws[0][2]);
      al[i][0] = 1.0 - al[i][1] - al[i][2] - al[i][3];
    }
    double M[10][12];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < 4; j++) {
        M[2 * i][3 * j] = al[i][j] * fu; M[2 * i][3 * j + 1] = 0.0; M[2 * i][3 * j + 2] = al[i][j] * (uc - us[i][0]);
        M[2 * i + 1][3 * j] = 0.0; M[2 * i + 1][3 * j + 1] = al[i][j] * fv; M[2 * i + 1][3 * j + 2] = al[i][j] * (vc - us[i][1]);
      }
    double MtM[144], VV[144];
    for (int a = 0; a < 12; a++)
      for (int b = 0; b < 12; b++) {
        double s = 0;
        for (int r = 0; r < 10; r++) s += M[r][a] * M[r][b];
        MtM[a * 12 + b] = s;
      }
    jacobi_eig12_rr(MtM, VV);
    int ord12[12];
    sort_diag(12, MtM, ord12, false);
    for (int i = 0; i < 4; i++)
      for (int k = 0; k < 12; k++) ut[i][k] = VV[k * 12 + ord12[i]];
    // L_6x10 and rho
    double dv[4][6][3];
    for (int i = 0; i < 4; i++) {
      int a = 0, b = 1;
      for (int j = 0; j < 6; j++) {
        for (int k = 0; k < 3; k++) dv[i][j][k] = ut[i][3 * a + k] - ut[i][3 * b + k];
        b++;
        if (b > 3) { a++; b = a + 1; }
      }
    }
    auto dot = [](const double* x, const double* y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
    for (int i = 0; i < 6; i++) {
      L[i][0] = dot(dv[0][i], dv[0][i]);
      L[i][1] = 2.0 * dot(dv[0][i], dv[1][i]);
      L[i][2] = dot(dv[1][i], dv[1][i]);
      L[i][3] = 2.0 * dot(dv[0][i], dv[2][i]);
      L[i][4] = 2.0 * dot(dv[1][i], dv[2][i]);
      L[i][5] = dot(dv[2][i], dv[2][i]);
      L[i][6] = 2.0 * dot(dv[0][i], dv[3][i]);
      L[i][7] = 2.0 * dot(dv[1][i], dv[3][i]);
      L[i][8] = 2.0 * dot(dv[2][i], dv[3][i]);
      L[i][9] = dot(dv[3][i], dv[3][i]);
    }
    const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    for (int i = 0; i < 6; i++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += (cws[pa[i]][k] - cws[pb[i]][k]) * (cws[pa[i]][k] - cws[pb[i]][k]);
      rho[i] = s;
    }
    return true;
  }

  bool betas_approx(int which, double* be) {
    static const int cols[3][5] = {{0, 1, 3, 6, -1}, {0, 1, 2, -1, -1}, {0, 1, 2, 3, 4}};
    const int nc = which == 0 ? 4 : which == 1 ? 3 : 5;
    double A[30], b[6], x[5];
    for (int i = 0; i < 6; i++) {
      for (int j = 0; j < nc; j++) A[i * nc + j] = L[i][cols[which][j]];
      b[i] = rho[i];
    }
    if (!qr_solve(6, nc, A, b, x)) return false;
    be[0] = be[1] = be[2] = be[3] = 0.0;
    if (which == 0) {
      if (x[0] < 0) { be[0] = std::sqrt(-x[0]); be[1] = -x[1] / be[0]; be[2] = -x[2] / be[0]; be[3] = -x[3] / be[0]; }
      else { be[0] = std::sqrt(x[0]); be[1] = x[1] / be[0]; be[2] = x[2] / be[0]; be[3] = x[3] / be[0]; }
    } else {
      if (x[0] < 0) { be[0] = std::sqrt(-x[0]); be[1] = x[2] < 0 ? std::sqrt(-x[2]) : 0.0; }
      else { be[0] = std::sqrt(x[0]); be[1] = x[2] > 0 ? std::sqrt(x[2]) : 0.0; }
      if (x[1] < 0) be[0] = -be[0];
      if (which == 2) be[2] = x[3] / be[0];
    }
    return true;
  }

  bool gauss_newton(double* be) {
    for (int it = 0; it < 5; it++) {
      double A[24], b[6], x[4];
      for (int i = 0; i < 6; i++) {
        const double* r = L[i];
        A[i * 4 + 0] = 2 * r[0] * be[0] + r[1] * be[1] + r[3] * be[2] + r[6] * be[3];
        A[i * 4 + 1] = r[1] * be[0] + 2 * r[2] * be[1] + r[4] * be[2] + r[7] * be[3];
        A[i * 4 + 2] = r[3] * be[0] + r[4] * be[1] + 2 * r[5] * be[2] + r[8] * be[3];
        A[i * 4 + 3] = r[6] * be[0] + r[7] * be[1] + r[8] * be[2] + 2 * r[9] * be[3];
        b[i] = rho[i] - (r[0] * be[0] * be[0] + r[1] * be[0] * be[1] + r[2] * be[1] * be[1] + r[3] * be[0] * be[2] +
                         r[4] * be[1] * be[2] + r[5] * be[2] * be[2] + r[6] * be[0] * be[3] + r[7] * be[1] * be[3] +
                         r[8] * be[2] * be[3] + r[9] * be[3] * be[3]);
      }
      if (!qr_solve(6, 4, A, b, x)) return false;
      for (int i = 0; i < 4; i++) be[i] += x[i];
    }
    return true;
  }

  // compute_ccs, compute_pcs, solve_for_sign, estimate_R_and_t, reprojection_error
  double R_and_t(const double* be, double* R, double* t) {
    const int n = 5;
    double ccs[4][3] = {{0}}, pcs[5][3];
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++)
        for (int k = 0; k < 3; k++) ccs[j][k] += be[i] * ut[i][3 * j + k];
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++)
        pcs[i][k] = al[i][0] * ccs[0][k] + al[i][1] * ccs[1][k] + al[i][2] * ccs[2][k] + al[i][3] * ccs[3][k];
    if (pcs[0][2] < 0.0)
      for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) pcs[i][k] = -pcs[i][k];
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) { pc0[k] += pcs[i][k]; pw0[k] += pws[i][k]; }
    for (int k = 0; k < 3; k++) { pc0[k] /= n; pw0[k] /= n; }
    double B[9] = {0};  // ABt = sum (pc - pc0)(pw - pw0)^T
    for (int i = 0; i < n; i++)
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) B[a * 3 + b] += (pcs[i][a] - pc0[a]) * (pws[i][b] - pw0[b]);
    // SVD of B through the eigen-decomposition of B^T B: B = U S V^T, R = U V^T
    double BtB[9], V[9];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) BtB[a * 3 + b] = B[0 * 3 + a] * B[0 * 3 + b] + B[1 * 3 + a] * B[1 * 3 + b] + B[2 * 3 + a] * B[2 * 3 + b];
    jacobi_eig(3, BtB, V);
    int ord[3];
    sort_diag(3, BtB, ord, true);
    double Vs[3][3], Us[3][3], sv[3];
    for (int c = 0; c < 3; c++) {
      const int o = ord[c];
      sv[c] = std::sqrt(BtB[o * 3 + o] > 0 ? BtB[o * 3 + o] : 0.0);
      for (int r = 0; r < 3; r++) Vs[r][c] = V[r * 3 + o];
    }
    for (int c = 0; c < 3; c++) {
      if (c < 2 || sv[2] > 1e-9 * sv[0]) {
        double u[3], nn = 0;
        for (int r = 0; r < 3; r++) {
          u[r] = B[r * 3 + 0] * Vs[0][c] + B[r * 3 + 1] * Vs[1][c] + B[r * 3 + 2] * Vs[2][c];
          nn += u[r] * u[r];
        }
        nn = std::sqrt(nn);
        if (!(nn > 0)) return 1e300;
        for (int r = 0; r < 3; r++) Us[r][c] = u[r] / nn;
      } else {  // rank 2: complete U to a rotation-compatible frame
        Us[0][2] = Us[1][0] * Us[2][1] - Us[2][0] * Us[1][1];
        Us[1][2] = Us[2][0] * Us[0][1] - Us[0][0] * Us[2][1];
        Us[2][2] = Us[0][0] * Us[1][1] - Us[1][0] * Us[0][1];
        const double dv = Vs[0][0] * (Vs[1][1] * Vs[2][2] - Vs[1][2] * Vs[2][1]) - Vs[0][1] * (Vs[1][0] * Vs[2][2] - Vs[1][2] * Vs[2][0]) +
                          Vs[0][2] * (Vs[1][0] * Vs[2][1] - Vs[1][1] * Vs[2][0]);
        if (dv < 0)
          for (int r = 0; r < 3; r++) Us[r][2] = -Us[r][2];
      }
    }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) R[i * 3 + j] = Us[i][0] * Vs[j][0] + Us[i][1] * Vs[j][1] + Us[i][2] * Vs[j][2];
    const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] - R[1] * R[3] * R[8] -
                       R[0] * R[5] * R[7];
    if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
    for (int k = 0; k < 3; k++) t[k] = pc0[k] - (R[k * 3] * pw0[0] + R[k * 3 + 1] * pw0[1] + R[k * 3 + 2] * pw0[2]);
    double sum = 0;
    for (int i = 0; i < n; i++) {
      const double Xc = R[0] * pws[i][0] + R[1] * pws[i][1] + R[2] * pws[i][2] + t[0];
      const double Yc = R[3] * pws[i][0] + R[4] * pws[i][1] + R[5] * pws[i][2] + t[1];
      const double iz = 1.0 / (R[6] * pws[i][0] + R[7] * pws[i][1] + R[8] * pws[i][2] + t[2]);
      const double ue = uc + fu * Xc * iz, ve = vc + fv * Yc * iz;
      sum += std::sqrt((us[i][0] - ue) * (us[i][0] - ue) + (us[i][1] - ve) * (us[i][1] - ve));
    }
    return sum / n;
  }

  bool compute_pose(double* R, double* t) {
    if (!prepare()) return false;
    double best = 1e300;
    bool any = false;
    for (int which = 0; which < 3; which++) {
      double be[4], Rc[9], tc[3];
      if (!betas_approx(which, be) || !gauss_newton(be)) continue;
      const double e = R_and_t(be, Rc, tc);
      if (!(e == e)) continue;  // NaN
      if (!any || e < best) {   // N = 1; 2 if smaller; 3 if smaller than the current one
        best = e;
        std::memcpy(R, Rc, sizeof(Rc));
        std::memcpy(t, tc, sizeof(tc));
        any = true;
      }
    }
    return any;
  }
};

// minimal solver on 5 correspondences given by idx
bool epnp5(const float* obj, const float* img, const double* K4, const int* idx, double* R, double* t) {
  Epnp e;
  e.fu = K4[0]; e.fv = K4[1]; e.uc = K4[2]; e.vc = K4[3];
  const double ifx = 1.0 / K4[0], ify = 1.0 / K4[1];
  for (int i = 0; i < 5; i++) {
    const int p = idx[i];
    for (int k = 0; k < 3; k++) e.pws[i][k] = obj[3 * p + k];
    // cv::undistortPoints on float points with zero distortion: double arithmetic, FLOAT result; epnp::init_points
    // maps the normalised point back to pixels in double
    const float xn = (float)(((double)img[2 * p] - K4[2]) * ifx), yn = (float)(((double)img[2 * p + 1] - K4[3]) * ify);
    e.us[i][0] = (double)xn * K4[0] + K4[2];
    e.us[i][1] = (double)yn * K4[1] + K4[3];
  }
  return e.compute_pose(R, t);
}

// PnPRansacCallback::computeError: projectPoints in double, float result, float difference and float squared norm
inline float reproj_err(const double* R, const double* t, const double* K4, const float* X, const float* uv) {
  const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
  z = z ? 1.0 / z : 1.0;
  const float pu = (float)(x * z * K4[0] + K4[2]), pv = (float)(y * z * K4[1] + K4[3]);
  const float dx = uv[0] - pu, dy = uv[1] - pv;
  float s = dx * dx;
  s += dy * dy;
  return s;
}

void so3_exp(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
  double a, b;
  if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
  else { a = std::sin(th) / th; b = (1.0 - std::cos(th)) / th2; }
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double w2 = 0;
      for (int k = 0; k < 3; k++) w2 += W[i * 3 + k] * W[k * 3 + j];
      R[i * 3 + j] = (i == j ? 1.0 : 0.0) + a * W[i * 3 + j] + b * w2;
    }
}

// Levenberg-Marquardt over the inliers: minimise sum |uv - proj(R X + t)|^2; update R <- exp(w) R, t <- exp(w) t + v
double refine(int N, const float* obj, const float* img, const double* K4, const uint8_t* mask, double* R, double* t) {
  auto cost_and_normal = [&](const double* Rr, const double* tt, double* H, double* g) {
    double c = 0;
    if (H) { std::memset(H, 0, 36 * sizeof(double)); std::memset(g, 0, 6 * sizeof(double)); }
    for (int i = 0; i < N; i++) {
      if (!mask[i]) continue;
      const double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
      const double x = Rr[0] * X + Rr[1] * Y + Rr[2] * Z + tt[0], y = Rr[3] * X + Rr[4] * Y + Rr[5] * Z + tt[1];
      const double z = Rr[6] * X + Rr[7] * Y + Rr[8] * Z + tt[2], iz = 1.0 / z;
      const double e0 = (double)img[2 * i] - (K4[0] * x * iz + K4[2]), e1 = (double)img[2 * i + 1] - (K4[1] * y * iz + K4[3]);
      c += e0 * e0 + e1 * e1;
      if (!H) continue;
      // d proj / d (w, v) at the left perturbation exp([w v]) * T: p' = p + w x p + v
      const double a = K4[0] * iz, b = K4[1] * iz, xz = x * iz, yz = y * iz;
      const double J0[6] = {-a * xz * y, a * (z + x * xz), -a * y, a, 0, -a * xz};
      const double J1[6] = {-b * (z + y * yz), b * yz * x, b * x, 0, b, -b * yz};
      for (int p = 0; p < 6; p++) {
        g[p] += J0[p] * e0 + J1[p] * e1;
        for (int q = 0; q < 6; q++) H[p * 6 + q] += J0[p] * J0[q] + J1[p] * J1[q];
      }
    }
    return c;
  };
  double lambda = 1e-3, H[36], g[6];
  double cost = cost_and_normal(R, t, H, g);
  for (int it = 0; it < 100; it++) {
    double A[36], x[6];
    std::memcpy(A, H, sizeof(A));
    for (int p = 0; p < 6; p++) A[p * 6 + p] *= 1.0 + lambda;
    // Cholesky solve
    bool ok = true;
    double Lc[36] = {0};
    for (int i = 0; i < 6 && ok; i++)
      for (int j = 0; j <= i; j++) {
        double s = A[i * 6 + j];
        for (int k = 0; k < j; k++) s -= Lc[i * 6 + k] * Lc[j * 6 + k];
        if (i == j) { if (!(s > 0)) { ok = false; break; } Lc[i * 6 + i] = std::sqrt(s); }
        else Lc[i * 6 + j] = s / Lc[j * 6 + j];
      }
    if (!ok) { lambda *= 10; if (lambda > 1e16) break; continue; }
    double y[6];
    for (int i = 0; i < 6; i++) { double s = g[i]; for (int k = 0; k < i; k++) s -= Lc[i * 6 + k] * y[k]; y[i] = s / Lc[i * 6 + i]; }
    for (int i = 5; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 6; k++) s -= Lc[k * 6 + i] * x[k]; x[i] = s / Lc[i * 6 + i]; }
    double dR[9], Rn[9], tn[3];
    so3_exp(x, dR);
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) Rn[i * 3 + j] = dR[i * 3] * R[j] + dR[i * 3 + 1] * R[3 + j] + dR[i * 3 + 2] * R[6 + j];
      tn[i] = dR[i * 3] * t[0] + dR[i * 3 + 1] * t[1] + dR[i * 3 + 2] * t[2] + x[3 + i];
    }
    const double cn = cost_and_normal(Rn, tn, nullptr, nullptr);
    double step = 0;
    for (int p = 0; p < 6; p++) step = std::max(step, std::fabs(x[p]));
    if (cn <= cost) {
      std::memcpy(R, Rn, sizeof(Rn));
      std::memcpy(t, tn, sizeof(tn));
      cost = cost_and_normal(R, t, H, g);
      lambda = std::max(lambda * 0.1, 1e-12);
      if (step < 1e-13) break;
    } else {
      lambda *= 10;
      if (lambda > 1e16 || step < 1e-13) break;
    }
  }
  return cost;
}

}  // namespace

extern "C" int urmvo_oracle_pnp_subsets(int N, int max_iters, int32_t* idx) {
  CvRng rng;
  for (int n = 0; n < max_iters; n++) draw_subset(N, rng, idx + 5 * n);
  return max_iters;
}

extern "C" int urmvo_oracle_pnp_epnp5(const float* obj, const float* img, const double* K4, const int32_t* idx5,
                                      double* R9, double* t3) {
  return epnp5(obj, img, K4, idx5, R9, t3) ? 1 : 0;
}

extern "C" int urmvo_oracle_pnp_ransac(int N, const float* obj, const float* img, const double* K4, int max_iters,
                                       double reproj_thr, double confidence, double* R9, double* t3, uint8_t* mask,
                                       int32_t* stats4, int32_t* counts) {
  if (N < 6) return -1;  // N == 4 / N == 5 are other OpenCV branches; the reference returns 0 below 8 points
  CvRng rng;
  int niters = std::max(max_iters, 1), max_good = 0, iter = 0, models = 0;
  const float thr = (float)(reproj_thr * reproj_thr);
  double bestR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, bestt[3] = {0, 0, 0};
  std::memset(mask, 0, N);
  std::vector<uint8_t> cur(N);
  for (iter = 0; iter < niters; iter++) {
    int idx[5];
    draw_subset(N, rng, idx);
    double R[9], t[3];
    if (counts) counts[iter] = -1;
    if (!epnp5(obj, img, K4, idx, R, t)) continue;
    models++;
    int good = 0;
    for (int i = 0; i < N; i++) {
      const int f = reproj_err(R, t, K4, obj + 3 * i, img + 2 * i) <= thr;
      cur[i] = (uint8_t)f;
      good += f;
    }
    if (counts) counts[iter] = good;
    if (good > std::max(max_good, 4)) {
      std::memcpy(mask, cur.data(), N);
      std::memcpy(bestR, R, sizeof(R));
      std::memcpy(bestt, t, sizeof(t));
      max_good = good;
      niters = update_num_iters(confidence, (double)(N - good) / N, 5, niters);
    }
  }
  if (stats4) { stats4[0] = iter; stats4[1] = max_good; stats4[2] = models; stats4[3] = 0; }
  if (max_good <= 0) return 0;
  std::memcpy(R9, bestR, sizeof(bestR));
  std::memcpy(t3, bestt, sizeof(bestt));
  refine(N, obj, img, K4, mask, R9, t3);
  return 1;
}
