// oracle/tri_oracle.cpp — CPU restatement of Mapping::TriangulateMappoint
// (reference src/mapping.cc:151-205): multi-view midpoint triangulation of one mappoint from the
// keyframes that observe it, with the rank test of Eigen::ColPivHouseholderQR (threshold 1e-5).
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED: Eigen is not installed here; the QR with
// column pivoting and Eigen's rank rule (|R_ii| > threshold * max_j |R_jj|) are restated from the
// published Eigen sources and cross-checked against numpy in tests/test_oracle_triangulate.py.
//
// Per observer k: bearing b_k = R_k * ((u - cx)/fx, (v - cy)/fy, 1)  (Camera::BackProjectMono,
// src/camera.cc:168-174; R_k, p_k from the keyframe pose T_wc), then
//   A = N I - sum_k b_k b_k^T / |b_k|^2,   rhs = sum_k p_k - sum_k b_k (b_k . p_k) / |b_k|^2,
// fewer than 2 observers or rank(A) < 3 -> false, else X = A^-1 rhs by the QR.
#include <cmath>

#include "oracle.h"

namespace {

// Householder QR with column pivoting of a 3x3 matrix (Eigen::ColPivHouseholderQR semantics):
// returns the rank for `threshold` and, if it is 3, solves A x = b.
int colpiv_qr_solve3(const double* A_in, const double* b_in, double threshold, double* x) {
  double A[3][3], b[3] = {b_in[0], b_in[1], b_in[2]};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A[i][j] = A_in[i * 3 + j];
  int perm[3] = {0, 1, 2};
  double maxpivot = 0.0, diag[3] = {0, 0, 0};
  for (int k = 0; k < 3; k++) {
    // column with the largest remaining norm (first one on ties)
    int best = k;
    double bestn = -1.0;
    for (int j = k; j < 3; j++) {
      double s = 0.0;
      for (int i = k; i < 3; i++) s += A[i][j] * A[i][j];
      if (s > bestn) { bestn = s; best = j; }
    }
    if (best != k) {
      for (int i = 0; i < 3; i++) { const double t = A[i][k]; A[i][k] = A[i][best]; A[i][best] = t; }
      const int t = perm[k]; perm[k] = perm[best]; perm[best] = t;
    }
    // Householder reflector for column k (rows k..2)
    double tail = 0.0;
    for (int i = k + 1; i < 3; i++) tail += A[i][k] * A[i][k];
    const double c0 = A[k][k];
    double beta, tau, v[3] = {0, 0, 0};
    if (tail == 0.0) {
      beta = c0; tau = 0.0;
    } else {
      beta = std::sqrt(c0 * c0 + tail);
      if (c0 >= 0.0) beta = -beta;
      for (int i = k + 1; i < 3; i++) v[i] = A[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    v[k] = 1.0;
    // apply H = I - tau v v^T to the remaining columns and to b
    for (int j = k + 1; j < 3; j++) {
      double s = 0.0;
      for (int i = k; i < 3; i++) s += v[i] * A[i][j];
      s *= tau;
      for (int i = k; i < 3; i++) A[i][j] -= s * v[i];
    }
    {
      double s = 0.0;
      for (int i = k; i < 3; i++) s += v[i] * b[i];
      s *= tau;
      for (int i = k; i < 3; i++) b[i] -= s * v[i];
    }
    A[k][k] = beta;
    for (int i = k + 1; i < 3; i++) A[i][k] = 0.0;
    diag[k] = std::fabs(beta);
    if (diag[k] > maxpivot) maxpivot = diag[k];
  }
  int rank = 0;
  for (int k = 0; k < 3; k++) rank += diag[k] > maxpivot * threshold ? 1 : 0;
  if (rank < 3) return rank;
  double y[3];
  for (int k = 2; k >= 0; k--) {
    double s = b[k];
    for (int j = k + 1; j < 3; j++) s -= A[k][j] * y[j];
    y[k] = s / A[k][k];
  }
  for (int k = 0; k < 3; k++) x[perm[k]] = y[k];
  return 3;
}

}  // namespace

extern "C" int urmvo_oracle_triangulate(int n_obs, const double* Rp /*n_obs*12: R row-major | p*/,
                                        const double* uv /*n_obs*2*/, const double* intr, double* X) {
  if (n_obs < 2) return 0;
  const double fx_inv = 1.0 / intr[0], fy_inv = 1.0 / intr[1], cx = intr[2], cy = intr[3];
  double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, rhs[3] = {0, 0, 0};
  for (int k = 0; k < n_obs; k++) {
    const double* R = Rp + 12 * k;
    const double* p = R + 9;
    const double bx = (uv[2 * k] - cx) * fx_inv, by = (uv[2 * k + 1] - cy) * fy_inv;
    const double b[3] = {R[0] * bx + R[1] * by + R[2], R[3] * bx + R[4] * by + R[5], R[6] * bx + R[7] * by + R[8]};
    const double inv = 1.0 / (b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    const double bp = b[0] * p[0] + b[1] * p[1] + b[2] * p[2];
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) A[i * 3 + j] -= b[i] * inv * b[j];
      rhs[i] += p[i] - b[i] * inv * bp;
    }
  }
  for (int i = 0; i < 3; i++) A[i * 3 + i] += (double)n_obs;
  return colpiv_qr_solve3(A, rhs, 1e-5, X) == 3 ? 1 : 0;
}
