// csrc/pnp_kernels.cu — SolvePnPWithCV on the GPU.
//
// Reference: /root/reference/src/g2o_optimization.cc:323-377
//     cv::solvePnPRansac(object_points, image_points, K, dist = 0, rvec, tvec, false, 100, 20.0, 0.99, inliers)
// called for every tracked frame before FrameOptimization (src/tracking.cc:799).  OpenCV's loop (calib3d
// solvepnp.cpp / ptsetreg.cpp / epnp.cpp) draws 5-point subsets with cv::RNG, fits EPnP to each, counts
// the points whose float reprojection error is <= 400 and stops after RANSACUpdateNumIters iterations;
// then solvePnP(ITERATIVE) refines over the inliers of the best model.
//
// Here all (<= 100) hypotheses of a frame — and of every frame of a batch — are evaluated at once:
//   pnp_epnp_kernel    one WARP per hypothesis: EPnP on 5 correspondences.  The hypothesis is a chain of tiny
//                      dense factorisations (3x3 and 12x12 symmetric eigen-problems by cyclic Jacobi, 6x{3,4,5}
//                      Householder least squares, 5 Gauss-Newton steps x 3 initialisations, a 3x3 absolute
//                      orientation).  The 12x12 eigen-problem is 9/10 of the work: the lanes of the warp share
//                      it in shared memory (lane k = entry k of the rotated rows / columns), lane 0 runs the rest;
//   pnp_score_kernel   one WARP per hypothesis: lanes stride over the points, ballot -> inlier-mask words and
//                      the count (HBM traffic: 20 B per point per hypothesis from L2);
//   (host)             OpenCV's "strictly more inliers wins" / RANSACUpdateNumIters replay over the counts;
//   pnp_refine_kernel  one CTA per frame: Levenberg-Marquardt over the inliers of the winner, normal equations
//                      reduced in a fixed order, and the mask words expanded to the caller's byte flags.
// fp64, compiled with -fmad=false: the arithmetic of a hypothesis follows the CPU restatement
// (the pnp file of the parity oracle) operation by operation, so counts and masks are identical to it.

#include "common.cuh"
#include "kernels.h"

namespace urmvo {

namespace {

// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (row-major, destroyed), eigenvectors as columns
// of V.  Pairs in row order, a pair is rotated when a_pq != 0, <= 60 sweeps, stop at off^2 <= 1e-32 |A|_F^2.
template <int n>
__device__ void jacobi_eig(double* A, double* V) {
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
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll 4
        for (int k = 0; k < n; k++) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
#pragma unroll 4
        for (int k = 0; k < n; k++) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
#pragma unroll 4
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
}

// The 12 x 12 matrix of a hypothesis is diagonalised by the 32 lanes of its warp in shared memory with the PARALLEL
// (round-robin) Jacobi ordering: a sweep is 11 rounds of 6 disjoint pairs (pair 0 of round r is (r, 11); pair
// k = 1..5 is ((r + k) mod 11, (r - k) mod 11)).  Lanes 0..5 compute the six rotations of a round at once — the two
// square roots and three divisions of a rotation are a ~1000-cycle dependency chain, paid 11 times per sweep instead
// of 66 —, then all lanes apply them: first to the columns of A and V, then to the rows of A.  The order of the
// element operations is that of the CPU restatement (its jacobi_eig12_rr).
__device__ void jacobi_eig12_warp(double* A, double* V, double* cs /* [12] */, int lane) {
  constexpr int n = 12;
  for (int e = lane; e < n * n; e += 32) V[e] = (e / n == e % n) ? 1.0 : 0.0;
  __syncwarp();
  double fro = 0;
  for (int i = 0; i < n * n; i++) fro += A[i] * A[i];
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
    if (off <= 1e-32 * fro) break;
    for (int r = 0; r < 11; r++) {
      if (lane < 6) {
        const int k = lane;
        const int a = k == 0 ? r : (r + k) % 11, b = k == 0 ? 11 : (r - k + 11) % 11;
        const int p = a < b ? a : b, q = a < b ? b : a;
        const double apq = A[p * n + q];
        double c = 1.0, s = 0.0;
        if (apq != 0.0) {
          const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
          const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          c = 1.0 / sqrt(t * t + 1.0);
          s = t * c;
        }
        cs[2 * k] = c;
        cs[2 * k + 1] = s;
      }
      __syncwarp();
      for (int item = lane; item < 144; item += 32) {  // columns of A (items 0..71) and of V (72..143)
        const bool isV = item >= 72;
        const int it2 = isV ? item - 72 : item;
        const int k = it2 / 12, i = it2 - k * 12;
        const int a = k == 0 ? r : (r + k) % 11, b = k == 0 ? 11 : (r - k + 11) % 11;
        const int p = a < b ? a : b, q = a < b ? b : a;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        double* Mx = isV ? V : A;
        const double xp = Mx[i * n + p], xq = Mx[i * n + q];
        Mx[i * n + p] = c * xp - s * xq;
        Mx[i * n + q] = s * xp + c * xq;
      }
      __syncwarp();
      for (int item = lane; item < 72; item += 32) {  // rows of A
        const int k = item / 12, i = item - k * 12;
        const int a = k == 0 ? r : (r + k) % 11, b = k == 0 ? 11 : (r - k + 11) % 11;
        const int p = a < b ? a : b, q = a < b ? b : a;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        const double apk = A[p * n + i], aqk = A[q * n + i];
        A[p * n + i] = c * apk - s * aqk;
        A[q * n + i] = s * apk + c * aqk;
      }
      __syncwarp();
    }
  }
}

template <int n>
__device__ void sort_diag(const double* A, int* order, bool descending) {
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

// least squares by Householder QR (epnp::qr_solve); A: 6 x NC row-major, destroyed.  Fully unrolled: the small
// systems live in registers (same operations in the same order as the run-time-sized CPU routine).
template <int NC>
__device__ __forceinline__ bool qr_solve(double* A, double* b, double* X) {
  constexpr int nr = 6, nc = NC;
  double A1[NC], A2[NC];
#pragma unroll
  for (int k = 0; k < nc; k++) {
    double eta = 0;
#pragma unroll
    for (int i = k; i < nr; i++) eta = fmax(eta, fabs(A[i * nc + k]));
    if (eta == 0) return false;
    const double inv_eta = 1.0 / eta;
    double sum2 = 0;
#pragma unroll
    for (int i = k; i < nr; i++) {
      A[i * nc + k] *= inv_eta;
      sum2 += A[i * nc + k] * A[i * nc + k];
    }
    double sigma = sqrt(sum2);
    if (A[k * nc + k] < 0) sigma = -sigma;
    A[k * nc + k] += sigma;
    A1[k] = sigma * A[k * nc + k];
    A2[k] = -eta * sigma;
#pragma unroll
    for (int j = k + 1; j < nc; j++) {
      double sum = 0;
#pragma unroll
      for (int i = k; i < nr; i++) sum += A[i * nc + k] * A[i * nc + j];
      const double tau = sum / A1[k];
#pragma unroll
      for (int i = k; i < nr; i++) A[i * nc + j] -= tau * A[i * nc + k];
    }
  }
#pragma unroll
  for (int j = 0; j < nc; j++) {
    double tau = 0;
#pragma unroll
    for (int i = j; i < nr; i++) tau += A[i * nc + j] * b[i];
    tau /= A1[j];
#pragma unroll
    for (int i = j; i < nr; i++) b[i] -= tau * A[i * nc + j];
  }
  X[nc - 1] = b[nc - 1] / A2[nc - 1];
#pragma unroll
  for (int i = nc - 2; i >= 0; i--) {
    double sum = 0;
#pragma unroll
    for (int j = i + 1; j < nc; j++) sum += A[i * nc + j] * X[j];
    X[i] = (b[i] - sum) / A2[i];
  }
  return true;
}

struct EpnpState {
  double fu, fv, uc, vc;
  double pws[5][3], us[5][2], al[5][4], cws[4][3];
  double ut[4][12];
  double L[6][10], rho[6];
};

// Part 1 (one lane): control points, barycentric coordinates, the 10 x 12 system M (row-major, to shared memory)
__device__ bool epnp_prepare_M(EpnpState& e, double* M) {
  const int n = 5;
  for (int j = 0; j < 3; j++) {
    double s = 0;
    for (int i = 0; i < n; i++) s += e.pws[i][j];
    e.cws[0][j] = s / n;
  }
  {
    double A[9], V[9];
    for (int a = 0; a < 9; a++) A[a] = 0.0;
    for (int i = 0; i < n; i++)
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) A[a * 3 + b] += (e.pws[i][a] - e.cws[0][a]) * (e.pws[i][b] - e.cws[0][b]);
    jacobi_eig<3>(A, V);
    int ord[3];
    sort_diag<3>(A, ord, true);
    for (int i = 1; i < 4; i++) {
      const int o = ord[i - 1];
      const double ev = A[o * 3 + o] > 0 ? A[o * 3 + o] : 0.0;
      const double k = sqrt(ev / n);
      for (int j = 0; j < 3; j++) e.cws[i][j] = e.cws[0][j] + k * V[j * 3 + o];
    }
  }
  double cc[9];
  for (int i = 0; i < 3; i++)
    for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = e.cws[j][i] - e.cws[0][i];
  const double c00 = cc[4] * cc[8] - cc[5] * cc[7], c01 = cc[5] * cc[6] - cc[3] * cc[8], c02 = cc[3] * cc[7] - cc[4] * cc[6];
  const double det = cc[0] * c00 + cc[1] * c01 + cc[2] * c02;
  if (!(fabs(det) > 0)) return false;
  const double id = 1.0 / det;
  double ci[9];
  ci[0] = c00 * id; ci[1] = (cc[2] * cc[7] - cc[1] * cc[8]) * id; ci[2] = (cc[1] * cc[5] - cc[2] * cc[4]) * id;
  ci[3] = c01 * id; ci[4] = (cc[0] * cc[8] - cc[2] * cc[6]) * id; ci[5] = (cc[2] * cc[3] - cc[0] * cc[5]) * id;
  ci[6] = c02 * id; ci[7] = (cc[1] * cc[6] - cc[0] * cc[7]) * id; ci[8] = (cc[0] * cc[4] - cc[1] * cc[3]) * id;
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < 3; j++)
      e.al[i][1 + j] = ci[3 * j] * (e.pws[i][0] - e.cws[0][0]) + ci[3 * j + 1] * (e.pws[i][1] - e.cws[0][1]) +
                       ci[3 * j + 2] * (e.pws[i][2] - e.cws[0][2]);
    e.al[i][0] = 1.0 - e.al[i][1] - e.al[i][2] - e.al[i][3];
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 4; j++) {
      M[(2 * i) * 12 + 3 * j] = e.al[i][j] * e.fu; M[(2 * i) * 12 + 3 * j + 1] = 0.0; M[(2 * i) * 12 + 3 * j + 2] = e.al[i][j] * (e.uc - e.us[i][0]);
      M[(2 * i + 1) * 12 + 3 * j] = 0.0; M[(2 * i + 1) * 12 + 3 * j + 1] = e.al[i][j] * e.fv; M[(2 * i + 1) * 12 + 3 * j + 2] = e.al[i][j] * (e.vc - e.us[i][1]);
    }
  return true;
}

// Part 2 (one lane, after the warp has diagonalised M^T M into (MtM, VV)): the four smallest eigenvectors,
// L_6x10 and rho
__device__ void epnp_prepare_L(EpnpState& e, const double* MtM, const double* VV) {
  {
    int ord12[12];
    sort_diag<12>(MtM, ord12, false);
    for (int i = 0; i < 4; i++)
      for (int k = 0; k < 12; k++) e.ut[i][k] = VV[k * 12 + ord12[i]];
  }
  double dv[4][6][3];
  for (int i = 0; i < 4; i++) {
    int a = 0, b = 1;
    for (int j = 0; j < 6; j++) {
      for (int k = 0; k < 3; k++) dv[i][j][k] = e.ut[i][3 * a + k] - e.ut[i][3 * b + k];
      b++;
      if (b > 3) { a++; b = a + 1; }
    }
  }
  auto dot = [](const double* x, const double* y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
  for (int i = 0; i < 6; i++) {
    e.L[i][0] = dot(dv[0][i], dv[0][i]);
    e.L[i][1] = 2.0 * dot(dv[0][i], dv[1][i]);
    e.L[i][2] = dot(dv[1][i], dv[1][i]);
    e.L[i][3] = 2.0 * dot(dv[0][i], dv[2][i]);
    e.L[i][4] = 2.0 * dot(dv[1][i], dv[2][i]);
    e.L[i][5] = dot(dv[2][i], dv[2][i]);
    e.L[i][6] = 2.0 * dot(dv[0][i], dv[3][i]);
    e.L[i][7] = 2.0 * dot(dv[1][i], dv[3][i]);
    e.L[i][8] = 2.0 * dot(dv[2][i], dv[3][i]);
    e.L[i][9] = dot(dv[3][i], dv[3][i]);
  }
  const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
  for (int i = 0; i < 6; i++) {
    double s = 0;
    for (int k = 0; k < 3; k++) s += (e.cws[pa[i]][k] - e.cws[pb[i]][k]) * (e.cws[pa[i]][k] - e.cws[pb[i]][k]);
    e.rho[i] = s;
  }
}

template <int NC>
__device__ __forceinline__ bool epnp_betas_solve(const EpnpState& e, const int (&cols)[NC], double* x) {
  double A[6 * NC], b[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int j = 0; j < NC; j++) A[i * NC + j] = e.L[i][cols[j]];
    b[i] = e.rho[i];
  }
  return qr_solve<NC>(A, b, x);
}

__device__ bool epnp_betas(const EpnpState& e, int which, double* be) {
  double x[5];
  bool ok;
  if (which == 0) { const int c[4] = {0, 1, 3, 6}; ok = epnp_betas_solve<4>(e, c, x); }
  else if (which == 1) { const int c[3] = {0, 1, 2}; ok = epnp_betas_solve<3>(e, c, x); }
  else { const int c[5] = {0, 1, 2, 3, 4}; ok = epnp_betas_solve<5>(e, c, x); }
  if (!ok) return false;
  be[0] = be[1] = be[2] = be[3] = 0.0;
  if (which == 0) {
    if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = -x[1] / be[0]; be[2] = -x[2] / be[0]; be[3] = -x[3] / be[0]; }
    else { be[0] = sqrt(x[0]); be[1] = x[1] / be[0]; be[2] = x[2] / be[0]; be[3] = x[3] / be[0]; }
  } else {
    if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = x[2] < 0 ? sqrt(-x[2]) : 0.0; }
    else { be[0] = sqrt(x[0]); be[1] = x[2] > 0 ? sqrt(x[2]) : 0.0; }
    if (x[1] < 0) be[0] = -be[0];
    if (which == 2) be[2] = x[3] / be[0];
  }
  return true;
}

__device__ bool epnp_gauss_newton(const EpnpState& e, double* be) {
  for (int it = 0; it < 5; it++) {
    double A[24], b[6], x[4];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const double* r = e.L[i];
      A[i * 4 + 0] = 2 * r[0] * be[0] + r[1] * be[1] + r[3] * be[2] + r[6] * be[3];
      A[i * 4 + 1] = r[1] * be[0] + 2 * r[2] * be[1] + r[4] * be[2] + r[7] * be[3];
      A[i * 4 + 2] = r[3] * be[0] + r[4] * be[1] + 2 * r[5] * be[2] + r[8] * be[3];
      A[i * 4 + 3] = r[6] * be[0] + r[7] * be[1] + r[8] * be[2] + 2 * r[9] * be[3];
      b[i] = e.rho[i] - (r[0] * be[0] * be[0] + r[1] * be[0] * be[1] + r[2] * be[1] * be[1] + r[3] * be[0] * be[2] +
                         r[4] * be[1] * be[2] + r[5] * be[2] * be[2] + r[6] * be[0] * be[3] + r[7] * be[1] * be[3] +
                         r[8] * be[2] * be[3] + r[9] * be[3] * be[3]);
    }
    if (!qr_solve<4>(A, b, x)) return false;
    for (int i = 0; i < 4; i++) be[i] += x[i];
  }
  return true;
}

__device__ double epnp_R_and_t(const EpnpState& e, const double* be, double* R, double* t) {
  const int n = 5;
  double ccs[4][3], pcs[5][3];
  for (int j = 0; j < 4; j++)
    for (int k = 0; k < 3; k++) ccs[j][k] = 0.0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 3; k++) ccs[j][k] += be[i] * e.ut[i][3 * j + k];
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 3; k++)
      pcs[i][k] = e.al[i][0] * ccs[0][k] + e.al[i][1] * ccs[1][k] + e.al[i][2] * ccs[2][k] + e.al[i][3] * ccs[3][k];
  if (pcs[0][2] < 0.0)
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) pcs[i][k] = -pcs[i][k];
  double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) { pc0[k] += pcs[i][k]; pw0[k] += e.pws[i][k]; }
  for (int k = 0; k < 3; k++) { pc0[k] /= n; pw0[k] /= n; }
  double B[9];
  for (int a = 0; a < 9; a++) B[a] = 0.0;
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) B[a * 3 + b] += (pcs[i][a] - pc0[a]) * (e.pws[i][b] - pw0[b]);
  double BtB[9], V[9];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) BtB[a * 3 + b] = B[0 * 3 + a] * B[0 * 3 + b] + B[1 * 3 + a] * B[1 * 3 + b] + B[2 * 3 + a] * B[2 * 3 + b];
  jacobi_eig<3>(BtB, V);
  int ord[3];
  sort_diag<3>(BtB, ord, true);
  double Vs[3][3], Us[3][3], sv[3];
  for (int c = 0; c < 3; c++) {
    const int o = ord[c];
    sv[c] = sqrt(BtB[o * 3 + o] > 0 ? BtB[o * 3 + o] : 0.0);
    for (int r = 0; r < 3; r++) Vs[r][c] = V[r * 3 + o];
  }
  for (int c = 0; c < 3; c++) {
    if (c < 2 || sv[2] > 1e-9 * sv[0]) {
      double u[3], nn = 0;
      for (int r = 0; r < 3; r++) {
        u[r] = B[r * 3 + 0] * Vs[0][c] + B[r * 3 + 1] * Vs[1][c] + B[r * 3 + 2] * Vs[2][c];
        nn += u[r] * u[r];
      }
      nn = sqrt(nn);
      if (!(nn > 0)) return 1e300;
      for (int r = 0; r < 3; r++) Us[r][c] = u[r] / nn;
    } else {
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
    const double Xc = R[0] * e.pws[i][0] + R[1] * e.pws[i][1] + R[2] * e.pws[i][2] + t[0];
    const double Yc = R[3] * e.pws[i][0] + R[4] * e.pws[i][1] + R[5] * e.pws[i][2] + t[1];
    const double iz = 1.0 / (R[6] * e.pws[i][0] + R[7] * e.pws[i][1] + R[8] * e.pws[i][2] + t[2]);
    const double ue = e.uc + e.fu * Xc * iz, ve = e.vc + e.fv * Yc * iz;
    sum += sqrt((e.us[i][0] - ue) * (e.us[i][0] - ue) + (e.us[i][1] - ve) * (e.us[i][1] - ve));
  }
  return sum / n;
}

// hypothesis h = (problem b, iteration it): sets [H][5] point indices local to the problem, models [H][12] = R | t.
// One WARP per hypothesis: lane 0 runs the scalar chain, the 12 x 12 eigen-problem (9/10 of the work) is shared.
constexpr int kEpnpWarps = 4;
__global__ void __launch_bounds__(kEpnpWarps * 32)
pnp_epnp_kernel(int H, int max_iters, const int* __restrict__ off, const int* __restrict__ sets,
                const float* __restrict__ obj, const float* __restrict__ img, const double* __restrict__ K4,
                double* __restrict__ models, int* __restrict__ valid) {
  __shared__ double s_M[kEpnpWarps][120], s_A[kEpnpWarps][144], s_V[kEpnpWarps][144];
  __shared__ EpnpState s_e[kEpnpWarps];
  __shared__ double s_cs[kEpnpWarps][12];
  __shared__ int s_ok[kEpnpWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x * kEpnpWarps + warp;
  if (h >= H) return;
  EpnpState& e = s_e[warp];
  if (lane == 0) {
    const int b = h / max_iters;
    const int o0 = off[b];
    e.fu = K4[0]; e.fv = K4[1]; e.uc = K4[2]; e.vc = K4[3];
    const double ifx = 1.0 / K4[0], ify = 1.0 / K4[1];
    for (int i = 0; i < 5; i++) {
      const int p = o0 + sets[h * 5 + i];
      for (int k = 0; k < 3; k++) e.pws[i][k] = obj[3 * p + k];
      const float xn = (float)(((double)img[2 * p] - K4[2]) * ifx), yn = (float)(((double)img[2 * p + 1] - K4[3]) * ify);
      e.us[i][0] = (double)xn * K4[0] + K4[2];
      e.us[i][1] = (double)yn * K4[1] + K4[3];
    }
    s_ok[warp] = epnp_prepare_M(e, s_M[warp]) ? 1 : 0;
  }
  __syncwarp();
  const bool ok = s_ok[warp] != 0;
  if (ok) {
    for (int en = lane; en < 144; en += 32) {  // M^T M, every entry summed over the rows in order
      const int a = en / 12, b = en - a * 12;
      double sm = 0;
      for (int r = 0; r < 10; r++) sm += s_M[warp][r * 12 + a] * s_M[warp][r * 12 + b];
      s_A[warp][en] = sm;
    }
    __syncwarp();
    jacobi_eig12_warp(s_A[warp], s_V[warp], s_cs[warp], lane);
    __syncwarp();
    if (lane == 0) epnp_prepare_L(e, s_A[warp], s_V[warp]);
    __syncwarp();
  }
  // the three beta initialisations run on lanes 0..2 side by side (Gauss-Newton and the absolute orientation are
  // the same code on different data); lane 0 then keeps the candidate OpenCV's sequential comparison keeps
  double err = 1e300, Rc[9], tc[3];
  bool have = false;
  if (ok && lane < 3) {
    double be[4];
    if (epnp_betas(e, lane, be) && epnp_gauss_newton(e, be)) {
      err = epnp_R_and_t(e, be, Rc, tc);
      have = err == err;
    }
  }
  bool any = false;
  double best = 1e300;
  int pick = 0;
  for (int which = 0; which < 3; which++) {
    const bool hv = __shfl_sync(0xffffffffu, have ? 1 : 0, which) != 0;
    const double ev = __shfl_sync(0xffffffffu, err, which);
    if (hv && (!any || ev < best)) { best = ev; pick = which; any = true; }
  }
  if (lane == 0) valid[h] = any ? 1 : 0;
  if (any && lane == pick) {
    for (int a = 0; a < 9; a++) models[h * 12 + a] = Rc[a];
    for (int a = 0; a < 3; a++) models[h * 12 + 9 + a] = tc[a];
  }
}

// PnPRansacCallback::computeError + findInliers: one warp per hypothesis
__global__ void __launch_bounds__(256)
pnp_score_kernel(int H, int max_iters, int words_max, const int* __restrict__ off, const float* __restrict__ obj,
                 const float* __restrict__ img, const double* __restrict__ K4, const double* __restrict__ models,
                 const int* __restrict__ valid, float thr2, unsigned* __restrict__ masks, int* __restrict__ counts) {
  const int h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (h >= H) return;
  if (!valid[h]) { if (lane == 0) counts[h] = -1; return; }
  const int b = h / max_iters, o0 = off[b], N = off[b + 1] - o0;
  double R[9], t[3];
#pragma unroll
  for (int a = 0; a < 9; a++) R[a] = models[h * 12 + a];
#pragma unroll
  for (int a = 0; a < 3; a++) t[a] = models[h * 12 + 9 + a];
  const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
  int cnt = 0;
  for (int i0 = 0; i0 < N; i0 += 32) {
    const int i = i0 + lane;
    bool in = false;
    if (i < N) {
      const float* X = obj + 3 * (size_t)(o0 + i);
      const float* uv = img + 2 * (size_t)(o0 + i);
      const double X0 = X[0], X1 = X[1], X2 = X[2];
      const double x = R[0] * X0 + R[1] * X1 + R[2] * X2 + t[0];
      const double y = R[3] * X0 + R[4] * X1 + R[5] * X2 + t[1];
      double z = R[6] * X0 + R[7] * X1 + R[8] * X2 + t[2];
      z = z != 0.0 ? 1.0 / z : 1.0;
      const float pu = (float)(x * z * fx + cx), pv = (float)(y * z * fy + cy);
      const float dx = uv[0] - pu, dy = uv[1] - pv;
      float s = dx * dx;
      s += dy * dy;
      in = s <= thr2;
    }
    const unsigned w = __ballot_sync(0xffffffffu, in);
    if (lane == 0) masks[(size_t)h * words_max + (i0 >> 5)] = w;
    cnt += __popc(w);
  }
  if (lane == 0) counts[h] = cnt;
}

__device__ __forceinline__ void so3_exp(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double a, b;
  if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
  else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double w2 = 0;
      for (int k = 0; k < 3; k++) w2 += W[i * 3 + k] * W[k * 3 + j];
      R[i * 3 + j] = (i == j ? 1.0 : 0.0) + a * W[i * 3 + j] + b * w2;
    }
}

constexpr int kRefThreads = 256;
constexpr int kRefVals = 28;  // 21 (upper triangle of J^T J) + 6 (J^T e) + 1 (cost)

// One CTA per problem: LM over the inliers of hypothesis best[b] (solvePnP(ITERATIVE) of solvePnPRansac's last step)
__global__ void __launch_bounds__(kRefThreads)
pnp_refine_kernel(int max_iters, int words_max, const int* __restrict__ off, const float* __restrict__ obj,
                  const float* __restrict__ img, const double* __restrict__ K4, const double* __restrict__ models,
                  const unsigned* __restrict__ masks, const int* __restrict__ best, double* __restrict__ out_Rt,
                  uint8_t* __restrict__ inlier) {
  __shared__ double s_part[kRefThreads / 32][kRefVals];
  __shared__ double s_sum[kRefVals];
  __shared__ double s_R[9], s_t[3], s_Rn[9], s_tn[3];
  __shared__ int s_state;  // 0: continue, 1: done
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int o0 = off[b], N = off[b + 1] - o0;
  const int hb = best[b];
  if (hb < 0) {  // no model: all flags 0
    for (int i = tid; i < N; i += kRefThreads) inlier[o0 + i] = 0;
    return;
  }
  const unsigned* mw = masks + (size_t)(b * max_iters + hb) * words_max;
  for (int i = tid; i < N; i += kRefThreads) inlier[o0 + i] = (mw[i >> 5] >> (i & 31)) & 1u;
  if (tid < 9) s_Rn[tid] = models[(size_t)(b * max_iters + hb) * 12 + tid];
  if (tid < 3) s_tn[tid] = models[(size_t)(b * max_iters + hb) * 12 + 9 + tid];
  if (tid == 0) s_state = 0;
  __syncthreads();
  const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
  double lambda = 1e-3, cost = 0.0, step = 0.0, H[36], g[6];
  bool first = true;
  // every pass evaluates cost AND normal equations at the candidate (s_Rn, s_tn); thread 0 accepts / rejects
  for (int it = 0; it < 101; it++) {
    double acc[kRefVals];
#pragma unroll
    for (int a = 0; a < kRefVals; a++) acc[a] = 0.0;
    double R[9], t[3];
#pragma unroll
    for (int a = 0; a < 9; a++) R[a] = s_Rn[a];
#pragma unroll
    for (int a = 0; a < 3; a++) t[a] = s_tn[a];
    for (int i = tid; i < N; i += kRefThreads) {
      if (!((mw[i >> 5] >> (i & 31)) & 1u)) continue;
      const float* Xf = obj + 3 * (size_t)(o0 + i);
      const double X = Xf[0], Y = Xf[1], Z = Xf[2];
      const double x = R[0] * X + R[1] * Y + R[2] * Z + t[0], y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
      const double z = R[6] * X + R[7] * Y + R[8] * Z + t[2], iz = 1.0 / z;
      const double e0 = (double)img[2 * (size_t)(o0 + i)] - (fx * x * iz + cx), e1 = (double)img[2 * (size_t)(o0 + i) + 1] - (fy * y * iz + cy);
      const double a = fx * iz, bb = fy * iz, xz = x * iz, yz = y * iz;
      const double J0[6] = {-a * xz * y, a * (z + x * xz), -a * y, a, 0, -a * xz};
      const double J1[6] = {-bb * (z + y * yz), bb * yz * x, bb * x, 0, bb, -bb * yz};
      int k = 0;
#pragma unroll
      for (int p = 0; p < 6; p++)
#pragma unroll
        for (int q = p; q < 6; q++) acc[k++] += J0[p] * J0[q] + J1[p] * J1[q];
#pragma unroll
      for (int p = 0; p < 6; p++) acc[21 + p] += J0[p] * e0 + J1[p] * e1;
      acc[27] += e0 * e0 + e1 * e1;
    }
#pragma unroll
    for (int a = 0; a < kRefVals; a++) {
      double v = acc[a];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_part[warp][a] = v;
    }
    __syncthreads();
    if (tid < kRefVals) {
      double v = 0;
      for (int w = 0; w < kRefThreads / 32; w++) v += s_part[w][tid];
      s_sum[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
      const double cn = s_sum[27];
      bool stop = false, accept = false;
      if (first) {
        accept = true;
        first = false;
      } else if (cn <= cost) {
        accept = true;
        lambda = fmax(lambda * 0.1, 1e-12);
        if (step < 1e-13) stop = true;
      } else {
        lambda *= 10;
        if (lambda > 1e16 || step < 1e-13) stop = true;
      }
      if (accept) {
        for (int a = 0; a < 9; a++) s_R[a] = s_Rn[a];
        for (int a = 0; a < 3; a++) s_t[a] = s_tn[a];
        cost = cn;
        int k = 0;
        for (int p = 0; p < 6; p++)
          for (int q = p; q < 6; q++) { H[p * 6 + q] = s_sum[k]; H[q * 6 + p] = s_sum[k]; k++; }
        for (int p = 0; p < 6; p++) g[p] = s_sum[21 + p];
      }
      if (it >= 100) stop = true;
      while (!stop) {  // damped solve (H + lambda diag H) x = g at the current point -> next candidate
        double A[36], Lc[36], x[6], y[6];
        for (int a = 0; a < 36; a++) { A[a] = H[a]; Lc[a] = 0.0; }
        for (int p = 0; p < 6; p++) A[p * 6 + p] *= 1.0 + lambda;
        bool ok = true;
        for (int i = 0; i < 6 && ok; i++)
          for (int j = 0; j <= i; j++) {
            double sm = A[i * 6 + j];
            for (int k = 0; k < j; k++) sm -= Lc[i * 6 + k] * Lc[j * 6 + k];
            if (i == j) { if (!(sm > 0)) { ok = false; break; } Lc[i * 6 + i] = sqrt(sm); }
            else Lc[i * 6 + j] = sm / Lc[j * 6 + j];
          }
        if (!ok) { lambda *= 10; if (lambda > 1e16) stop = true; continue; }
        for (int i = 0; i < 6; i++) { double sm = g[i]; for (int k = 0; k < i; k++) sm -= Lc[i * 6 + k] * y[k]; y[i] = sm / Lc[i * 6 + i]; }
        for (int i = 5; i >= 0; i--) { double sm = y[i]; for (int k = i + 1; k < 6; k++) sm -= Lc[k * 6 + i] * x[k]; x[i] = sm / Lc[i * 6 + i]; }
        double dR[9];
        so3_exp(x, dR);
        for (int i = 0; i < 3; i++) {
          for (int j = 0; j < 3; j++) s_Rn[i * 3 + j] = dR[i * 3] * s_R[j] + dR[i * 3 + 1] * s_R[3 + j] + dR[i * 3 + 2] * s_R[6 + j];
          s_tn[i] = dR[i * 3] * s_t[0] + dR[i * 3 + 1] * s_t[1] + dR[i * 3 + 2] * s_t[2] + x[3 + i];
        }
        step = 0;
        for (int p = 0; p < 6; p++) step = fmax(step, fabs(x[p]));
        break;
      }
      if (stop) s_state = 1;
    }
    __syncthreads();
    if (s_state) break;
  }
  if (tid < 9) out_Rt[b * 12 + tid] = s_R[tid];
  if (tid < 3) out_Rt[b * 12 + 9 + tid] = s_t[tid];
}

}  // namespace

cudaError_t launch_pnp_hypotheses(int H, int max_iters, int words_max, const int* off, const int* sets, const float* obj,
                                  const float* img, const double* K4, double* models, int* valid, float thr2,
                                  unsigned* masks, int* counts, cudaStream_t s) {
  pnp_epnp_kernel<<<(H + kEpnpWarps - 1) / kEpnpWarps, kEpnpWarps * 32, 0, s>>>(H, max_iters, off, sets, obj, img, K4, models, valid);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  pnp_score_kernel<<<(H + 7) / 8, 256, 0, s>>>(H, max_iters, words_max, off, obj, img, K4, models, valid, thr2, masks, counts);
  return cudaGetLastError();
}

cudaError_t launch_pnp_refine(int B, int max_iters, int words_max, const int* off, const float* obj, const float* img,
                              const double* K4, const double* models, const unsigned* masks, const int* best,
                              double* out_Rt, uint8_t* inlier, cudaStream_t s) {
  pnp_refine_kernel<<<B, kRefThreads, 0, s>>>(max_iters, words_max, off, obj, img, K4, models, masks, best, out_Rt, inlier);
  return cudaGetLastError();
}

}  // namespace urmvo
