// oracle/twoview_oracle.cpp — CPU restatement of the reference's two-view RANSAC / reconstruction.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED (no reference tests/fixtures exist).
// Follows /root/reference/src/epipolar_geometry.cc function by function (lines cited below).
// All arithmetic is fp32 like the reference.  Must be compiled with -ffp-contract=off: the CUDA
// kernels are compiled with -fmad=false and the parity contract is bit-exact masks/scores for
// identical 8-point sets.
//
// Eigen::JacobiSVD (Eigen 3.3.7, absent here) cannot be reproduced bit for bit; it is replaced by
// a fully specified one-sided (Hestenes) Jacobi SVD, cyclic-by-rows pair order, a pair is left alone
// when g^2 <= (5e-7)^2 a b, rotation t = 2g / (d +- sqrt(d^2 + 4 g^2)) with d = b - a, at most 30
// sweeps; the null vector of the 8x9 fundamental system comes from a Householder QR instead.  All uses in the reference are invariant to the sign /
// ordering conventions of the SVD (SURVEY.md §7 "RANSAC bit-exactness").

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "oracle.h"

namespace {

constexpr int kMaxSweeps = 30;
constexpr float kJacobiTol = 5e-7f;
constexpr float kJacobiTol2 = kJacobiTol * kJacobiTol;  // the test is gamma^2 <= tol^2 alpha beta (no square root)
// A column whose squared norm is below kJacobiTiny * ||A||_F^2 is numerically zero (the null column
// of a rank-deficient DLT system reaches ~1e-20 after a few sweeps); rotating it against the other
// columns only chases round-off and would keep every solve at the sweep limit.
constexpr float kJacobiTiny = 1e-14f;

// Sum over the m rows of a design matrix.  Part of the specification: for m < 8 (the 3x3 and 4x4
// problems, solved by one thread) the terms are added in row order; for m >= 8 (the 8x9 / 16x9
// systems of a hypothesis, whose rows live in the lanes of a sub-warp on the GPU) they are added as
// a pairwise tree over 8 or 16 terms (zero-padded) — the order of an xor-butterfly of shuffles.
template <class F>
inline float row_sum(int m, F term) {
  if (m < 8) {
    float s = 0.0f;
    for (int k = 0; k < m; k++) s = (k == 0) ? term(k) : s + term(k);
    return s;
  }
  const int P = m <= 8 ? 8 : 16;
  float t[16];
  for (int k = 0; k < P; k++) t[k] = k < m ? term(k) : 0.0f;
  for (int w = 1; w < P; w <<= 1)
    for (int i = 0; i < P; i += 2 * w) t[i] = t[i] + t[i + w];
  return t[0];
}

// One-sided Jacobi: A (m x n, row-major, leading dim n) becomes U*Sigma, V (n x n) accumulates.
void jacobi_onesided(int m, int n, float* A, float* V) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0f : 0.0f;
  // sum over rows of the row sums (this order is part of the specification)
  const float fro2 = row_sum(m, [&](int k) {
    float row = 0.0f;
    for (int j = 0; j < n; j++) row += A[k * n + j] * A[k * n + j];
    return row;
  });
  const float tiny = kJacobiTiny * fro2;
  for (int sweep = 0; sweep < kMaxSweeps; sweep++) {
    bool rotated = false;
    for (int p = 0; p < n - 1; p++) {
      for (int q = p + 1; q < n; q++) {
        const float alpha = row_sum(m, [&](int k) { return A[k * n + p] * A[k * n + p]; });
        const float beta = row_sum(m, [&](int k) { return A[k * n + q] * A[k * n + q]; });
        const float gamma = row_sum(m, [&](int k) { return A[k * n + p] * A[k * n + q]; });
        if (alpha <= tiny || beta <= tiny) continue;
        if (gamma * gamma <= kJacobiTol2 * (alpha * beta)) continue;
        rotated = true;
        const float dlt = beta - alpha;
        const float rad = std::sqrt(dlt * dlt + 4.0f * (gamma * gamma));
        const float t = (2.0f * gamma) / (dlt >= 0.0f ? dlt + rad : dlt - rad);
        float c = 1.0f / std::sqrt(1.0f + t * t);
        float s = c * t;
        for (int k = 0; k < m; k++) {
          float ap = A[k * n + p], aq = A[k * n + q];
          A[k * n + p] = c * ap - s * aq;
          A[k * n + q] = s * ap + c * aq;
        }
        for (int k = 0; k < n; k++) {
          float vp = V[k * n + p], vq = V[k * n + q];
          V[k * n + p] = c * vp - s * vq;
          V[k * n + q] = s * vp + c * vq;
        }
      }
    }
    if (!rotated) break;
  }
}

inline float col_norm(int m, int n, const float* A, int j) {
  return std::sqrt(row_sum(m, [&](int k) { return A[k * n + j] * A[k * n + j]; }));
}

// Null vector of the 8x9 system of an 8-point fundamental fit (what V.col(8) of the reference's
// JacobiSVD is, up to sign): Householder QR of M = A^T (9x8), Q = H_0 ... H_7, null vector = Q e_8.
// Specification shared with the CUDA kernel (lane r of a 16-lane group holds row r of M, rows 9..15
// are zero): reflector k uses x = M[k.., k], sigma = row_sum16(x^2), norm = sqrt(sigma),
// alpha = -sign(x_k) norm, v = x - alpha e_k, beta = 1 / (norm (norm + |x_k|)) (0 if norm == 0),
// M[:, j] -= (beta * row_sum16(v . M[:, j])) * v; row_sum16 is the pairwise tree of row_sum(16, .).
// 8 square roots and 8 divisions instead of ~200 Jacobi rotations.
void null_vector_qr_8x9(const float* A, float* v) {
  float M[16][8];
  for (int r = 0; r < 16; r++)
    for (int k = 0; k < 8; k++) M[r][k] = r < 9 ? A[k * 9 + r] : 0.0f;
  float V[8][16], beta[8];
  for (int k = 0; k < 8; k++) {
    float x[16];
    for (int r = 0; r < 16; r++) x[r] = r >= k ? M[r][k] : 0.0f;
    const float sigma = row_sum(16, [&](int r) { return x[r] * x[r]; });
    const float xkk = M[k][k];
    const float norm = std::sqrt(sigma);
    const float alpha = xkk >= 0.0f ? -norm : norm;
    for (int r = 0; r < 16; r++) V[k][r] = r == k ? x[r] - alpha : x[r];
    beta[k] = norm > 0.0f ? 1.0f / (norm * (norm + std::fabs(xkk))) : 0.0f;
    for (int j = k + 1; j < 8; j++) {
      const float w = beta[k] * row_sum(16, [&](int r) { return V[k][r] * M[r][j]; });
      for (int r = 0; r < 16; r++) M[r][j] = M[r][j] - w * V[k][r];
    }
  }
  float y[16];
  for (int r = 0; r < 16; r++) y[r] = r == 8 ? 1.0f : 0.0f;
  for (int k = 7; k >= 0; k--) {
    const float w = beta[k] * row_sum(16, [&](int r) { return V[k][r] * y[r]; });
    for (int r = 0; r < 16; r++) y[r] = y[r] - w * V[k][r];
  }
  for (int r = 0; r < 9; r++) v[r] = y[r];
}

// Right singular vector of the smallest singular value (JacobiSVD::matrixV().col(n-1)).
void null_vector(int m, int n, float* A, float* v) {
  float V[81];
  jacobi_onesided(m, n, A, V);
  int best = 0;
  float bn = col_norm(m, n, A, 0);
  for (int j = 1; j < n; j++) {
    float nj = col_norm(m, n, A, j);
    if (nj < bn) { bn = nj; best = j; }
  }
  for (int k = 0; k < n; k++) v[k] = V[k * n + best];
}

// Full 3x3 SVD, singular values descending. U,V row-major with singular vectors in columns.
// U.col(2) = +-(u0 x u1), sign chosen so that U*diag(w)*V^T reproduces A.
void svd3(const float* Ain, float* U, float* w, float* V) {
  float A[9], Vt[9];
  std::memcpy(A, Ain, sizeof(A));
  jacobi_onesided(3, 3, A, Vt);
  float nrm[3];
  int idx[3] = {0, 1, 2};
  for (int j = 0; j < 3; j++) nrm[j] = col_norm(3, 3, A, j);
  for (int i = 0; i < 2; i++) {  // selection sort, descending, first index wins ties
    int b = i;
    for (int j = i + 1; j < 3; j++)
      if (nrm[idx[j]] > nrm[idx[b]]) b = j;
    int tmp = idx[i]; idx[i] = idx[b]; idx[b] = tmp;
  }
  for (int j = 0; j < 3; j++) {
    int s = idx[j];
    w[j] = nrm[s];
    for (int k = 0; k < 3; k++) V[k * 3 + j] = Vt[k * 3 + s];
    if (j < 2) {
      for (int k = 0; k < 3; k++) U[k * 3 + j] = (nrm[s] > 0.0f) ? A[k * 3 + s] / nrm[s] : 0.0f;
    }
  }
  float c0 = U[3 + 0] * U[6 + 1] - U[6 + 0] * U[3 + 1];
  float c1 = U[6 + 0] * U[0 + 1] - U[0 + 0] * U[6 + 1];
  float c2 = U[0 + 0] * U[3 + 1] - U[3 + 0] * U[0 + 1];
  int s2 = idx[2];
  float d = c0 * A[0 + s2] + c1 * A[3 + s2] + c2 * A[6 + s2];
  if (d < 0.0f) { c0 = -c0; c1 = -c1; c2 = -c2; }
  U[2] = c0; U[5] = c1; U[8] = c2;
}

inline void mat3_mul(const float* A, const float* B, float* C) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
inline void mat3_T(const float* A, float* B) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[i * 3 + j] = A[j * 3 + i];
}
inline float det3(const float* a) {
  return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
         a[2] * (a[3] * a[7] - a[4] * a[6]);
}
inline void inv3f(const float* a, float* r) {
  float c00 = a[4] * a[8] - a[5] * a[7];
  float c01 = a[5] * a[6] - a[3] * a[8];
  float c02 = a[3] * a[7] - a[4] * a[6];
  float det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  float id = 1.0f / det;
  r[0] = c00 * id;
  r[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  r[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  r[3] = c01 * id;
  r[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  r[6] = c02 * id;
  r[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

struct TwoView {
  int n1, n2, N;
  const float* keys1;
  const float* keys2;
  std::vector<int> m1, m2;  // _vMatches12 (first, second)
  float K[9];
  float sigma, sigma2;
  int n_hyp;
  const int32_t* sets;
  int score_mode = 0;  // 0: the reference's symmetric point-line chi2 (:372-449); 1: Sampson error (extra mode)
  // per-image normalisation (:735-780)
  std::vector<float> pn1, pn2;
  float T1[9], T2[9];
};

// _normalize (:735-780): statistics over ALL keypoints of the image.
void normalize(int n, const float* keys, std::vector<float>& pn, float* T) {
  float meanX = 0, meanY = 0;
  pn.resize((size_t)n * 2);
  for (int i = 0; i < n; i++) { meanX += keys[i * 2]; meanY += keys[i * 2 + 1]; }
  meanX = meanX / n;
  meanY = meanY / n;
  float meanDevX = 0, meanDevY = 0;
  for (int i = 0; i < n; i++) {
    pn[i * 2] = keys[i * 2] - meanX;
    pn[i * 2 + 1] = keys[i * 2 + 1] - meanY;
    meanDevX += std::fabs(pn[i * 2]);
    meanDevY += std::fabs(pn[i * 2 + 1]);
  }
  meanDevX = meanDevX / n;
  meanDevY = meanDevY / n;
  float sX = 1.0f / meanDevX;
  float sY = 1.0f / meanDevY;
  for (int i = 0; i < n; i++) { pn[i * 2] = pn[i * 2] * sX; pn[i * 2 + 1] = pn[i * 2 + 1] * sY; }
  for (int i = 0; i < 9; i++) T[i] = 0.0f;
  T[0] = sX; T[4] = sY; T[2] = -meanX * sX; T[5] = -meanY * sY; T[8] = 1.0f;
}

// _compute_F21 (:247-283), then F21i = T2^T * Fn * T1 (:195)
void fit_F(const TwoView& tv, const int32_t* set, float* F21) {
  float A[8 * 9];
  for (int j = 0; j < 8; j++) {
    int idx = set[j];
    float u1 = tv.pn1[tv.m1[idx] * 2], v1 = tv.pn1[tv.m1[idx] * 2 + 1];
    float u2 = tv.pn2[tv.m2[idx] * 2], v2 = tv.pn2[tv.m2[idx] * 2 + 1];
    float* r = A + j * 9;
    r[0] = u2 * u1; r[1] = u2 * v1; r[2] = u2;
    r[3] = v2 * u1; r[4] = v2 * v1; r[5] = v2;
    r[6] = u1; r[7] = v1; r[8] = 1.0f;
  }
  float Fpre[9];
  null_vector_qr_8x9(A, Fpre);
  float U[9], w[3], V[9];
  svd3(Fpre, U, w, V);
  w[2] = 0.0f;
  float Fn[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      Fn[i * 3 + j] = (U[i * 3] * w[0]) * V[j * 3] + (U[i * 3 + 1] * w[1]) * V[j * 3 + 1] +
                      (U[i * 3 + 2] * w[2]) * V[j * 3 + 2];
  float T2t[9], tmp[9];
  mat3_T(tv.T2, T2t);
  mat3_mul(T2t, Fn, tmp);
  mat3_mul(tmp, tv.T1, F21);
}

// _compute_H21 (:207-245), then H21i = T2inv * Hn * T1, H12i = H21i^-1 (:147-149)
void fit_H(const TwoView& tv, const int32_t* set, float* H21, float* H12) {
  float A[16 * 9];
  for (int j = 0; j < 8; j++) {
    int idx = set[j];
    float u1 = tv.pn1[tv.m1[idx] * 2], v1 = tv.pn1[tv.m1[idx] * 2 + 1];
    float u2 = tv.pn2[tv.m2[idx] * 2], v2 = tv.pn2[tv.m2[idx] * 2 + 1];
    float* r0 = A + (2 * j) * 9;
    float* r1 = A + (2 * j + 1) * 9;
    r0[0] = 0.0f; r0[1] = 0.0f; r0[2] = 0.0f;
    r0[3] = -u1; r0[4] = -v1; r0[5] = -1.0f;
    r0[6] = v2 * u1; r0[7] = v2 * v1; r0[8] = v2;
    r1[0] = u1; r1[1] = v1; r1[2] = 1.0f;
    r1[3] = 0.0f; r1[4] = 0.0f; r1[5] = 0.0f;
    r1[6] = -u2 * u1; r1[7] = -u2 * v1; r1[8] = -u2;
  }
  float Hn[9];
  null_vector(16, 9, A, Hn);
  float T2inv[9], tmp[9];
  inv3f(tv.T2, T2inv);
  mat3_mul(T2inv, Hn, tmp);
  mat3_mul(tmp, tv.T1, H21);
  inv3f(H21, H12);
}

// Sampson-error variant of _check_F (BASELINE.json north_star (4); NOT what the reference computes):
//   d = (x2^T F x1)^2 / ((F x1)_1^2 + (F x1)_2^2 + (F^T x2)_1^2 + (F^T x2)_2^2), chi2 = d / sigma^2,
// one test per match against the 1-dof threshold 3.841, score += 5.991 - chi2 for an inlier.  Same
// operation order as the CUDA kernel (fp32, no contraction).
float check_F_sampson(const TwoView& tv, const float* F, uint8_t* mask) {
  const float f11 = F[0], f12 = F[1], f13 = F[2], f21 = F[3], f22 = F[4], f23 = F[5],
              f31 = F[6], f32 = F[7], f33 = F[8];
  float score = 0;
  const float th = 3.841;
  const float thScore = 5.991;
  const float invSigmaSquare = 1.0 / (tv.sigma * tv.sigma);
  for (int i = 0; i < tv.N; i++) {
    const float u1 = tv.keys1[tv.m1[i] * 2], v1 = tv.keys1[tv.m1[i] * 2 + 1];
    const float u2 = tv.keys2[tv.m2[i] * 2], v2 = tv.keys2[tv.m2[i] * 2 + 1];
    const float a2 = f11 * u1 + f12 * v1 + f13;
    const float b2 = f21 * u1 + f22 * v1 + f23;
    const float c2 = f31 * u1 + f32 * v1 + f33;
    const float num2 = a2 * u2 + b2 * v2 + c2;
    const float a1 = f11 * u2 + f21 * v2 + f31;
    const float b1 = f12 * u2 + f22 * v2 + f32;
    const float den = (a2 * a2 + b2 * b2) + (a1 * a1 + b1 * b1);
    const float sampson = num2 * num2 / den;
    const float chiSquare = sampson * invSigmaSquare;
    bool bIn = true;
    if (chiSquare > th) bIn = false; else score += thScore - chiSquare;
    mask[i] = bIn ? 1 : 0;
  }
  return score;
}

// _check_F (:372-449). mask: N bytes.
float check_F(const TwoView& tv, const float* F, uint8_t* mask) {
  if (tv.score_mode == 1) return check_F_sampson(tv, F, mask);
  const float f11 = F[0], f12 = F[1], f13 = F[2], f21 = F[3], f22 = F[4], f23 = F[5],
              f31 = F[6], f32 = F[7], f33 = F[8];
  float score = 0;
  const float th = 3.841;
  const float thScore = 5.991;
  const float invSigmaSquare = 1.0 / (tv.sigma * tv.sigma);
  for (int i = 0; i < tv.N; i++) {
    bool bIn = true;
    const float u1 = tv.keys1[tv.m1[i] * 2], v1 = tv.keys1[tv.m1[i] * 2 + 1];
    const float u2 = tv.keys2[tv.m2[i] * 2], v2 = tv.keys2[tv.m2[i] * 2 + 1];
    const float a2 = f11 * u1 + f12 * v1 + f13;
    const float b2 = f21 * u1 + f22 * v1 + f23;
    const float c2 = f31 * u1 + f32 * v1 + f33;
    const float num2 = a2 * u2 + b2 * v2 + c2;
    const float squareDist1 = num2 * num2 / (a2 * a2 + b2 * b2);
    const float chiSquare1 = squareDist1 * invSigmaSquare;
    if (chiSquare1 > th) bIn = false; else score += thScore - chiSquare1;
    const float a1 = f11 * u2 + f21 * v2 + f31;
    const float b1 = f12 * u2 + f22 * v2 + f32;
    const float c1 = f13 * u2 + f23 * v2 + f33;
    const float num1 = a1 * u1 + b1 * v1 + c1;
    const float squareDist2 = num1 * num1 / (a1 * a1 + b1 * b1);
    const float chiSquare2 = squareDist2 * invSigmaSquare;
    if (chiSquare2 > th) bIn = false; else score += thScore - chiSquare2;
    mask[i] = bIn ? 1 : 0;
  }
  return score;
}

// _check_H (:285-370)
float check_H(const TwoView& tv, const float* H21, const float* H12, uint8_t* mask) {
  const float h11 = H21[0], h12 = H21[1], h13 = H21[2], h21 = H21[3], h22 = H21[4], h23 = H21[5],
              h31 = H21[6], h32 = H21[7], h33 = H21[8];
  const float h11inv = H12[0], h12inv = H12[1], h13inv = H12[2], h21inv = H12[3], h22inv = H12[4],
              h23inv = H12[5], h31inv = H12[6], h32inv = H12[7], h33inv = H12[8];
  float score = 0;
  const float th = 5.991;
  const float invSigmaSquare = 1.0 / (tv.sigma * tv.sigma);
  for (int i = 0; i < tv.N; i++) {
    bool bIn = true;
    const float u1 = tv.keys1[tv.m1[i] * 2], v1 = tv.keys1[tv.m1[i] * 2 + 1];
    const float u2 = tv.keys2[tv.m2[i] * 2], v2 = tv.keys2[tv.m2[i] * 2 + 1];
    const float w2in1inv = 1.0f / (h31inv * u2 + h32inv * v2 + h33inv);
    const float u2in1 = (h11inv * u2 + h12inv * v2 + h13inv) * w2in1inv;
    const float v2in1 = (h21inv * u2 + h22inv * v2 + h23inv) * w2in1inv;
    const float squareDist1 = (u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1);
    const float chiSquare1 = squareDist1 * invSigmaSquare;
    if (chiSquare1 > th) bIn = false; else score += th - chiSquare1;
    const float w1in2inv = 1.0f / (h31 * u1 + h32 * v1 + h33);
    const float u1in2 = (h11 * u1 + h12 * v1 + h13) * w1in2inv;
    const float v1in2 = (h21 * u1 + h22 * v1 + h23) * w1in2inv;
    const float squareDist2 = (u2 - u1in2) * (u2 - u1in2) + (v2 - v1in2) * (v2 - v1in2);
    const float chiSquare2 = squareDist2 * invSigmaSquare;
    if (chiSquare2 > th) bIn = false; else score += th - chiSquare2;
    mask[i] = bIn ? 1 : 0;
  }
  return score;
}

// _find_F (:161-205) / _find_H (:114-159): best = strictly greater score, earliest wins ties.
void find_model(const TwoView& tv, int model, std::vector<uint8_t>& best_mask, float& score,
                float* best_M, int& best_idx) {
  score = 0.0f;
  best_idx = -1;
  best_mask.assign(tv.N, 0);
  for (int i = 0; i < 9; i++) best_M[i] = 0.0f;
  std::vector<uint8_t> cur(tv.N, 0);
  for (int it = 0; it < tv.n_hyp; it++) {
    float M[9], Minv[9];
    float s;
    if (model == 0) {
      fit_F(tv, tv.sets + (size_t)it * 8, M);
      s = check_F(tv, M, cur.data());
    } else {
      fit_H(tv, tv.sets + (size_t)it * 8, M, Minv);
      s = check_H(tv, M, Minv, cur.data());
    }
    if (s > score) {
      std::memcpy(best_M, M, sizeof(M));
      best_mask = cur;
      score = s;
      best_idx = it;
    }
  }
}

// _triangulate (:928-950): null vector of the 4x4 DLT matrix, de-homogenised.
void triangulate(const float* x1, const float* x2, const float* P1, const float* P2, float* X) {
  float A[16];
  for (int j = 0; j < 4; j++) {
    A[0 * 4 + j] = x1[0] * P1[2 * 4 + j] - P1[0 * 4 + j];
    A[1 * 4 + j] = x1[1] * P1[2 * 4 + j] - P1[1 * 4 + j];
    A[2 * 4 + j] = x2[0] * P2[2 * 4 + j] - P2[0 * 4 + j];
    A[3 * 4 + j] = x2[1] * P2[2 * 4 + j] - P2[1 * 4 + j];
  }
  float v[4];
  null_vector(4, 4, A, v);
  X[0] = v[0] / v[3];
  X[1] = v[1] / v[3];
  X[2] = v[2] / v[3];
}

// _check_R_T (:782-898)
int check_RT(const TwoView& tv, const float* R, const float* t, const std::vector<uint8_t>& inl,
             std::vector<float>& P3D, float th2, std::vector<uint8_t>& good, float& parallax) {
  const float fx = tv.K[0], fy = tv.K[4], cx = tv.K[2], cy = tv.K[5];
  good.assign(tv.n1, 0);
  P3D.assign((size_t)tv.n1 * 3, 0.0f);
  std::vector<float> vCos;
  vCos.reserve(tv.n1);
  float P1[12], P2[12], Rt[12];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) { P1[i * 4 + j] = tv.K[i * 3 + j]; Rt[i * 4 + j] = R[i * 3 + j]; }
    P1[i * 4 + 3] = 0.0f;
    Rt[i * 4 + 3] = t[i];
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++)
      P2[i * 4 + j] = tv.K[i * 3] * Rt[j] + tv.K[i * 3 + 1] * Rt[4 + j] + tv.K[i * 3 + 2] * Rt[8 + j];
  float O2[3];
  for (int i = 0; i < 3; i++) O2[i] = (-R[0 * 3 + i]) * t[0] + (-R[1 * 3 + i]) * t[1] + (-R[2 * 3 + i]) * t[2];
  int nGood = 0;
  for (int i = 0; i < tv.N; i++) {
    if (!inl[i]) continue;
    const float* kp1 = tv.keys1 + tv.m1[i] * 2;
    const float* kp2 = tv.keys2 + tv.m2[i] * 2;
    float p[3];
    triangulate(kp1, kp2, P1, P2, p);
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) {
      good[tv.m1[i]] = 0;
      continue;
    }
    float dist1 = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    float n2[3] = {p[0] - O2[0], p[1] - O2[1], p[2] - O2[2]};
    float dist2 = std::sqrt(n2[0] * n2[0] + n2[1] * n2[1] + n2[2] * n2[2]);
    float cosParallax = (p[0] * n2[0] + p[1] * n2[1] + p[2] * n2[2]) / (dist1 * dist2);
    if (p[2] <= 0 && cosParallax < 0.99998) continue;
    float p2[3];
    for (int r = 0; r < 3; r++) p2[r] = (R[r * 3] * p[0] + R[r * 3 + 1] * p[1] + R[r * 3 + 2] * p[2]) + t[r];
    if (p2[2] <= 0 && cosParallax < 0.99998) continue;
    float invZ1 = 1.0f / p[2];
    float im1x = fx * p[0] * invZ1 + cx;
    float im1y = fy * p[1] * invZ1 + cy;
    float squareError1 = (im1x - kp1[0]) * (im1x - kp1[0]) + (im1y - kp1[1]) * (im1y - kp1[1]);
    if (squareError1 > th2) continue;
    float invZ2 = 1.0f / p2[2];
    float im2x = fx * p2[0] * invZ2 + cx;
    float im2y = fy * p2[1] * invZ2 + cy;
    float squareError2 = (im2x - kp2[0]) * (im2x - kp2[0]) + (im2y - kp2[1]) * (im2y - kp2[1]);
    if (squareError2 > th2) continue;
    vCos.push_back(cosParallax);
    P3D[(size_t)tv.m1[i] * 3] = p[0];
    P3D[(size_t)tv.m1[i] * 3 + 1] = p[1];
    P3D[(size_t)tv.m1[i] * 3 + 2] = p[2];
    nGood++;
    if (cosParallax < 0.99998) good[tv.m1[i]] = 1;
  }
  if (nGood > 0) {
    std::sort(vCos.begin(), vCos.end());
    size_t idx = std::min(50, int(vCos.size() - 1));
    parallax = std::acos(vCos[idx]) * 180 / 3.1415926535897932384626433832795;
  } else {
    parallax = 0;
  }
  return nGood;
}

void compose(const float* R, const float* t, float* T) {
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j];
    T[i * 4 + 3] = t[i];
  }
}

// _reconstruct_F (:451-562) with _decompose_E (:900-926)
bool reconstruct_F(const TwoView& tv, const std::vector<uint8_t>& inl, const float* F21, float* T21,
                   float* P3D_out, uint8_t* tri_out, float minParallax, int minTriangulated,
                   urmvo_oracle_tv_stats* st) {
  int N = 0;
  for (int i = 0; i < tv.N; i++) N += inl[i] ? 1 : 0;
  float Kt[9], tmp[9], E[9];
  mat3_T(tv.K, Kt);
  mat3_mul(Kt, F21, tmp);
  mat3_mul(tmp, tv.K, E);
  float U[9], w[3], V[9], Vt[9];
  svd3(E, U, w, V);
  mat3_T(V, Vt);
  float t[3] = {U[2], U[5], U[8]};
  float tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  for (int i = 0; i < 3; i++) t[i] = t[i] / tn;
  const float W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  float Wt[9], R1[9], R2[9];
  mat3_T(W, Wt);
  mat3_mul(U, W, tmp);
  mat3_mul(tmp, Vt, R1);
  if (det3(R1) < 0) for (int i = 0; i < 9; i++) R1[i] = -R1[i];
  mat3_mul(U, Wt, tmp);
  mat3_mul(tmp, Vt, R2);
  if (det3(R2) < 0) for (int i = 0; i < 9; i++) R2[i] = -R2[i];
  float t1[3] = {t[0], t[1], t[2]}, t2[3] = {-t[0], -t[1], -t[2]};
  const float* Rs[4] = {R1, R2, R1, R2};
  const float* ts[4] = {t1, t1, t2, t2};
  std::vector<float> P3D[4];
  std::vector<uint8_t> tri[4];
  float parallax[4];
  int nGood[4];
  const float th2 = 4.0 * tv.sigma2;
  for (int h = 0; h < 4; h++) {
    nGood[h] = check_RT(tv, Rs[h], ts[h], inl, P3D[h], th2, tri[h], parallax[h]);
    if (st) { st->n_good[h] = nGood[h]; st->parallax[h] = parallax[h]; }
  }
  int maxGood = std::max(nGood[0], std::max(nGood[1], std::max(nGood[2], nGood[3])));
  int nMinGood = std::max(static_cast<int>(0.9 * N), minTriangulated);
  int nsimilar = 0;
  for (int h = 0; h < 4; h++)
    if (nGood[h] > 0.7 * maxGood) nsimilar++;
  if (maxGood < nMinGood || nsimilar > 1) return false;
  for (int h = 0; h < 4; h++) {
    if (maxGood == nGood[h]) {  // if / else-if chain: only the first match is tested
      if (parallax[h] > minParallax) {
        std::memcpy(P3D_out, P3D[h].data(), (size_t)tv.n1 * 3 * sizeof(float));
        std::memcpy(tri_out, tri[h].data(), (size_t)tv.n1);
        compose(Rs[h], ts[h], T21);
        if (st) st->best_motion = h;
        return true;
      }
      return false;
    }
  }
  return false;
}

// _reconstruct_H (:564-733), Faugeras' 8 hypotheses.  The reference never assigns vP3D on
// success (ORB-SLAM3 does vP3D = bestP3D); we output bestP3D.
bool reconstruct_H(const TwoView& tv, const std::vector<uint8_t>& inl, const float* H21, float* T21,
                   float* P3D_out, uint8_t* tri_out, float minParallax, int minTriangulated,
                   urmvo_oracle_tv_stats* st) {
  int N = 0;
  for (int i = 0; i < tv.N; i++) N += inl[i] ? 1 : 0;
  float invK[9], tmp[9], A[9];
  inv3f(tv.K, invK);
  mat3_mul(invK, H21, tmp);
  mat3_mul(tmp, tv.K, A);
  float U[9], w[3], V[9], Vt[9];
  svd3(A, U, w, V);
  mat3_T(V, Vt);
  float s = det3(U) * det3(Vt);
  float d1 = w[0], d2 = w[1], d3 = w[2];
  if (d1 / d2 < 1.00001 || d2 / d3 < 1.00001) return false;
  float vR[8][9], vt[8][3];
  float aux1 = std::sqrt((d1 * d1 - d2 * d2) / (d1 * d1 - d3 * d3));
  float aux3 = std::sqrt((d2 * d2 - d3 * d3) / (d1 * d1 - d3 * d3));
  float x1[] = {aux1, aux1, -aux1, -aux1};
  float x3[] = {aux3, -aux3, aux3, -aux3};
  float aux_stheta = std::sqrt((d1 * d1 - d2 * d2) * (d2 * d2 - d3 * d3)) / ((d1 + d3) * d2);
  float ctheta = (d2 * d2 + d1 * d3) / ((d1 + d3) * d2);
  float stheta[] = {aux_stheta, -aux_stheta, -aux_stheta, aux_stheta};
  float sU[9];
  for (int i = 0; i < 9; i++) sU[i] = s * U[i];
  for (int i = 0; i < 4; i++) {
    float Rp[9] = {ctheta, 0, -stheta[i], 0, 1.f, 0, stheta[i], 0, ctheta};
    mat3_mul(sU, Rp, tmp);
    mat3_mul(tmp, Vt, vR[i]);
    float tp[3] = {x1[i], 0, -x3[i]};
    for (int k = 0; k < 3; k++) tp[k] *= d1 - d3;
    float t[3];
    for (int r = 0; r < 3; r++) t[r] = U[r * 3] * tp[0] + U[r * 3 + 1] * tp[1] + U[r * 3 + 2] * tp[2];
    float tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    for (int r = 0; r < 3; r++) vt[i][r] = t[r] / tn;
  }
  float aux_sphi = std::sqrt((d1 * d1 - d2 * d2) * (d2 * d2 - d3 * d3)) / ((d1 - d3) * d2);
  float cphi = (d1 * d3 - d2 * d2) / ((d1 - d3) * d2);
  float sphi[] = {aux_sphi, -aux_sphi, -aux_sphi, aux_sphi};
  for (int i = 0; i < 4; i++) {
    float Rp[9] = {cphi, 0, sphi[i], 0, -1, 0, sphi[i], 0, -cphi};
    mat3_mul(sU, Rp, tmp);
    mat3_mul(tmp, Vt, vR[4 + i]);
    float tp[3] = {x1[i], 0, x3[i]};
    for (int k = 0; k < 3; k++) tp[k] *= d1 + d3;
    float t[3];
    for (int r = 0; r < 3; r++) t[r] = U[r * 3] * tp[0] + U[r * 3 + 1] * tp[1] + U[r * 3 + 2] * tp[2];
    float tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    for (int r = 0; r < 3; r++) vt[4 + i][r] = t[r] / tn;
  }
  int bestGood = 0, secondBestGood = 0, bestSolutionIdx = -1;
  float bestParallax = -1;
  std::vector<float> bestP3D;
  std::vector<uint8_t> bestTri;
  const float th2 = 4.0 * tv.sigma2;
  for (int i = 0; i < 8; i++) {
    float parallaxi;
    std::vector<float> P3Di;
    std::vector<uint8_t> trii;
    int nGood = check_RT(tv, vR[i], vt[i], inl, P3Di, th2, trii, parallaxi);
    if (st) { st->n_good[i] = nGood; st->parallax[i] = parallaxi; }
    if (nGood > bestGood) {
      secondBestGood = bestGood;
      bestGood = nGood;
      bestSolutionIdx = i;
      bestParallax = parallaxi;
      bestP3D = P3Di;
      bestTri = trii;
    } else if (nGood > secondBestGood) {
      secondBestGood = nGood;
    }
  }
  if (secondBestGood < 0.75 * bestGood && bestParallax >= minParallax && bestGood > minTriangulated &&
      bestGood > 0.9 * N) {
    compose(vR[bestSolutionIdx], vt[bestSolutionIdx], T21);
    std::memcpy(P3D_out, bestP3D.data(), (size_t)tv.n1 * 3 * sizeof(float));
    std::memcpy(tri_out, bestTri.data(), (size_t)tv.n1);
    if (st) st->best_motion = bestSolutionIdx;
    return true;
  }
  return false;
}

void tv_setup(TwoView& tv, int n1, const float* keys1, int n2, const float* keys2,
              const int32_t* matches12, const float* K, float sigma, int n_hyp,
              const int32_t* sets) {
  tv.n1 = n1; tv.n2 = n2; tv.keys1 = keys1; tv.keys2 = keys2;
  for (int i = 0; i < n1; i++)  // :34-40
    if (matches12[i] >= 0) { tv.m1.push_back(i); tv.m2.push_back(matches12[i]); }
  tv.N = (int)tv.m1.size();
  if (K) std::memcpy(tv.K, K, sizeof(tv.K));
  tv.sigma = sigma;
  tv.sigma2 = sigma * sigma;
  tv.n_hyp = n_hyp;
  tv.sets = sets;
  normalize(n1, keys1, tv.pn1, tv.T1);
  normalize(n2, keys2, tv.pn2, tv.T2);
}

}  // namespace

extern "C" int urmvo_oracle_two_view_mode(int n1, const float* keys1, int n2, const float* keys2,
                                          const int32_t* matches12, const float* K, float sigma,
                                          int n_hyp, const int32_t* sets, int score_mode, float* T21, float* P3D,
                                          uint8_t* triangulated, uint8_t* mask_H, uint8_t* mask_F,
                                          urmvo_oracle_tv_stats* stats);
extern "C" int urmvo_oracle_two_view(int n1, const float* keys1, int n2, const float* keys2,
                                     const int32_t* matches12, const float* K, float sigma,
                                     int n_hyp, const int32_t* sets, float* T21, float* P3D,
                                     uint8_t* triangulated, uint8_t* mask_H, uint8_t* mask_F,
                                     urmvo_oracle_tv_stats* stats) {
  return urmvo_oracle_two_view_mode(n1, keys1, n2, keys2, matches12, K, sigma, n_hyp, sets, 0, T21, P3D,
                                    triangulated, mask_H, mask_F, stats);
}

extern "C" int urmvo_oracle_two_view_mode(int n1, const float* keys1, int n2, const float* keys2,
                                          const int32_t* matches12, const float* K, float sigma,
                                          int n_hyp, const int32_t* sets, int score_mode, float* T21, float* P3D,
                                          uint8_t* triangulated, uint8_t* mask_H, uint8_t* mask_F,
                                          urmvo_oracle_tv_stats* stats) {
  TwoView tv;
  tv.score_mode = score_mode;
  tv_setup(tv, n1, keys1, n2, keys2, matches12, K, sigma, n_hyp, sets);
  urmvo_oracle_tv_stats local;
  urmvo_oracle_tv_stats* st = stats ? stats : &local;
  std::memset(st, 0, sizeof(*st));
  st->best_motion = -1;
  std::vector<uint8_t> inlH, inlF;
  float SH = 0, SF = 0, H[9], F[9];
  int bH = -1, bF = -1;
  // :77-84 two threads, H || F
  std::thread thH([&] { find_model(tv, 1, inlH, SH, H, bH); });
  std::thread thF([&] { find_model(tv, 0, inlF, SF, F, bF); });
  thH.join();
  thF.join();
  st->SH = SH; st->SF = SF; st->best_H = bH; st->best_F = bF;
  std::memcpy(st->H21, H, sizeof(H));
  std::memcpy(st->F21, F, sizeof(F));
  if (mask_H) std::memcpy(mask_H, inlH.data(), (size_t)tv.N);
  if (mask_F) std::memcpy(mask_F, inlF.data(), (size_t)tv.N);
  std::memset(P3D, 0, (size_t)n1 * 3 * sizeof(float));
  std::memset(triangulated, 0, (size_t)n1);
  if (SH + SF == 0.f) { st->used_H = -1; return 0; }  // :87-88
  float RH = SH / (SH + SF);
  float minParallax = 1.0;
  if (RH > 0.50) {  // :92-97
    st->used_H = 1;
    return reconstruct_H(tv, inlH, H, T21, P3D, triangulated, minParallax, 50, st) ? 1 : 0;
  }
  st->used_H = 0;
  return reconstruct_F(tv, inlF, F, T21, P3D, triangulated, minParallax, 50, st) ? 1 : 0;
}

extern "C" int urmvo_oracle_score_all_mode(int n1, const float* keys1, int n2, const float* keys2,
                                           const int32_t* matches12, float sigma, int n_hyp,
                                           const int32_t* sets, int model, int score_mode, float* scores,
                                           uint32_t* masks, float* models);
extern "C" int urmvo_oracle_score_all(int n1, const float* keys1, int n2, const float* keys2,
                                      const int32_t* matches12, float sigma, int n_hyp,
                                      const int32_t* sets, int model, float* scores,
                                      uint32_t* masks, float* models) {
  return urmvo_oracle_score_all_mode(n1, keys1, n2, keys2, matches12, sigma, n_hyp, sets, model, 0, scores, masks, models);
}

extern "C" int urmvo_oracle_score_all_mode(int n1, const float* keys1, int n2, const float* keys2,
                                           const int32_t* matches12, float sigma, int n_hyp,
                                           const int32_t* sets, int model, int score_mode, float* scores,
                                           uint32_t* masks, float* models) {
  TwoView tv;
  tv.score_mode = score_mode;
  tv_setup(tv, n1, keys1, n2, keys2, matches12, nullptr, sigma, n_hyp, sets);
  const int words = (tv.N + 31) / 32;
  std::vector<uint8_t> cur(tv.N);
  for (int it = 0; it < n_hyp; it++) {
    float M[9], Minv[9];
    float s;
    if (model == 0) {
      fit_F(tv, sets + (size_t)it * 8, M);
      s = check_F(tv, M, cur.data());
    } else {
      fit_H(tv, sets + (size_t)it * 8, M, Minv);
      s = check_H(tv, M, Minv, cur.data());
    }
    scores[it] = s;
    if (models) std::memcpy(models + (size_t)it * 9, M, sizeof(M));
    if (masks) {
      uint32_t* mw = masks + (size_t)it * words;
      for (int wd = 0; wd < words; wd++) mw[wd] = 0;
      for (int i = 0; i < tv.N; i++)
        if (cur[i]) mw[i >> 5] |= (1u << (i & 31));
    }
  }
  return tv.N;
}

// reconstruct() :45-71 with Random::RandomInt (:109-112) on glibc rand().
extern "C" void urmvo_oracle_draw_sets(int N, int n_hyp, int reseed, int seed, int32_t* sets) {
  if (reseed) srand(seed);
  std::vector<size_t> all(N), avail;
  for (int i = 0; i < N; i++) all[i] = i;
  for (int it = 0; it < n_hyp; it++) {
    avail = all;
    for (int j = 0; j < 8; j++) {
      int d = (int)avail.size() - 1 - 0 + 1;
      int randi = int(((double)rand() / ((double)RAND_MAX + 1.0)) * d) + 0;
      sets[(size_t)it * 8 + j] = (int32_t)avail[randi];
      avail[randi] = avail.back();
      avail.pop_back();
    }
  }
}

extern "C" void urmvo_oracle_svd(int m, int n, const float* Ain, float* sigma, float* U, float* V) {
  std::vector<float> A(Ain, Ain + (size_t)m * n), Vt((size_t)n * n);
  jacobi_onesided(m, n, A.data(), Vt.data());
  std::vector<float> nrm(n);
  std::vector<int> idx(n);
  for (int j = 0; j < n; j++) { nrm[j] = col_norm(m, n, A.data(), j); idx[j] = j; }
  for (int i = 0; i < n - 1; i++) {
    int b = i;
    for (int j = i + 1; j < n; j++)
      if (nrm[idx[j]] > nrm[idx[b]]) b = j;
    std::swap(idx[i], idx[b]);
  }
  for (int j = 0; j < n; j++) {
    int s = idx[j];
    sigma[j] = nrm[s];
    for (int k = 0; k < n; k++) V[k * n + j] = Vt[k * n + s];
    if (U)
      for (int k = 0; k < m; k++) U[k * n + j] = (nrm[s] > 0.0f) ? A[k * n + s] / nrm[s] : 0.0f;
  }
}
