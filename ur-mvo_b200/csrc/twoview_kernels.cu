// csrc/twoview_kernels.cu — two-view RANSAC (fundamental + homography) and motion recovery on sm_100a.
//
// Reference behaviour reproduced, /root/reference/src/epipolar_geometry.cc:
//   _normalize :735-780, _compute_F21 :247-283, _compute_H21 :207-245, _check_F :372-449,
//   _check_H :285-370, _find_F/_find_H arg-max :153-157/:199-203, _decompose_E :900-926,
//   _reconstruct_F :451-562, _reconstruct_H :564-733, _check_R_T :782-898, _triangulate :928-950.
//
// Everything is fp32 like the reference.  THIS FILE MUST BE COMPILED WITH -fmad=false (and the
// default IEEE -prec-div=true -prec-sqrt=true): the parity contract is bit-exact scores and inlier
// masks for identical 8-point sets, which needs the same operation sequence with no contraction.
// Eigen::JacobiSVD is replaced by a fully specified one-sided Jacobi SVD (cyclic pair order,
// threshold 5e-7, <= 30 sweeps) — see DESIGN.md §6.
//
//   tv_normalize_kernel   one CTA per image; the running sums are sequential like the reference's
//   tv_fit_sub_kernel     8 (F) / 16 (H) lanes per hypothesis: 8-point DLT + Jacobi SVD, lane = row
//   tv_score_kernel       one warp per (model, hypothesis): lanes stride over the matches,
//                         __ballot_sync builds the inlier mask words, the score is accumulated in
//                         match order (the reference's summation order) by a lane-ordered chain
//   tv_argmax_kernel      deterministic arg-max: highest score, earliest hypothesis on ties
//   tv_motion_kernel      model selection, E / H decomposition, per-match DLT triangulation and
//                         cheirality vote for the 4 (F) or 8 (H) motion hypotheses

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "kernels.h"

namespace urmvo {

namespace {

constexpr int kMaxSweeps = 30;
constexpr float kJacobiTol = 5e-7f;
constexpr float kJacobiTol2 = kJacobiTol * kJacobiTol;  // the test is gamma^2 <= tol^2 alpha beta (no square root)
// columns with squared norm <= kJacobiTiny * ||A||_F^2 are numerically zero and are not rotated
// (otherwise the null column of a rank-deficient system keeps every solve at the sweep limit)
constexpr float kJacobiTiny = 1e-14f;

// One-sided (Hestenes) Jacobi on an m x n row-major matrix; V (n x n) accumulates the rotations.
// Deliberately a real (non-inlined) function with run-time sizes: one copy of the loop nest serves
// the 8x9, 16x9, 4x4 and 3x3 uses, and the operation order is literally the restatement's.
__device__ __noinline__ void jacobi_onesided(int m, int n, float* A, float* V) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0f : 0.0f;
  float fro2 = 0.0f;  // sum over rows of the row sums
  for (int k = 0; k < m; k++) {
    float row = 0.0f;
    for (int j = 0; j < n; j++) row += A[k * n + j] * A[k * n + j];
    fro2 = (k == 0) ? row : fro2 + row;
  }
  const float tiny = kJacobiTiny * fro2;
  for (int sweep = 0; sweep < kMaxSweeps; sweep++) {
    bool rotated = false;
    for (int p = 0; p < n - 1; p++) {
      for (int q = p + 1; q < n; q++) {
        float alpha = 0.0f, beta = 0.0f, gamma = 0.0f;
        for (int k = 0; k < m; k++) {
          const float ap = A[k * n + p], aq = A[k * n + q];
          alpha += ap * ap;
          beta += aq * aq;
          gamma += ap * aq;
        }
        if (alpha <= tiny || beta <= tiny) continue;
        if (gamma * gamma <= kJacobiTol2 * (alpha * beta)) continue;
        rotated = true;
        const float dlt = beta - alpha;
        const float rad = sqrtf(dlt * dlt + 4.0f * (gamma * gamma));
        const float t = (2.0f * gamma) / (dlt >= 0.0f ? dlt + rad : dlt - rad);
        const float c = 1.0f / sqrtf(1.0f + t * t);
        const float s = c * t;
        for (int k = 0; k < m; k++) {
          const float ap = A[k * n + p], aq = A[k * n + q];
          A[k * n + p] = c * ap - s * aq;
          A[k * n + q] = s * ap + c * aq;
        }
        for (int k = 0; k < n; k++) {
          const float vp = V[k * n + p], vq = V[k * n + q];
          V[k * n + p] = c * vp - s * vq;
          V[k * n + q] = s * vp + c * vq;
        }
      }
    }
    if (!rotated) break;
  }
}

__device__ __forceinline__ float col_norm(int m, int n, const float* A, int j) {
  float s = 0.0f;
  for (int k = 0; k < m; k++) s += A[k * n + j] * A[k * n + j];
  return sqrtf(s);
}

// Right singular vector of the smallest singular value (first index on ties). n <= 9.
__device__ __noinline__ void null_vector(int m, int n, float* A, float* v) {
  float V[81];
  jacobi_onesided(m, n, A, V);
  int best = 0;
  float bn = col_norm(m, n, A, 0);
  for (int j = 1; j < n; j++) {
    const float nj = col_norm(m, n, A, j);
    if (nj < bn) { bn = nj; best = j; }
  }
  for (int k = 0; k < n; k++) v[k] = V[k * n + best];
}

// One Jacobi pair (P,Q) of the 3x3 problem with compile-time column indices: the matrices stay in
// registers.  Same operations in the same order as jacobi_onesided(3, 3, ...).
template <int P, int Q>
__device__ __forceinline__ bool jacobi3_pair(float (&A)[9], float (&V)[9], float tiny) {
  float alpha = 0.0f, beta = 0.0f, gamma = 0.0f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float ap = A[k * 3 + P], aq = A[k * 3 + Q];
    alpha += ap * ap;
    beta += aq * aq;
    gamma += ap * aq;
  }
  if (alpha <= tiny || beta <= tiny) return false;
  if (gamma * gamma <= kJacobiTol2 * (alpha * beta)) return false;
  const float dlt = beta - alpha;
  const float rad = sqrtf(dlt * dlt + 4.0f * (gamma * gamma));
  const float t = (2.0f * gamma) / (dlt >= 0.0f ? dlt + rad : dlt - rad);
  const float c = 1.0f / sqrtf(1.0f + t * t);
  const float s = c * t;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float ap = A[k * 3 + P], aq = A[k * 3 + Q];
    A[k * 3 + P] = c * ap - s * aq;
    A[k * 3 + Q] = s * ap + c * aq;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float vp = V[k * 3 + P], vq = V[k * 3 + Q];
    V[k * 3 + P] = c * vp - s * vq;
    V[k * 3 + Q] = s * vp + c * vq;
  }
  return true;
}

// 3x3 SVD, singular values descending; U.col(2) = +-(u0 x u1) such that U diag(w) V^T = A.
__device__ __noinline__ void svd3(const float* Ain, float* U, float* w, float* V) {
  float A[9], Vt[9];
#pragma unroll
  for (int i = 0; i < 9; i++) { A[i] = Ain[i]; Vt[i] = (i % 4 == 0) ? 1.0f : 0.0f; }
  float fro2 = 0.0f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float row = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; j++) row += A[k * 3 + j] * A[k * 3 + j];
    fro2 = (k == 0) ? row : fro2 + row;
  }
  const float tiny = kJacobiTiny * fro2;
  for (int sweep = 0; sweep < kMaxSweeps; sweep++) {
    bool rotated = jacobi3_pair<0, 1>(A, Vt, tiny);
    rotated = jacobi3_pair<0, 2>(A, Vt, tiny) || rotated;
    rotated = jacobi3_pair<1, 2>(A, Vt, tiny) || rotated;
    if (!rotated) break;
  }
  float nrm[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float sq = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) sq += A[k * 3 + j] * A[k * 3 + j];
    nrm[j] = sqrtf(sq);
  }
  // selection sort, descending, first index wins ties (indices kept in registers)
  int i0 = 0, i1 = 1, i2 = 2;
  {
    int bsel = 0;
    if (nrm[1] > nrm[0]) bsel = 1;
    if (nrm[2] > nrm[bsel]) bsel = 2;
    if (bsel == 1) { i0 = 1; i1 = 0; }
    else if (bsel == 2) { i0 = 2; i2 = 0; }
    auto nv = [&](int idx) { return idx == 0 ? nrm[0] : (idx == 1 ? nrm[1] : nrm[2]); };
    if (nv(i2) > nv(i1)) { const int tmp = i1; i1 = i2; i2 = tmp; }
  }
  auto colA = [&](int k, int idx) { return idx == 0 ? A[k * 3] : (idx == 1 ? A[k * 3 + 1] : A[k * 3 + 2]); };
  auto colV = [&](int k, int idx) { return idx == 0 ? Vt[k * 3] : (idx == 1 ? Vt[k * 3 + 1] : Vt[k * 3 + 2]); };
  auto nsel = [&](int idx) { return idx == 0 ? nrm[0] : (idx == 1 ? nrm[1] : nrm[2]); };
  const int idx[3] = {i0, i1, i2};
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const int sidx = idx[j];
    const float nj = nsel(sidx);
    w[j] = nj;
#pragma unroll
    for (int k = 0; k < 3; k++) V[k * 3 + j] = colV(k, sidx);
    if (j < 2) {
#pragma unroll
      for (int k = 0; k < 3; k++) U[k * 3 + j] = (nj > 0.0f) ? colA(k, sidx) / nj : 0.0f;
    }
  }
  float c0 = U[3 + 0] * U[6 + 1] - U[6 + 0] * U[3 + 1];
  float c1 = U[6 + 0] * U[0 + 1] - U[0 + 0] * U[6 + 1];
  float c2 = U[0 + 0] * U[3 + 1] - U[3 + 0] * U[0 + 1];
  const float d = c0 * colA(0, i2) + c1 * colA(1, i2) + c2 * colA(2, i2);
  if (d < 0.0f) { c0 = -c0; c1 = -c1; c2 = -c2; }
  U[2] = c0; U[5] = c1; U[8] = c2;
}

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3_T(const float* A, float* B) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[i * 3 + j] = A[j * 3 + i];
}
__device__ __forceinline__ float det3(const float* a) {
  return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
         a[2] * (a[3] * a[7] - a[4] * a[6]);
}
__device__ __forceinline__ void inv3f(const float* a, float* r) {
  const float c00 = a[4] * a[8] - a[5] * a[7];
  const float c01 = a[5] * a[6] - a[3] * a[8];
  const float c02 = a[3] * a[7] - a[4] * a[6];
  const float det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  const float id = 1.0f / det;
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

}  // namespace

// ------------------------------------------------------------------------------- normalisation

// grid = 2 (image 1, image 2). keys: n*2, pn: n*2 out, T: 9 out.
__global__ void tv_normalize_kernel(int n1, const float* __restrict__ keys1, float* __restrict__ pn1,
                                    float* __restrict__ T1, int n2, const float* __restrict__ keys2,
                                    float* __restrict__ pn2, float* __restrict__ T2) {
  const int n = blockIdx.x == 0 ? n1 : n2;
  const float* keys = blockIdx.x == 0 ? keys1 : keys2;
  float* pn = blockIdx.x == 0 ? pn1 : pn2;
  float* T = blockIdx.x == 0 ? T1 : T2;
  constexpr int CH = 2048;  // keypoints staged per chunk (16 KB)
  __shared__ float buf[CH * 2];
  __shared__ float sh[4];
  // the reference's running sums are sequential (src/epipolar_geometry.cc:744-761): one thread adds
  // in keypoint order, the others only stage the data into shared memory
  float meanX = 0, meanY = 0;
  for (int c0 = 0; c0 < n; c0 += CH) {
    const int m = min(CH, n - c0);
    for (int i = threadIdx.x; i < m * 2; i += blockDim.x) buf[i] = keys[(size_t)c0 * 2 + i];
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < m; i++) { meanX += buf[i * 2]; meanY += buf[i * 2 + 1]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { sh[0] = meanX / n; sh[1] = meanY / n; }
  __syncthreads();
  meanX = sh[0]; meanY = sh[1];
  float meanDevX = 0, meanDevY = 0;
  for (int c0 = 0; c0 < n; c0 += CH) {
    const int m = min(CH, n - c0);
    for (int i = threadIdx.x; i < m * 2; i += blockDim.x) buf[i] = keys[(size_t)c0 * 2 + i];
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < m; i++) {
        meanDevX += fabsf(buf[i * 2] - meanX);
        meanDevY += fabsf(buf[i * 2 + 1] - meanY);
      }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    meanDevX = meanDevX / n;
    meanDevY = meanDevY / n;
    const float sX = 1.0f / meanDevX, sY = 1.0f / meanDevY;
    sh[2] = sX; sh[3] = sY;
    for (int i = 0; i < 9; i++) T[i] = 0.0f;
    T[0] = sX; T[4] = sY; T[2] = -meanX * sX; T[5] = -meanY * sY; T[8] = 1.0f;
  }
  __syncthreads();
  const float sX = sh[2], sY = sh[3];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    pn[i * 2] = (keys[i * 2] - meanX) * sX;
    pn[i * 2 + 1] = (keys[i * 2 + 1] - meanY) * sY;
  }
}

// Packs the valid matches: uv[i] = (u1, v1, u2, v2) in pixels, pnm likewise in normalised coords.
__global__ void tv_gather_kernel(int N, const int* __restrict__ m1, const int* __restrict__ m2,
                                 const float* __restrict__ keys1, const float* __restrict__ keys2,
                                 const float* __restrict__ pn1, const float* __restrict__ pn2,
                                 float4* __restrict__ uv, float4* __restrict__ pnm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int a = m1[i], b = m2[i];
  uv[i] = make_float4(keys1[a * 2], keys1[a * 2 + 1], keys2[b * 2], keys2[b * 2 + 1]);
  pnm[i] = make_float4(pn1[a * 2], pn1[a * 2 + 1], pn2[b * 2], pn2[b * 2 + 1]);
}

// ------------------------------------------------------------------------------- model fitting

// ------------------------------------------------------------------------------- cooperative fit
//
// The first version kept one thread per hypothesis (matrices in local memory, ~3.5 warps per SM for
// 16384 hypotheses: latency-bound, profiles/README.md).  tv_fit_sub_kernel gives every hypothesis a SUB-WARP of
// L = M lanes (8 for the 8x9 fundamental system, 16 for the 16x9 homography system): lane k owns
// row k of A (and row k of V), the matrices live in shared memory, the three column dot products
// of a Jacobi pair are formed as per-lane products followed by an xor butterfly over the rows
// (log2 M shuffle steps): the specification sums the rows of these systems as a pairwise tree, which
// is exactly what the butterfly computes, so every value is bit-identical to the CPU restatement.
// (A lane-ordered chain of M adds, the first version, cost 2-3x the shuffles for the same result
// class.)  4 (F) or 2 (H) hypotheses per warp.

// Sum over the M rows (= lanes of the sub-warp): xor butterfly, i.e. the pairwise tree
// ((v0+v1)+(v2+v3))+... of the specification for m >= 8; every lane ends with the same bits.
template <int L>
__device__ __forceinline__ float row_sum(float v, unsigned mask) {
#pragma unroll
  for (int w = 1; w < L; w <<= 1) v = v + __shfl_xor_sync(mask, v, w, L);
  return v;
}

// A: M x 9 in shared memory (row-major) on entry, sub-warp of L = M lanes, r = lane in sub-warp.  Lane r keeps
// row r of A and (r < 9) row r of V in REGISTERS for the whole decomposition: the 36 pairs of a sweep are fully
// unrolled, so every column index is static and a pair costs its three butterflies and the rotation — no
// shared-memory round trip between consecutive pairs (the first version re-read and re-wrote A and V per pair:
// 170 us for 8192 homography fits, this one 0.1 ms; same operations in the same order, bit-identical results).
template <int M>
__device__ void jacobi_null_vector_sub(float* A, float* V, int r, unsigned mask, float* v_out) {
  constexpr int N = 9;
  static_assert(M >= N, "lane r < 9 holds row r of V");
  float a[N], v[N];
#pragma unroll
  for (int j = 0; j < N; j++) { a[j] = A[r * N + j]; v[j] = (j == r) ? 1.0f : 0.0f; }
  float tiny;
  {
    float row = 0.0f;
#pragma unroll
    for (int j = 0; j < N; j++) row += a[j] * a[j];
    tiny = kJacobiTiny * row_sum<M>(row, mask);  // sum over rows of the row sums
  }
  for (int sweep = 0; sweep < kMaxSweeps; sweep++) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < N - 1; p++) {
#pragma unroll
      for (int q = p + 1; q < N; q++) {
        const float ap = a[p], aq = a[q];
        const float alpha = row_sum<M>(ap * ap, mask);
        const float beta = row_sum<M>(aq * aq, mask);
        const float gamma = row_sum<M>(ap * aq, mask);
        if (alpha <= tiny || beta <= tiny) continue;
        if (gamma * gamma <= kJacobiTol2 * (alpha * beta)) continue;
        rotated = true;
        const float dlt = beta - alpha;
        const float rad = sqrtf(dlt * dlt + 4.0f * (gamma * gamma));
        const float t = (2.0f * gamma) / (dlt >= 0.0f ? dlt + rad : dlt - rad);
        const float c = 1.0f / sqrtf(1.0f + t * t);
        const float s = c * t;
        a[p] = c * ap - s * aq;
        a[q] = s * ap + c * aq;
        const float vp = v[p], vq = v[q];  // rows >= 9 of a 16-lane group carry zeros: harmless
        v[p] = c * vp - s * vq;
        v[q] = s * vp + c * vq;
      }
    }
    if (!rotated) break;
  }
  // column with the smallest norm (first index on ties)
  int best = 0;
  float bn = 0.0f;
#pragma unroll
  for (int j = 0; j < N; j++) {
    const float nj = sqrtf(row_sum<M>(a[j] * a[j], mask));
    if (j == 0 || nj < bn) { bn = nj; best = j; }
  }
  float vb = v[0];
#pragma unroll
  for (int j = 1; j < N; j++) vb = (best == j) ? v[j] : vb;
  __syncwarp(mask);
  if (r < N) V[r] = vb;  // V.col(best): lane k holds V[k][best]
  __syncwarp(mask);
#pragma unroll
  for (int k = 0; k < N; k++) v_out[k] = V[k];
}

// Null vector of the 8x9 fundamental system: Householder QR of A^T in registers, lane r of a 16-lane
// group = row r of the 9x8 matrix (specification: null_vector_qr_8x9 of the CPU restatement).
__device__ void qr_null_vector_8x9(float* A, int r, unsigned mask, float* v_out) {
  float mrow[8], vk[8], bk[8];
#pragma unroll
  for (int k = 0; k < 8; k++) mrow[k] = r < 9 ? A[k * 9 + r] : 0.0f;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const float x = r >= k ? mrow[k] : 0.0f;
    const float sigma = row_sum<16>(x * x, mask);
    const float xkk = __shfl_sync(mask, mrow[k], k, 16);
    const float norm = sqrtf(sigma);
    const float alpha = xkk >= 0.0f ? -norm : norm;
    const float v = r == k ? x - alpha : x;
    const float beta = norm > 0.0f ? 1.0f / (norm * (norm + fabsf(xkk))) : 0.0f;
    vk[k] = v;
    bk[k] = beta;
#pragma unroll
    for (int j = k + 1; j < 8; j++) {
      const float w = beta * row_sum<16>(v * mrow[j], mask);
      mrow[j] = mrow[j] - w * v;
    }
  }
  float y = r == 8 ? 1.0f : 0.0f;
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    const float w = bk[k] * row_sum<16>(vk[k] * y, mask);
    y = y - w * vk[k];
  }
  __syncwarp(mask);
  if (r < 9) A[r] = y;  // the staging area is free now
  __syncwarp(mask);
  for (int k = 0; k < 9; k++) v_out[k] = A[k];
}

// MODEL 0: fundamental (8x9 system, Householder null vector), MODEL 1: homography (16x9 system,
// Jacobi SVD); 16 lanes per hypothesis in both.  models: [2][n_hyp][18].
template <int MODEL>
__global__ void __launch_bounds__(256)
tv_fit_sub_kernel(int n_hyp, const int* __restrict__ sets, const float4* __restrict__ pnm,
                  const float* __restrict__ T1, const float* __restrict__ T2, float* __restrict__ models) {
  constexpr int M = 16;
  constexpr int PER_WARP = 32 / M;
  constexpr int FL = MODEL == 0 ? 8 * 9 : M * 9 + 81;  // floats per hypothesis
  extern __shared__ float sm_fit[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane / M, r = lane - sub * M;
  const unsigned mask = 0xFFFFu << (sub * M);
  const int hyp = (blockIdx.x * (blockDim.x >> 5) + wid) * PER_WARP + sub;
  if (hyp >= n_hyp) return;  // whole sub-warps leave together
  float* A = sm_fit + (size_t)(wid * PER_WARP + sub) * FL;
  float* V = A + M * 9;
  const int* set = sets + (size_t)hyp * 8;
  if (MODEL == 1 || r < 8) {
    const float4 m = pnm[set[MODEL == 0 ? r : (r >> 1)]];
    const float u1 = m.x, v1 = m.y, u2 = m.z, v2 = m.w;
    float* row = A + r * 9;
    if (MODEL == 0) {
      row[0] = u2 * u1; row[1] = u2 * v1; row[2] = u2;
      row[3] = v2 * u1; row[4] = v2 * v1; row[5] = v2;
      row[6] = u1; row[7] = v1; row[8] = 1.0f;
    } else if ((r & 1) == 0) {
      row[0] = 0.0f; row[1] = 0.0f; row[2] = 0.0f;
      row[3] = -u1; row[4] = -v1; row[5] = -1.0f;
      row[6] = v2 * u1; row[7] = v2 * v1; row[8] = v2;
    } else {
      row[0] = u1; row[1] = v1; row[2] = 1.0f;
      row[3] = 0.0f; row[4] = 0.0f; row[5] = 0.0f;
      row[6] = -u2 * u1; row[7] = -u2 * v1; row[8] = -u2;
    }
  }
  __syncwarp(mask);
  float nv[9];
  if (MODEL == 0) qr_null_vector_8x9(A, r, mask, nv);
  else jacobi_null_vector_sub<M>(A, V, r, mask, nv);
  if (r != 0) return;
  float t1[9], t2[9];
  for (int i = 0; i < 9; i++) { t1[i] = T1[i]; t2[i] = T2[i]; }
  float* out = models + ((size_t)MODEL * n_hyp + hyp) * 18;
  if (MODEL == 0) {
    float U[9], w[3], Vs[9];
    svd3(nv, U, w, Vs);
    w[2] = 0.0f;
    float Fn[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        Fn[i * 3 + j] = (U[i * 3] * w[0]) * Vs[j * 3] + (U[i * 3 + 1] * w[1]) * Vs[j * 3 + 1] +
                        (U[i * 3 + 2] * w[2]) * Vs[j * 3 + 2];
    float T2t[9], tmp[9], F21[9];
    mat3_T(t2, T2t);
    mat3_mul(T2t, Fn, tmp);
    mat3_mul(tmp, t1, F21);
    for (int i = 0; i < 9; i++) { out[i] = F21[i]; out[9 + i] = 0.0f; }
  } else {
    float T2inv[9], tmp[9], H21[9], H12[9];
    inv3f(t2, T2inv);
    mat3_mul(T2inv, nv, tmp);
    mat3_mul(tmp, t1, H21);
    inv3f(H21, H12);
    for (int i = 0; i < 9; i++) { out[i] = H21[i]; out[9 + i] = H12[i]; }
  }
}

// ------------------------------------------------------------------------------- scoring

// One warp per (model, hypothesis).  uv staged in shared memory when it fits.
// scores: [2][n_hyp], masks: [2][n_hyp][words].
__global__ void __launch_bounds__(256)
tv_score_kernel(int N, int n_hyp, const float4* __restrict__ uv_g, const float* __restrict__ models,
                float inv_sigma2, int stage_uv, int score_mode, float* __restrict__ scores,
                uint32_t* __restrict__ masks) {
  extern __shared__ float4 uv_s[];
  if (stage_uv) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) uv_s[i] = uv_g[i];
    __syncthreads();
  }
  const float4* uvp = stage_uv ? uv_s : uv_g;
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int words = (N + 31) >> 5;
  const float thF = (float)3.841, thScore = (float)5.991, thH = (float)5.991;
  for (int g = blockIdx.x * wpc + (threadIdx.x >> 5); g < 2 * n_hyp; g += gridDim.x * wpc) {
    const int model = g / n_hyp;
    const float* Mp = models + (size_t)g * 18;
    float m[18];
#pragma unroll
    for (int i = 0; i < 18; i++) m[i] = Mp[i];
    float score = 0.0f;
    uint32_t* mw = masks + (size_t)g * words;
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      float c1 = 0.0f, c2 = 0.0f;
      bool bIn = false;
      if (i < N) {
        const float4 p = uvp[i];
        const float u1 = p.x, v1 = p.y, u2 = p.z, v2 = p.w;
        bIn = true;
        if (model == 0 && score_mode == 1) {
          // Sampson error (BASELINE.json north_star (4); extra mode, not the reference's metric):
          // one 1-dof test per match, same operation order as oracle check_F_sampson
          const float a2 = m[0] * u1 + m[1] * v1 + m[2];
          const float b2 = m[3] * u1 + m[4] * v1 + m[5];
          const float cc2 = m[6] * u1 + m[7] * v1 + m[8];
          const float num2 = a2 * u2 + b2 * v2 + cc2;
          const float a1 = m[0] * u2 + m[3] * v2 + m[6];
          const float b1 = m[1] * u2 + m[4] * v2 + m[7];
          const float den = (a2 * a2 + b2 * b2) + (a1 * a1 + b1 * b1);
          const float sampson = num2 * num2 / den;
          const float chiSquare = sampson * inv_sigma2;
          if (chiSquare > thF) bIn = false; else c1 = thScore - chiSquare;
        } else if (model == 0) {
          const float a2 = m[0] * u1 + m[1] * v1 + m[2];
          const float b2 = m[3] * u1 + m[4] * v1 + m[5];
          const float cc2 = m[6] * u1 + m[7] * v1 + m[8];
          const float num2 = a2 * u2 + b2 * v2 + cc2;
          const float squareDist1 = num2 * num2 / (a2 * a2 + b2 * b2);
          const float chiSquare1 = squareDist1 * inv_sigma2;
          if (chiSquare1 > thF) bIn = false; else c1 = thScore - chiSquare1;
          const float a1 = m[0] * u2 + m[3] * v2 + m[6];
          const float b1 = m[1] * u2 + m[4] * v2 + m[7];
          const float cc1 = m[2] * u2 + m[5] * v2 + m[8];
          const float num1 = a1 * u1 + b1 * v1 + cc1;
          const float squareDist2 = num1 * num1 / (a1 * a1 + b1 * b1);
          const float chiSquare2 = squareDist2 * inv_sigma2;
          if (chiSquare2 > thF) bIn = false; else c2 = thScore - chiSquare2;
        } else {
          const float w2in1inv = 1.0f / (m[15] * u2 + m[16] * v2 + m[17]);
          const float u2in1 = (m[9] * u2 + m[10] * v2 + m[11]) * w2in1inv;
          const float v2in1 = (m[12] * u2 + m[13] * v2 + m[14]) * w2in1inv;
          const float squareDist1 = (u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1);
          const float chiSquare1 = squareDist1 * inv_sigma2;
          if (chiSquare1 > thH) bIn = false; else c1 = thH - chiSquare1;
          const float w1in2inv = 1.0f / (m[6] * u1 + m[7] * v1 + m[8]);
          const float u1in2 = (m[0] * u1 + m[1] * v1 + m[2]) * w1in2inv;
          const float v1in2 = (m[3] * u1 + m[4] * v1 + m[5]) * w1in2inv;
          const float squareDist2 = (u2 - u1in2) * (u2 - u1in2) + (v2 - v1in2) * (v2 - v1in2);
          const float chiSquare2 = squareDist2 * inv_sigma2;
          if (chiSquare2 > thH) bIn = false; else c2 = thH - chiSquare2;
        }
      }
      const uint32_t word = __ballot_sync(0xffffffffu, bIn);
      if (lane == 0) mw[base >> 5] = word;
      // the reference adds in match order (score += ... twice per match): replay that order.
      // Directions that failed their test contribute +0.0f, which leaves the fp32 sum unchanged, so
      // only the non-zero terms are visited (most hypotheses have few inliers): same bits, a chain
      // of popc(nz1) + popc(nz2) dependent adds instead of 64 per 32 matches.
      const uint32_t nz1 = __ballot_sync(0xffffffffu, c1 != 0.0f);
      const uint32_t nz2 = __ballot_sync(0xffffffffu, c2 != 0.0f);
      for (uint32_t any = nz1 | nz2; any; any &= any - 1) {
        const int l = __ffs(any) - 1;
        if ((nz1 >> l) & 1u) score += __shfl_sync(0xffffffffu, c1, l);
        if ((nz2 >> l) & 1u) score += __shfl_sync(0xffffffffu, c2, l);
      }
    }
    if (lane == 0) scores[g] = score;
  }
}

// grid = 2 (F, H). best[model] = index of the highest score > 0 (earliest on ties) or -1.
__global__ void tv_argmax_kernel(int n_hyp, const float* __restrict__ scores, int* __restrict__ best_idx,
                                 float* __restrict__ best_score) {
  const int model = blockIdx.x;
  const float* s = scores + (size_t)model * n_hyp;
  __shared__ float ss[256];
  __shared__ int si[256];
  float bs = 0.0f;
  int bi = -1;
  for (int i = threadIdx.x; i < n_hyp; i += blockDim.x) {
    const float v = s[i];
    if (v > bs) { bs = v; bi = i; }  // ascending i per thread: earliest wins inside a thread
  }
  ss[threadIdx.x] = bs;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      const float v = ss[threadIdx.x + off];
      const int j = si[threadIdx.x + off];
      const float cur = ss[threadIdx.x];
      const int ci = si[threadIdx.x];
      if (j >= 0 && (v > cur || (v == cur && (ci < 0 || j < ci)))) { ss[threadIdx.x] = v; si[threadIdx.x] = j; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { best_idx[model] = si[0]; best_score[model] = ss[0]; }
}

// ------------------------------------------------------------------------------- motion recovery


// One CTA per motion hypothesis (grid = 8; every CTA repeats the cheap model selection /
// decomposition, CTA 0 publishes it).  P3D: [8][n1*3], good: [8][n1], cosbuf: [8][N] scratch.
__global__ void __launch_bounds__(256)
tv_motion_kernel(int N, int n1, int n_hyp, const float4* __restrict__ uv, const int* __restrict__ m1,
                 const float* __restrict__ Kg, float th2, const float* __restrict__ models,
                 const uint32_t* __restrict__ masks, const int* __restrict__ best_idx,
                 const float* __restrict__ best_score, float* __restrict__ P3D,
                 uint8_t* __restrict__ good, float* __restrict__ cosbuf, TVMotionOut* __restrict__ out) {
  __shared__ float sR[8][9], st[8][3], sK[9];
  __shared__ int s_nm, s_model, s_cnt[8];
  __shared__ int s_red[256];
  const int words = (N + 31) >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; i++) sK[i] = Kg[i];
    const float SH = best_score[1], SF = best_score[0];
    int nm = 0, used = -1;
    if (SH + SF == 0.f) {
      used = -1;
    } else {
      const float RH = SH / (SH + SF);
      used = (RH > (float)0.50) ? 1 : 0;
      if (used == 0) {
        // _reconstruct_F: E = K^T F K, _decompose_E
        const float* F21 = models + ((size_t)0 * n_hyp + best_idx[0]) * 18;
        float Kt[9], tmp[9], E[9], Fl[9];
        for (int i = 0; i < 9; i++) Fl[i] = F21[i];
        mat3_T(sK, Kt);
        mat3_mul(Kt, Fl, tmp);
        mat3_mul(tmp, sK, E);
        float U[9], w[3], V[9], Vt[9];
        svd3(E, U, w, V);
        mat3_T(V, Vt);
        float t[3] = {U[2], U[5], U[8]};
        const float tn = sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
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
        for (int h = 0; h < 4; h++) {
          const float* Rs = (h & 1) ? R2 : R1;
          for (int i = 0; i < 9; i++) sR[h][i] = Rs[i];
          for (int i = 0; i < 3; i++) st[h][i] = (h < 2) ? t[i] : -t[i];
        }
        nm = 4;
      } else {
        // _reconstruct_H: Faugeras' 8 hypotheses
        const float* H21 = models + ((size_t)1 * n_hyp + best_idx[1]) * 18;
        float invK[9], tmp[9], A[9], Hl[9];
        for (int i = 0; i < 9; i++) Hl[i] = H21[i];
        inv3f(sK, invK);
        mat3_mul(invK, Hl, tmp);
        mat3_mul(tmp, sK, A);
        float U[9], w[3], V[9], Vt[9];
        svd3(A, U, w, V);
        mat3_T(V, Vt);
        const float s = det3(U) * det3(Vt);
        const float d1 = w[0], d2 = w[1], d3 = w[2];
        if ((double)(d1 / d2) < 1.00001 || (double)(d2 / d3) < 1.00001) {
          nm = 0;
        } else {
          const float aux1 = sqrtf((d1 * d1 - d2 * d2) / (d1 * d1 - d3 * d3));
          const float aux3 = sqrtf((d2 * d2 - d3 * d3) / (d1 * d1 - d3 * d3));
          const float x1[4] = {aux1, aux1, -aux1, -aux1};
          const float x3[4] = {aux3, -aux3, aux3, -aux3};
          const float aux_stheta = sqrtf((d1 * d1 - d2 * d2) * (d2 * d2 - d3 * d3)) / ((d1 + d3) * d2);
          const float ctheta = (d2 * d2 + d1 * d3) / ((d1 + d3) * d2);
          const float stheta[4] = {aux_stheta, -aux_stheta, -aux_stheta, aux_stheta};
          float sU[9];
          for (int i = 0; i < 9; i++) sU[i] = s * U[i];
          for (int i = 0; i < 4; i++) {
            const float Rp[9] = {ctheta, 0, -stheta[i], 0, 1.f, 0, stheta[i], 0, ctheta};
            float Rr[9];
            mat3_mul(sU, Rp, tmp);
            mat3_mul(tmp, Vt, Rr);
            for (int k = 0; k < 9; k++) sR[i][k] = Rr[k];
            float tp[3] = {x1[i], 0, -x3[i]};
            for (int k = 0; k < 3; k++) tp[k] *= d1 - d3;
            float t[3];
            for (int r = 0; r < 3; r++) t[r] = U[r * 3] * tp[0] + U[r * 3 + 1] * tp[1] + U[r * 3 + 2] * tp[2];
            const float tn = sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
            for (int r = 0; r < 3; r++) st[i][r] = t[r] / tn;
          }
          const float aux_sphi = sqrtf((d1 * d1 - d2 * d2) * (d2 * d2 - d3 * d3)) / ((d1 - d3) * d2);
          const float cphi = (d1 * d3 - d2 * d2) / ((d1 - d3) * d2);
          const float sphi[4] = {aux_sphi, -aux_sphi, -aux_sphi, aux_sphi};
          for (int i = 0; i < 4; i++) {
            const float Rp[9] = {cphi, 0, sphi[i], 0, -1, 0, sphi[i], 0, -cphi};
            float Rr[9];
            mat3_mul(sU, Rp, tmp);
            mat3_mul(tmp, Vt, Rr);
            for (int k = 0; k < 9; k++) sR[4 + i][k] = Rr[k];
            float tp[3] = {x1[i], 0, x3[i]};
            for (int k = 0; k < 3; k++) tp[k] *= d1 + d3;
            float t[3];
            for (int r = 0; r < 3; r++) t[r] = U[r * 3] * tp[0] + U[r * 3 + 1] * tp[1] + U[r * 3 + 2] * tp[2];
            const float tn = sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
            for (int r = 0; r < 3; r++) st[4 + i][r] = t[r] / tn;
          }
          nm = 8;
        }
      }
    }
    s_nm = nm;
    s_model = used;
    if (blockIdx.x == 0) {
      out->used_H = used;
      out->n_motion = nm;
      for (int h = 0; h < 8; h++) {
        for (int i = 0; i < 9; i++) out->R[h][i] = (h < nm) ? sR[h][i] : 0.0f;
        for (int i = 0; i < 3; i++) out->t[h][i] = (h < nm) ? st[h][i] : 0.0f;
      }
    }
    if (blockIdx.x < 8) {  // every CTA owns the entry of its hypothesis
      out->n_good[blockIdx.x] = 0;
      out->cos_kth[blockIdx.x] = 1.0f;
    }
  }
  __syncthreads();
  const int nm = s_nm, model = s_model;
  if (model < 0) return;
  cosbuf += (size_t)blockIdx.x * N;
  const uint32_t* mask = masks + ((size_t)model * n_hyp + best_idx[model]) * words;
  {  // N inliers of the selected model
    int c = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) c += (mask[i >> 5] >> (i & 31)) & 1u;
    s_red[threadIdx.x] = c;
    __syncthreads();
    for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
      if ((int)threadIdx.x < off) s_red[threadIdx.x] += s_red[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out->n_inl = s_red[0];
    __syncthreads();
  }
  const float fx = sK[0], fy = sK[4], cx = sK[2], cy = sK[5];
  for (int h = blockIdx.x; h < nm; h += gridDim.x) {
    // _check_R_T for motion hypothesis h
    float R[9], t[3], P1[12], P2[12], O2[3];
    for (int i = 0; i < 9; i++) R[i] = sR[h][i];
    for (int i = 0; i < 3; i++) t[i] = st[h][i];
    {
      float Rt[12];
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) { P1[i * 4 + j] = sK[i * 3 + j]; Rt[i * 4 + j] = R[i * 3 + j]; }
        P1[i * 4 + 3] = 0.0f;
        Rt[i * 4 + 3] = t[i];
      }
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++)
          P2[i * 4 + j] = sK[i * 3] * Rt[j] + sK[i * 3 + 1] * Rt[4 + j] + sK[i * 3 + 2] * Rt[8 + j];
      for (int i = 0; i < 3; i++)
        O2[i] = (-R[0 * 3 + i]) * t[0] + (-R[1 * 3 + i]) * t[1] + (-R[2 * 3 + i]) * t[2];
    }
    float* P3Dh = P3D + (size_t)h * n1 * 3;
    uint8_t* goodh = good + (size_t)h * n1;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) {
      goodh[i] = 0;
      P3Dh[i * 3] = 0.0f; P3Dh[i * 3 + 1] = 0.0f; P3Dh[i * 3 + 2] = 0.0f;
    }
    __syncthreads();
    int my_good = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      float cosv = 2.0f;  // marker: not counted
      if ((mask[i >> 5] >> (i & 31)) & 1u) {
        const float4 m = uv[i];
        // _triangulate: null vector of the 4x4 DLT system
        float A[16];
        for (int j = 0; j < 4; j++) {
          A[0 * 4 + j] = m.x * P1[2 * 4 + j] - P1[0 * 4 + j];
          A[1 * 4 + j] = m.y * P1[2 * 4 + j] - P1[1 * 4 + j];
          A[2 * 4 + j] = m.z * P2[2 * 4 + j] - P2[0 * 4 + j];
          A[3 * 4 + j] = m.w * P2[2 * 4 + j] - P2[1 * 4 + j];
        }
        float v[4];
        null_vector(4, 4, A, v);
        float p[3] = {v[0] / v[3], v[1] / v[3], v[2] / v[3]};
        if (isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2])) {
          const float dist1 = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
          const float n2[3] = {p[0] - O2[0], p[1] - O2[1], p[2] - O2[2]};
          const float dist2 = sqrtf(n2[0] * n2[0] + n2[1] * n2[1] + n2[2] * n2[2]);
          const float cosParallax = (p[0] * n2[0] + p[1] * n2[1] + p[2] * n2[2]) / (dist1 * dist2);
          const bool lowpar = (double)cosParallax < 0.99998;
          bool ok = !(p[2] <= 0 && lowpar);
          float p2[3];
          for (int r = 0; r < 3; r++) p2[r] = (R[r * 3] * p[0] + R[r * 3 + 1] * p[1] + R[r * 3 + 2] * p[2]) + t[r];
          if (ok && p2[2] <= 0 && lowpar) ok = false;
          if (ok) {
            const float invZ1 = 1.0f / p[2];
            const float im1x = fx * p[0] * invZ1 + cx;
            const float im1y = fy * p[1] * invZ1 + cy;
            const float squareError1 = (im1x - m.x) * (im1x - m.x) + (im1y - m.y) * (im1y - m.y);
            if (squareError1 > th2) ok = false;
          }
          if (ok) {
            const float invZ2 = 1.0f / p2[2];
            const float im2x = fx * p2[0] * invZ2 + cx;
            const float im2y = fy * p2[1] * invZ2 + cy;
            const float squareError2 = (im2x - m.z) * (im2x - m.z) + (im2y - m.w) * (im2y - m.w);
            if (squareError2 > th2) ok = false;
          }
          if (ok) {
            cosv = cosParallax;
            const int k1 = m1[i];
            P3Dh[k1 * 3] = p[0]; P3Dh[k1 * 3 + 1] = p[1]; P3Dh[k1 * 3 + 2] = p[2];
            my_good++;
            if (lowpar) goodh[k1] = 1;
          }
        }
      }
      cosbuf[i] = cosv;
    }
    s_red[threadIdx.x] = my_good;
    __syncthreads();
    for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
      if ((int)threadIdx.x < off) s_red[threadIdx.x] += s_red[threadIdx.x + off];
      __syncthreads();
    }
    const int nGood = s_red[0];
    __syncthreads();
    if (threadIdx.x == 0) { out->n_good[h] = nGood; s_cnt[h] = nGood; }
    if (nGood > 0) {
      // element of rank min(50, nGood-1) of the sorted cosines, by rank counting
      const int kth = min(50, nGood - 1);
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float ci = cosbuf[i];
        if (ci > 1.5f) continue;
        int rank = 0;
        for (int j = 0; j < N; j++) {
          const float cj = cosbuf[j];
          rank += (cj < ci || (cj == ci && j < i)) ? 1 : 0;
        }
        if (rank == kth) out->cos_kth[h] = ci;
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_tv_ransac(const TVBuffers& b, float sigma, int score_mode, int n_sm, cudaStream_t stream, int* n_launch) {
  int nl = 0;
  tv_normalize_kernel<<<2, 256, 0, stream>>>(b.n1, b.keys1, b.pn1, b.T1, b.n2, b.keys2, b.pn2, b.T2);
  nl++;
  tv_gather_kernel<<<(b.N + 255) / 256, 256, 0, stream>>>(b.N, b.m1, b.m2, b.keys1, b.keys2, b.pn1, b.pn2, b.uv, b.pnm);
  nl++;
  {
    // cooperative fit: 4 fundamental / 2 homography hypotheses per warp, 8 warps per CTA
    const int wpc = 8;
    const int gF = (b.n_hyp + wpc * 2 - 1) / (wpc * 2), gH = (b.n_hyp + wpc * 2 - 1) / (wpc * 2);
    const size_t smF = (size_t)wpc * 2 * (8 * 9) * sizeof(float), smH = (size_t)wpc * 2 * (16 * 9 + 81) * sizeof(float);
    tv_fit_sub_kernel<0><<<gF, wpc * 32, smF, stream>>>(b.n_hyp, b.sets, b.pnm, b.T1, b.T2, b.models);
    tv_fit_sub_kernel<1><<<gH, wpc * 32, smH, stream>>>(b.n_hyp, b.sets, b.pnm, b.T1, b.T2, b.models);
    nl += 2;
  }
  const float inv_sigma2 = 1.0f / (sigma * sigma);
  const size_t smem = (size_t)b.N * sizeof(float4);
  const int stage = smem <= 160 * 1024 ? 1 : 0;
  if (stage && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(tv_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int wpc = 8;
  int grid = (2 * b.n_hyp + wpc - 1) / wpc;
  const int cap = n_sm * 8;
  if (grid > cap) grid = cap;
  tv_score_kernel<<<grid, wpc * 32, stage ? smem : 0, stream>>>(b.N, b.n_hyp, b.uv, b.models, inv_sigma2, stage, score_mode, b.scores, b.masks);
  nl++;
  tv_argmax_kernel<<<2, 256, 0, stream>>>(b.n_hyp, b.scores, b.best_idx, b.best_score);
  nl++;
  if (n_launch) *n_launch = nl;
  return cudaGetLastError();
}

cudaError_t launch_tv_motion(const TVBuffers& b, float th2, cudaStream_t stream) {
  tv_motion_kernel<<<8, 256, 0, stream>>>(b.N, b.n1, b.n_hyp, b.uv, b.m1, b.K, th2, b.models, b.masks,
                                          b.best_idx, b.best_score, b.P3D, b.good, b.cosbuf, b.motion);
  return cudaGetLastError();
}

}  // namespace urmvo
