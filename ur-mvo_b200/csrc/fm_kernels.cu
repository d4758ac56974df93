// csrc/fm_kernels.cu — per-frame fundamental-matrix RANSAC on sm_100a.
//
// Reference behaviour reproduced: src/point_matching.cc:44-58,
//     cv::findFundamentalMat(points0, points1, cv::FM_RANSAC, 3, 0.99, inliers)
// i.e. OpenCV's RANSACPointSetRegistrator with the 7-point solver (calib3d fundam.cpp / ptsetreg.cpp;
// OpenCV is linked, not vendored, by the reference).  The sequential RANSAC loop is data-parallel in
// everything except two things, which stay on the host (capi.cu):
//   * the cv::RNG subset draws (a 64-bit multiply-with-carry chain with data-dependent re-draws),
//   * the "strictly more inliers than the best so far -> shrink the iteration budget" replay, which
//     needs only the inlier COUNT of every (iteration, model) and picks the same winner as the
//     sequential loop.
// The device evaluates every iteration of the budget at once:
//   fm_solve_kernel  16 lanes per iteration: run7Point (normalise, 7x9 system, 2-D null space by a
//                    Householder QR in fp64 registers, cubic det = 0, <= 3 models)      [fp64 latency]
//   fm_score_kernel  one warp per (iteration, model): FMEstimatorCallback::computeError in fp64,
//                    (float)err <= (float)thr^2, ballot/popc inlier count               [fp64 ALU]
//   fm_mask_kernel   the winner's inlier flags
// Arithmetic follows the CPU restatement fm_oracle.cpp operation by operation (same summation
// order: ordered shuffle chains / 16-lane butterflies; this file is compiled with -fmad=false), so the models agree to the
// last bits and the masks are identical; the oracle itself is pinned bit-for-bit against the real
// cv2.findFundamentalMat (tests/golden/golden_fm_r01.npz).
#include <cfloat>
#include <cmath>

#include "kernels.h"

namespace urmvo {
namespace {

constexpr int kFmLanes = 16;       // lanes per hypothesis in fm_solve_kernel

// sum over lanes 0..m-1 of the sub-warp, in lane order (the oracle's loop order)
__device__ __forceinline__ double ordered_sum64(double v, int m, unsigned mask) {
  double s = __shfl_sync(mask, v, 0, kFmLanes);
  for (int k = 1; k < m; k++) s = s + __shfl_sync(mask, v, k, kFmLanes);
  return s;
}

// 16-lane xor butterfly: every lane ends with the pairwise-tree sum ((v0+v1)+(v2+v3))+...
__device__ __forceinline__ double tree_sum16(double v, unsigned mask) {
#pragma unroll
  for (int w = 1; w < 16; w <<= 1) v = v + __shfl_xor_sync(mask, v, w, kFmLanes);
  return v;
}

// cv::solveCubic (c[0] x^3 + c[1] x^2 + c[2] x + c[3] = 0); returns the number of roots, -1: any x
__device__ int solve_cubic(const double* c, double* r) {
  double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
  int n = 0;
  double x0 = 0, x1 = 0, x2 = 0;
  if (a0 == 0) {
    if (a1 == 0) {
      if (a2 == 0) {
        n = a3 == 0 ? -1 : 0;
      } else {
        x0 = -a3 / a2;
        n = 1;
      }
    } else {
      double d = a2 * a2 - 4 * a1 * a3;
      if (d >= 0) {
        d = sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) {
          x0 = q1 / a1;
          x1 = a3 / q1;
        } else {
          x0 = q2 / a1;
          x1 = a3 / q2;
        }
        n = d > 0 ? 2 : 1;
      }
    }
  } else {
    a0 = 1. / a0;
    a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d > 0) {
      const double theta = acos(R / sqrt(Qcubed));
      const double sqrtQ = sqrt(Q);
      const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
      x0 = t0 * cos(t1) - t2;
      x1 = t0 * cos(t1 + (2. * 3.14159265358979323846 / 3)) - t2;
      x2 = t0 * cos(t1 + (4. * 3.14159265358979323846 / 3)) - t2;
      n = 3;
    } else if (d == 0) {
      if (R >= 0) {
        x0 = -2 * pow(R, 1. / 3) - a1 / 3;
        x1 = pow(R, 1. / 3) - a1 / 3;
      } else {
        x0 = 2 * pow(-R, 1. / 3) - a1 / 3;
        x1 = -pow(-R, 1. / 3) - a1 / 3;
      }
      x2 = 0;
      n = x0 == x1 ? 1 : 2;
      x1 = x0 == x1 ? 0 : x1;
    } else {
      d = sqrt(-d);
      double e = pow(d + fabs(R), 1. / 3);
      if (R > 0) e = -e;
      x0 = (e + Q / e) - a1 * (1. / 3);
      n = 1;
    }
  }
  r[0] = x0; r[1] = x1; r[2] = x2;
  return n;
}

// One RANSAC round: n hypotheses.  hyp_ids[i] = problem * max_iters + iteration addresses the
// persistent model store (models: [B*max_iters][27]); sets / n_models are compact per round.
// pts: float4 (x0,y0,x1,y1) per match, sets hold GLOBAL match indices (problem offset added).
__global__ void __launch_bounds__(256)
fm_solve_kernel(int n_hyp, const int* __restrict__ hyp_ids, const int* __restrict__ sets,
                const float4* __restrict__ pts, double* __restrict__ models_all, int* __restrict__ n_models) {
  constexpr int PER_WARP = 32 / kFmLanes;
  constexpr int DL = 7 * 9;  // doubles per hypothesis: the 7x9 system, later the two null vectors
  extern __shared__ double sm_fm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane / kFmLanes, r = lane - sub * kFmLanes;
  const unsigned mask = 0xFFFFu << (sub * kFmLanes);
  const int hyp = (blockIdx.x * (blockDim.x >> 5) + wid) * PER_WARP + sub;
  if (hyp >= n_hyp) return;  // whole sub-warps leave together
  double* A = sm_fm + (size_t)(wid * PER_WARP + sub) * DL;
  // ---- run7Point: normalisation (centroid, mean distance) in the oracle's summation order
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < 7) m = pts[sets[(size_t)hyp * 7 + r]];
  const double t = 1. / 7;
  const double c1x = ordered_sum64((double)m.x, 7, mask) * t, c1y = ordered_sum64((double)m.y, 7, mask) * t;
  const double c2x = ordered_sum64((double)m.z, 7, mask) * t, c2y = ordered_sum64((double)m.w, 7, mask) * t;
  const double dx1 = m.x - c1x, dy1 = m.y - c1y, dx2 = m.z - c2x, dy2 = m.w - c2y;
  double scale1 = ordered_sum64(sqrt(dx1 * dx1 + dy1 * dy1), 7, mask) * t;
  double scale2 = ordered_sum64(sqrt(dx2 * dx2 + dy2 * dy2), 7, mask) * t;
  if (scale1 < FLT_EPSILON || scale2 < FLT_EPSILON) {
    if (r == 0) n_models[hyp] = 0;
    return;
  }
  scale1 = sqrt(2.) / scale1;
  scale2 = sqrt(2.) / scale2;
  if (r < 7) {
    const double x0 = dx1 * scale1, y0 = dy1 * scale1, x1 = dx2 * scale2, y1 = dy2 * scale2;
    double* a = A + r * 9;
    a[0] = x1 * x0; a[1] = x1 * y0; a[2] = x1;
    a[3] = y1 * x0; a[4] = y1 * y0; a[5] = y1;
    a[6] = x0; a[7] = y0; a[8] = 1;
  }
  __syncwarp(mask);
  // ---- 2-D null space: Householder QR of M = A^T (9x7).  Lane r holds row r of M (lanes >= 9:
  // zeros); reflector k: sigma = sum x^2 over rows >= k (16-lane xor butterfly), alpha = -sign(x_k)
  // sqrt(sigma), v = x - alpha e_k, beta = 1/(norm (norm + |x_k|)); the last two columns of
  // Q = H_0 ... H_6 span the null space (f1 = Q e_8, f2 = Q e_7).  Spec: fm_oracle.cpp null_space_7x9.
  double mrow[7], vk[7], bk[7];
#pragma unroll
  for (int k = 0; k < 7; k++) mrow[k] = r < 9 ? A[k * 9 + r] : 0.0;
#pragma unroll
  for (int k = 0; k < 7; k++) {
    const double x = r >= k ? mrow[k] : 0.0;
    const double sigma = tree_sum16(x * x, mask);
    const double xkk = __shfl_sync(mask, mrow[k], k, kFmLanes);
    const double norm = sqrt(sigma);
    const double alpha = xkk >= 0.0 ? -norm : norm;
    const double v = r == k ? x - alpha : x;
    const double beta = norm > 0.0 ? 1.0 / (norm * (norm + fabs(xkk))) : 0.0;
    vk[k] = v;
    bk[k] = beta;
#pragma unroll
    for (int j = k + 1; j < 7; j++) {
      const double w = beta * tree_sum16(v * mrow[j], mask);
      mrow[j] = mrow[j] - w * v;
    }
  }
  double y1 = r == 8 ? 1.0 : 0.0, y2 = r == 7 ? 1.0 : 0.0;
#pragma unroll
  for (int k = 6; k >= 0; k--) {
    const double w1 = bk[k] * tree_sum16(vk[k] * y1, mask);
    const double w2 = bk[k] * tree_sum16(vk[k] * y2, mask);
    y1 = y1 - w1 * vk[k];
    y2 = y2 - w2 * vk[k];
  }
  __syncwarp(mask);
  if (r < 9) { A[r] = y1; A[9 + r] = y2; }  // the staging area is free now
  __syncwarp(mask);
  if (r != 0) return;
  double f1[9], f2[9];
  for (int k = 0; k < 9; k++) { f1[k] = A[k]; f2[k] = A[9 + k]; }
  for (int k = 0; k < 9; k++) f1[k] -= f2[k];
  double c[4], roots[3];
  double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
  c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
  c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
         f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
         f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
         f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
         f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
  int n = solve_cubic(c, roots);
  if (n < 1 || n > 3) {
    n_models[hyp] = n < 0 ? 0 : n;
    return;
  }
  const double T1[9] = {scale1, 0, -scale1 * c1x, 0, scale1, -scale1 * c1y, 0, 0, 1};
  const double T2[9] = {scale2, 0, -scale2 * c2x, 0, scale2, -scale2 * c2y, 0, 0, 1};
  for (int k = 0; k < n; k++) {
    double* fm = models_all + (size_t)hyp_ids[hyp] * 27 + 9 * k;
    double lambda = roots[k], mu = 1.;
    const double s = f1[8] * roots[k] + f2[8];
    double Fn[9];
    if (fabs(s) > DBL_EPSILON) {
      mu = 1. / s;
      lambda *= mu;
      Fn[8] = 1.;
    } else {
      Fn[8] = 0.;
    }
    for (int i = 0; i < 8; i++) Fn[i] = f1[i] * lambda + f2[i] * mu;
    double tmp[9], out[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double a = 0;
        for (int l = 0; l < 3; l++) a += T2[l * 3 + i] * Fn[l * 3 + j];
        tmp[i * 3 + j] = a;
      }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double a = 0;
        for (int l = 0; l < 3; l++) a += tmp[i * 3 + l] * T1[l * 3 + j];
        out[i * 3 + j] = a;
      }
    if (fabs(out[8]) > FLT_EPSILON) {
      const double inv = 1. / out[8];
      for (int i = 0; i < 9; i++) out[i] *= inv;
    }
    for (int i = 0; i < 9; i++) fm[i] = out[i];
  }
  n_models[hyp] = n;
}

// FMEstimatorCallback::computeError + the threshold test of findInliers for one correspondence:
//   err = (float)std::max(d1^2 * (1/(a1^2+b1^2)), d2^2 * (1/(a2^2+b2^2)));  inlier = err <= (float)thr^2
// Fast path without the two fp64 divisions: (float)x <= t  <=>  x <= tm (tm = the double midpoint
// between t and the next float; passed in by the host), and x = d^2/den is compared as
// d^2 vs tm*den with a relative guard band of 1e-12; anything inside the band (or den == 0) takes
// the exact path, which repeats OpenCV's expression operation by operation.  The decision is
// therefore always the exact one.
struct FmThresh {
  float t;      // (float)(thresh*thresh)
  double lo;    // tm * (1 - 1e-12)
  double hi;    // tm * (1 + 1e-12)
};

// FMEstimatorCallback::computeError for one correspondence, OpenCV's expression operation by operation.
__device__ __forceinline__ float fm_error_exact(const double* F, const float4 m) {
  const double x1 = m.x, y1 = m.y, x2 = m.z, y2 = m.w;
  const double a2 = F[0] * x1 + F[1] * y1 + F[2];
  const double b2 = F[3] * x1 + F[4] * y1 + F[5];
  const double c2 = F[6] * x1 + F[7] * y1 + F[8];
  const double s2 = 1. / (a2 * a2 + b2 * b2);
  const double d2 = x2 * a2 + y2 * b2 + c2;
  const double a1 = F[0] * x2 + F[3] * y2 + F[6];
  const double b1 = F[1] * x2 + F[4] * y2 + F[7];
  const double c1 = F[2] * x2 + F[5] * y2 + F[8];
  const double s1 = 1. / (a1 * a1 + b1 * b1);
  const double d1 = x1 * a1 + y1 * b1 + c1;
  const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
  return (float)((e1 < e2) ? e2 : e1);  // std::max(e1, e2)
}

__device__ __forceinline__ FmThresh fm_make_thresh(float t) {
  // (float)x <= t  <=>  x <= midpoint(t, nextafterf(t, +inf)) up to the tie, which the exact path decides
  const double tm = 0.5 * ((double)t + (double)nextafterf(t, INFINITY));
  FmThresh th;
  th.t = t;
  th.lo = tm * (1.0 - 1e-12);
  th.hi = tm * (1.0 + 1e-12);
  return th;
}

__device__ __forceinline__ bool fm_inlier(const double* F, const float4 m, const FmThresh th) {
  const double x1 = m.x, y1 = m.y, x2 = m.z, y2 = m.w;
  const double a2 = F[0] * x1 + F[1] * y1 + F[2];
  const double b2 = F[3] * x1 + F[4] * y1 + F[5];
  const double c2 = F[6] * x1 + F[7] * y1 + F[8];
  const double den2 = a2 * a2 + b2 * b2;
  const double d2 = x2 * a2 + y2 * b2 + c2;
  const double a1 = F[0] * x2 + F[3] * y2 + F[6];
  const double b1 = F[1] * x2 + F[4] * y2 + F[7];
  const double c1 = F[2] * x2 + F[5] * y2 + F[8];
  const double den1 = a1 * a1 + b1 * b1;
  const double d1 = x1 * a1 + y1 * b1 + c1;
  const double n1 = d1 * d1, n2 = d2 * d2;
  // clearly outside in either direction -> outlier; clearly inside in both -> inlier
  const bool out = n1 > th.hi * den1 || n2 > th.hi * den2;
  const bool in = n1 < th.lo * den1 && n2 < th.lo * den2;
  if (in) return true;
  if (out && den1 > 0.0 && den2 > 0.0) return false;
  // exact path (guard band, zero denominators, NaNs): OpenCV's own expression
  const double s2 = 1. / den2, s1 = 1. / den1;
  const double e1 = n1 * s1, e2 = n2 * s2;
  const float err = (float)((e1 < e2) ? e2 : e1);  // std::max(e1, e2)
  return err <= th.t;
}

// One warp per (round hypothesis, model slot).  counts: [n_hyp][3] compact (0 for unused slots).
__global__ void __launch_bounds__(256)
fm_score_kernel(int n_hyp, const int* __restrict__ hyp_ids, int max_iters, const int* __restrict__ off,
                const float4* __restrict__ pts, const double* __restrict__ models_all,
                const int* __restrict__ n_models, FmThresh th, int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  for (int g = blockIdx.x * wpc + (threadIdx.x >> 5); g < n_hyp * 3; g += gridDim.x * wpc) {
    const int i = g / 3, k = g - i * 3;
    if (k >= n_models[i]) {
      if (lane == 0) counts[g] = 0;
      continue;
    }
    const int id = hyp_ids[i];
    double F[9];
#pragma unroll
    for (int e = 0; e < 9; e++) F[e] = models_all[(size_t)id * 27 + 9 * k + e];
    const int b = id / max_iters;
    const int i0 = off[b], i1 = off[b + 1];
    if (i1 - i0 < 15) {
      // fewer than 15 matches: cv::findFundamentalMat runs LMeDSPointSetRegistrator::run — the score of a model is
      // the count/2-th smallest float error, ordered as integers (std::nth_element on errf.ptr<int>()).  counts[g]
      // carries the bits of that median.
      const int n = i1 - i0;
      const int mine = lane < n ? __float_as_int(fm_error_exact(F, pts[i0 + lane])) : 0x7fffffff;
      int rank = 0;
      for (int j = 0; j < n; j++) {
        const int other = __shfl_sync(0xffffffffu, mine, j);
        rank += (other < mine || (other == mine && j < lane)) ? 1 : 0;
      }
      if (lane < n && rank == n / 2) counts[g] = mine;
      continue;
    }
    int good = 0;
    for (int base = i0; base < i1; base += 32) {
      const int j = base + lane;
      bool in = false;
      if (j < i1) in = fm_inlier(F, pts[j], th);
      good += __popc(__ballot_sync(0xffffffffu, in));
    }
    if (lane == 0) counts[g] = good;
  }
}

// Inlier flags of the winning model of every problem; win_id[b] = hyp_id*3 + model slot or -1 (no
// model: all flags 0).  Also gathers the winning models into win_F [B][9].  thr_b[b] = the problem's own squared
// threshold as a float (the LMedS sigma^2 of problems with fewer than 15 matches); a negative value flags every
// match (the direct 7-point branch: mask.setTo(1)); thr_b == NULL: th for all.
__global__ void __launch_bounds__(256)
fm_mask_kernel(int B, const int* __restrict__ off, const float4* __restrict__ pts,
               const double* __restrict__ models_all, const int* __restrict__ win_id, FmThresh th,
               const float* __restrict__ thr_b, uint8_t* __restrict__ mask, double* __restrict__ win_F) {
  const int b = blockIdx.y;
  if (b >= B) return;
  const int i0 = off[b], i1 = off[b + 1];
  const int w = win_id[b];
  bool all = false;
  if (thr_b) {
    const float t = thr_b[b];
    all = t < 0.f;
    if (!all) th = fm_make_thresh(t);
  }
  double F[9];
#pragma unroll
  for (int i = 0; i < 9; i++) F[i] = w >= 0 ? models_all[(size_t)(w / 3) * 27 + 9 * (w % 3) + i] : 0.0;
  if (blockIdx.x == 0 && threadIdx.x < 9) win_F[(size_t)b * 9 + threadIdx.x] = F[threadIdx.x];
  for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += gridDim.x * blockDim.x)
    mask[i] = (all || (w >= 0 && fm_inlier(F, pts[i], th))) ? 1 : 0;
}

}  // namespace

static FmThresh make_thresh(float t) {
  // (float)x <= t  <=>  x <= midpoint(t, nextafterf(t, +inf)) up to the tie, which the exact path decides
  const double tm = 0.5 * ((double)t + (double)nextafterf(t, INFINITY));
  FmThresh th;
  th.t = t;
  th.lo = tm * (1.0 - 1e-12);
  th.hi = tm * (1.0 + 1e-12);
  return th;
}

cudaError_t launch_fm_solve(int n_hyp, const int* hyp_ids, const int* sets, const float4* pts, double* models_all,
                            int* n_models, cudaStream_t s) {
  if (n_hyp <= 0) return cudaSuccess;
  const int threads = 256, per_cta = (threads / 32) * (32 / kFmLanes);
  const size_t smem = (size_t)per_cta * 63 * sizeof(double);
  fm_solve_kernel<<<(n_hyp + per_cta - 1) / per_cta, threads, smem, s>>>(n_hyp, hyp_ids, sets, pts, models_all, n_models);
  return cudaGetLastError();
}

cudaError_t launch_fm_score(int n_hyp, const int* hyp_ids, int max_iters, const int* off, const float4* pts,
                            const double* models_all, const int* n_models, float thr2, int* counts, int n_sm,
                            cudaStream_t s) {
  if (n_hyp <= 0) return cudaSuccess;
  const int threads = 256, wpc = threads / 32;
  long long blocks = ((long long)n_hyp * 3 + wpc - 1) / wpc;
  const long long cap = (long long)n_sm * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  fm_score_kernel<<<(unsigned)blocks, threads, 0, s>>>(n_hyp, hyp_ids, max_iters, off, pts, models_all, n_models,
                                                       make_thresh(thr2), counts);
  return cudaGetLastError();
}

cudaError_t launch_fm_mask(int B, int max_n, const int* off, const float4* pts, const double* models_all,
                           const int* win_id, float thr2, const float* thr_b, uint8_t* mask, double* win_F,
                           cudaStream_t s) {
  if (B <= 0 || max_n <= 0) return cudaSuccess;
  dim3 grid((unsigned)((max_n + 255) / 256), (unsigned)B);
  fm_mask_kernel<<<grid, 256, 0, s>>>(B, off, pts, models_all, win_id, make_thresh(thr2), thr_b, mask, win_F);
  return cudaGetLastError();
}

}  // namespace urmvo
