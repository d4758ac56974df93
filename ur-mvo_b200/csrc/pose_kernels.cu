// csrc/pose_kernels.cu — batched pose-only optimisation (FrameOptimization) on sm_100a.
//
// Reference behaviour reproduced:
//   /root/reference/src/g2o_optimization.cc:179-321: 4 rounds, each restarting from the INPUT pose
//   (:265-266), optimize(10) with g2o's Levenberg-Marquardt, re-classification with a float-cast
//   chi2 (:277), robust kernel removed after round index 2 (:287-288), early exit for < 10 edges.
//   g2o EdgeSE3ProjectXYZOnlyPose error / Jacobian (upstream types_six_dof_expmap.cpp).
//
// One CTA per frame; the whole 4 x 10-iteration protocol runs inside one launch.  Each LM trial is
// two passes over the frame's observations (40 B each, L1/L2 resident after the first pass): one
// that accumulates the 6x6 normal equations (21 + 6 + 1 sums, fixed-order CTA reduction) and one
// that evaluates the trial cost.  The 6x6 damped system is solved by Cholesky in every thread
// (block-Jacobi PCG on a single 6x6 block is exactly one direct solve).

#include "ba_types.h"
#include "common.cuh"
#include "kernels.h"

namespace urmvo {

namespace {

__device__ __forceinline__ void pose_map(const double* R, const double* t, const double* X, double* pc) {
  pc[0] = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  pc[1] = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  pc[2] = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
}

// CTA-wide sum of the 28 accumulators of the normal equations (result in every thread), fixed order.
// Inside a warp the 28 (padded to 32) values are reduced "scatter" fashion: at the step with lane distance o a lane
// keeps one half of its values and exchanges the other half with lane ^ o, so the five steps move 16 + 8 + 4 + 2 + 1
// = 31 values per lane instead of 5 x 28, and lane k ends with the warp total of value k.  Warp totals are then
// added in warp order by thread k.  red: >= 29 * 32 doubles.
// (Measured and NOT adopted: evaluating the normal equations in the trial-cost pass, so that an accepted trial is
// already linearised for the next iteration — identical results, but the heavier trial pass cost more than the saved
// error passes: 0.278 -> 0.314 ms for one 1000-match frame.  512 threads per frame (two observations per thread and
// pass, 128 registers): 0.281 -> 0.287 ms for one frame, 0.34 -> 0.44 ms for 100 — the spills and the longer CTA
// reduction cost more than the shorter pass saves.)
__device__ __forceinline__ void block_sum28(double (&v)[28], double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  double a[16];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const double lo = v[j], hi = j + 16 < 28 ? v[j + 16] : 0.0;
      const double send = up ? lo : hi, keep = up ? hi : lo;
      a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; j++) {
      const double send = up ? a[j] : a[j + o], keep = up ? a[j + o] : a[j];
      a[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  __syncthreads();  // the previous totals in red have been read by everybody
  if (lane < 28) red[lane * 32 + wid] = a[0];
  __syncthreads();
  if (threadIdx.x < 28) {
    double t = 0.0;
    for (int w = 0; w < nw; w++) t += red[threadIdx.x * 32 + w];
    red[28 * 32 + threadIdx.x] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 28; k++) v[k] = red[28 * 32 + k];
}

// One reciprocal per edge (fp64 division is a software routine on the SM); see ba_kernels.cu.
__device__ __forceinline__ double pose_err(const double* pc, double u, double v, const double* K,
                                           double& e0, double& e1) {
  const double iz = 1.0 / pc[2];
  e0 = u - (pc[0] * iz * K[0] + K[2]);
  e1 = v - (pc[1] * iz * K[1] + K[3]);
  return e0 * e0 + e1 * e1;
}

// EdgeSE3ProjectXYZOnlyPose::linearizeOplus (invz form)
__device__ __forceinline__ void pose_jac(const double* pc, const double* K, double* J) {
  const double x = pc[0], y = pc[1];
  const double invz = 1.0 / pc[2], invz_2 = invz * invz;
  J[0] = x * y * invz_2 * K[0];
  J[1] = -(1 + (x * x * invz_2)) * K[0];
  J[2] = y * invz * K[0];
  J[3] = -invz * K[0];
  J[4] = 0;
  J[5] = x * invz_2 * K[0];
  J[6] = (1 + y * y * invz_2) * K[1];
  J[7] = -x * y * invz_2 * K[1];
  J[8] = -x * invz * K[1];
  J[9] = 0;
  J[10] = -invz * K[1];
  J[11] = y * invz_2 * K[1];
}

// EdgeStereoSE3ProjectXYZOnlyPose (src/g2o_optimization.cc:235-258): third residual row
// u_right - (u_left_projected - bf / z) and its Jacobian row (row 0 plus the bf terms).
__device__ __forceinline__ double pose_err_right(const double* pc, double ur, const double* K, double bf) {
  const double iz = 1.0 / pc[2];
  return ur - (pc[0] * iz * K[0] + K[2] - bf * iz);
}
__device__ __forceinline__ void pose_jac_right(const double* pc, const double* J, double bf, double* J2) {
  const double x = pc[0], y = pc[1];
  const double invz = 1.0 / pc[2], invz_2 = invz * invz;
  J2[0] = J[0] - bf * y * invz_2;
  J2[1] = J[1] + bf * x * invz_2;
  J2[2] = J[2];
  J2[3] = J[3];
  J2[4] = 0;
  J2[5] = J[5] - bf * invz_2;
}

// Solve (H + lambda I) x = b, H symmetric packed upper (21). False if not positive definite.
// LDL^T with reciprocal pivots: every thread of the CTA runs this redundantly between two passes over
// the observations, so its serial latency is what matters — 6 divisions on the critical path instead
// of the 6 square roots + 18 divisions of a textbook Cholesky with substitutions.
__device__ __forceinline__ bool solve6(const double* Hp, const double* b, double lambda, double* x) {
  double A[36], L[36], dinv[6], dd[6];
  int idx = 0;
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = i; j < 6; j++) { A[i * 6 + j] = Hp[idx]; A[j * 6 + i] = Hp[idx]; idx++; }
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double v[6];
    double d = A[j * 6 + j] + lambda;
#pragma unroll
    for (int k = 0; k < j; k++) { v[k] = L[j * 6 + k] * dd[k]; d -= L[j * 6 + k] * v[k]; }
    if (!(d > 0.0)) return false;
    dd[j] = d;
    dinv[j] = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double s = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[i * 6 + k] * v[k];
      L[i * 6 + j] = s * dinv[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {  // z = L^-1 b (unit lower triangle)
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[i * 6 + k] * x[k];
    x[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) x[i] *= dinv[i];
#pragma unroll
  for (int i = 5; i >= 0; i--) {  // x = L^-T y
    double s = x[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * x[k];
    x[i] = s;
  }
  return true;
}

}  // namespace

// level: per-observation scratch (0 active / 1 excluded).  inlier: in/out flags.
// MINB = CTAs per SM the register allocation is capped for: 1 keeps everything in registers (172) and
// is fastest when the batch fits one wave (B <= #SMs, and for single-frame latency); 2 caps at 128
// registers (a few spills) so that a batch of up to 2 x #SMs frames is resident at once instead of
// running a second, partly empty wave — +34 % on 256 frames.
// STEREO: the frame mixes mono edges and stereo edges (kind[o] = 1: measurement (u, v, u_right), 3-row
// residual, Huber delta / threshold of cfg.stereo_point) and reads the intrinsics of every edge from its camera
// model row (camera_list[mpc->id_camera]); the mono-only single-camera instantiation is unchanged.
struct PoseStereo {
  const double* ur;     // per observation, read for stereo edges only
  const uint8_t* kind;  // per observation: stereo bit | camera model << 1
  const double* tab;    // camera models, rows of (fx fy cx cy bf)
  double chi2_thr, delta;
};

template <int MINB, bool STEREO>
__global__ void __launch_bounds__(256, MINB)
pose_only_kernel(int B, const int* __restrict__ obs_off, const double* __restrict__ pose_in,
                 const double* __restrict__ uv, const double* __restrict__ Xw, double fx, double fy,
                 double cx, double cy, double chi2_thr, double delta, int rounds, int its_per_round,
                 uint8_t* __restrict__ inlier, uint8_t* __restrict__ level,
                 double* __restrict__ pose_out, int* __restrict__ n_inlier, int* __restrict__ lm_iters,
                 PoseStereo sp) {
  __shared__ double red[29 * 32 + 32];
  const double K[4] = {fx, fy, cx, cy};
  for (int f = blockIdx.x; f < B; f += gridDim.x) {
    const int o0 = obs_off[f], n = obs_off[f + 1] - o0;
    const double* fuv = uv + (size_t)o0 * 2;
    const double* fur = STEREO ? sp.ur + o0 : nullptr;
    const uint8_t* fkind = STEREO ? sp.kind + o0 : nullptr;
    const double* fX = Xw + (size_t)o0 * 3;
    uint8_t* flev = level + o0;
    uint8_t* finl = inlier + o0;
    // T0 = SE3Quat(q, p).inverse()  (:198-199)
    double q0[4], t0[3];
    {
      const double* in = pose_in + (size_t)f * 7;
      double q[4] = {in[0], in[1], in[2], in[3]}, t[3] = {in[4], in[5], in[6]};
      quat_normalize_w(q);
      se3_inverse(q, t, q0, t0);
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) flev[i] = 0;
    __syncthreads();
    double qc[4], tc[3];  // current estimate
    double ql[4], tl[3];  // state of the last computeActiveErrors()
    int num_outlier = 0, total_iters = 0;
    for (int round = 0; round < rounds; round++) {
      const bool robust = (round <= 2);  // kernel removed after round index 2 (:287-288)
#pragma unroll
      for (int a = 0; a < 4; a++) { qc[a] = q0[a]; ql[a] = q0[a]; }
#pragma unroll
      for (int a = 0; a < 3; a++) { tc[a] = t0[a]; tl[a] = t0[a]; }
      // any level-0 edge? (g2o: no active vertex -> optimize() does nothing)
      double na[1] = {0.0};
      for (int i = threadIdx.x; i < n; i += blockDim.x) na[0] += flev[i] ? 0.0 : 1.0;
      block_sum<1>(na, red);
      bool evaluated = false;
      if (na[0] > 0.0) {
        double lambda = 0.0, ni = 2.0;
        for (int it = 0; it < its_per_round; it++) {
          double R[9];
          quat_to_R(qc, R);
          // computeActiveErrors + buildSystem: H (21, packed upper), b (6), robust chi2
          double acc[28];
#pragma unroll
          for (int a = 0; a < 28; a++) acc[a] = 0.0;
          for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (flev[i]) continue;
            const double X[3] = {fX[i * 3], fX[i * 3 + 1], fX[i * 3 + 2]};
            double pc[3], e0, e1, w, J[12];
            pose_map(R, tc, X, pc);
            const unsigned kb = STEREO ? fkind[i] : 0u;  // stereo bit | camera model << 1
            const bool st = STEREO && (kb & 1u);
            const double* Ke = STEREO ? sp.tab + 5 * (kb >> 1) : K;  // camera_list[mpc->id_camera] (:221-224)
            double e2 = pose_err(pc, fuv[i * 2], fuv[i * 2 + 1], Ke, e0, e1);
            double er = 0.0;
            if (st) { er = pose_err_right(pc, fur[i], Ke, Ke[4]); e2 += er * er; }
            acc[27] += huber_rho(e2, st ? sp.delta : delta, robust, w);
            pose_jac(pc, Ke, J);
            const double r0 = -w * e0, r1 = -w * e1;
            int idx = 0;
#pragma unroll
            for (int a = 0; a < 6; a++) {
              acc[21 + a] += J[a] * r0 + J[6 + a] * r1;
#pragma unroll
              for (int b = a; b < 6; b++) { acc[idx] += w * (J[a] * J[b] + J[6 + a] * J[6 + b]); idx++; }
            }
            if (st) {
              double J2[6];
              pose_jac_right(pc, J, Ke[4], J2);
              const double r2 = -w * er;
              idx = 0;
#pragma unroll
              for (int a = 0; a < 6; a++) {
                acc[21 + a] += J2[a] * r2;
#pragma unroll
                for (int b = a; b < 6; b++) { acc[idx] += w * (J2[a] * J2[b]); idx++; }
              }
            }
          }
          block_sum28(acc, red);
          double currentChi = acc[27];
          if (it == 0) {
            double m = 0.0;
            int idx = 0;
#pragma unroll
            for (int a = 0; a < 6; a++) { m = fmax(m, fabs(acc[idx])); idx += 6 - a; }
            lambda = 1e-5 * m;
            ni = 2.0;
          }
          double rho = 0.0;
          int qmax = 0;
          bool lambda_bad = false;
          do {
            double x[6] = {0, 0, 0, 0, 0, 0};
            const bool ok2 = solve6(acc, acc + 21, lambda, x);
            double qt[4], tt[3];
            if (ok2) se3_oplus(x, qc, tc, qt, tt);
            else {
#pragma unroll
              for (int a = 0; a < 4; a++) qt[a] = qc[a];
#pragma unroll
              for (int a = 0; a < 3; a++) tt[a] = tc[a];
            }
            double Rt[9];
            quat_to_R(qt, Rt);
            double tchi[1] = {0.0};
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
              if (flev[i]) continue;
              const double X[3] = {fX[i * 3], fX[i * 3 + 1], fX[i * 3 + 2]};
              double pc[3], e0, e1, w;
              pose_map(Rt, tt, X, pc);
              const unsigned kb = STEREO ? fkind[i] : 0u;
              const bool st = STEREO && (kb & 1u);
              const double* Ke = STEREO ? sp.tab + 5 * (kb >> 1) : K;
              double e2 = pose_err(pc, fuv[i * 2], fuv[i * 2 + 1], Ke, e0, e1);
              if (st) { const double er = pose_err_right(pc, fur[i], Ke, Ke[4]); e2 += er * er; }
              tchi[0] += huber_rho(e2, st ? sp.delta : delta, robust, w);
            }
            block_sum<1>(tchi, red);
            evaluated = true;
#pragma unroll
            for (int a = 0; a < 4; a++) ql[a] = qt[a];
#pragma unroll
            for (int a = 0; a < 3; a++) tl[a] = tt[a];
            double tempChi = ok2 ? tchi[0] : 1.7976931348623157e308;
            double scale = 0.0;
#pragma unroll
            for (int a = 0; a < 6; a++) scale += x[a] * (lambda * x[a] + acc[21 + a]);
            rho = (currentChi - tempChi) / (scale + 1e-3);
            if (rho > 0 && isfinite(tempChi)) {
              double alpha = 1. - pow((2 * rho - 1), 3);
              alpha = fmin(alpha, 2. / 3.);
              lambda *= fmax(1. / 3., alpha);
              ni = 2;
              currentChi = tempChi;
#pragma unroll
              for (int a = 0; a < 4; a++) qc[a] = qt[a];
#pragma unroll
              for (int a = 0; a < 3; a++) tc[a] = tt[a];
            } else {
              lambda *= ni;
              ni *= 2;
              if (!isfinite(lambda)) { lambda_bad = true; qmax++; break; }
            }
            qmax++;
          } while (rho < 0 && qmax < 10);
          total_iters++;
          if (qmax == 10 || rho == 0 || lambda_bad || !isfinite(lambda)) break;
        }
      }
      // re-classification (:270-289).  Active inlier edges: chi2 cached by the last
      // computeActiveErrors (state ql); edges flagged !inlier: computeError at the current pose.
      double Rc[9], Rl[9];
      quat_to_R(qc, Rc);
      quat_to_R(evaluated ? ql : qc, Rl);
      const double* tle = evaluated ? tl : tc;
      double nout[1] = {0.0};
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double X[3] = {fX[i * 3], fX[i * 3 + 1], fX[i * 3 + 2]};
        double pc[3], e0, e1;
        const bool was_inl = finl[i] != 0;
        if (!was_inl || flev[i]) pose_map(Rc, tc, X, pc);
        else pose_map(Rl, tle, X, pc);
        const unsigned kb = STEREO ? fkind[i] : 0u;
        const bool st = STEREO && (kb & 1u);
        const double* Ke = STEREO ? sp.tab + 5 * (kb >> 1) : K;
        double e2 = pose_err(pc, fuv[i * 2], fuv[i * 2 + 1], Ke, e0, e1);
        if (st) { const double er = pose_err_right(pc, fur[i], Ke, Ke[4]); e2 += er * er; }
        const float chi2 = (float)e2;  // :277 / :297
        if ((double)chi2 > (st ? sp.chi2_thr : chi2_thr)) { finl[i] = 0; flev[i] = 1; nout[0] += 1.0; }
        else { finl[i] = 1; flev[i] = 0; }
      }
      block_sum<1>(nout, red);
      num_outlier = (int)nout[0];
      if (n < 10) break;  // :310-311
    }
    if (threadIdx.x == 0) {
      double qi[4], ti[3];
      se3_inverse(qc, tc, qi, ti);  // :315-317
      double* o = pose_out + (size_t)f * 7;
      o[0] = qi[0]; o[1] = qi[1]; o[2] = qi[2]; o[3] = qi[3];
      o[4] = ti[0]; o[5] = ti[1]; o[6] = ti[2];
      n_inlier[f] = n - num_outlier;  // :319-320
      if (lm_iters) lm_iters[f] = total_iters;
    }
    __syncthreads();
  }
}

// ur / kind / tab non-null: stereo-capable frames (tab = camera model rows, chi2_thr_s, delta_s for the stereo edges).
cudaError_t launch_pose_only(int B, const int* obs_off, const double* pose_in, const double* uv,
                             const double* Xw, const double* intr, double chi2_thr, double delta,
                             int rounds, int its_per_round, uint8_t* inlier, uint8_t* level,
                             double* pose_out, int* n_inlier, int* lm_iters, cudaStream_t stream,
                             const double* ur, const uint8_t* kind, const double* tab, double chi2_thr_s, double delta_s) {
  if (B <= 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const PoseStereo sp = {ur, kind, tab, chi2_thr_s, delta_s};
#define URMVO_POSE_LAUNCH(MINB, ST)                                                                              \
  pose_only_kernel<MINB, ST><<<B, 256, 0, stream>>>(B, obs_off, pose_in, uv, Xw, intr[0], intr[1], intr[2], intr[3], \
                                                    chi2_thr, delta, rounds, its_per_round, inlier, level, pose_out, \
                                                    n_inlier, lm_iters, sp)
  const bool stereo = ur != nullptr && kind != nullptr && tab != nullptr;
  if (B > sms) { if (stereo) URMVO_POSE_LAUNCH(2, true); else URMVO_POSE_LAUNCH(2, false); }
  else { if (stereo) URMVO_POSE_LAUNCH(1, true); else URMVO_POSE_LAUNCH(1, false); }
#undef URMVO_POSE_LAUNCH
  return cudaGetLastError();
}

}  // namespace urmvo
