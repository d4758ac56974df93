// csrc/ba_large_tail.cuh — the end of a reduced-camera-system solve in tile mode, shared by the
// sequential band solve (ba_large.cu) and the block cyclic reduction (ba_bcr.cu): x_p, the pose part of
// g2o's computeScale  sum x (lambda x + b_p), and the trial cameras exp(x_c) * T_c
// (OptimizationAlgorithmLevenberg::solve -> SparseOptimizer::update).
#pragma once
#include "ba_device.cuh"

namespace urmvo {

// Pose part of computeScale and x_p.  yv: the solution (Ncf*6, shared or global memory), redv: >= blockDim.x/32
// doubles of shared scratch.  One CTA.
__device__ __forceinline__ void lg_tail_scale(const BAWin& W, LgState* stt, const double* yv, double* redv, double lambda) {
  const int t = threadIdx.x, n6 = W.Ncf * 6;
  double sc = 0.0;
  for (int i = t; i < n6; i += blockDim.x) {
    const double x = yv[i];
    W.xp[i] = x;
    sc += x * (lambda * x + W.bp[i]);
  }
  sc = warp_sum(sc);
  if ((t & 31) == 0) redv[t >> 5] = sc;
  __syncthreads();
  if (t == 0) {
    double s2 = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s2 += redv[w];
    stt->scale_pose = s2;
    stt->ok2 = 1;
  }
}

// Trial cameras exp(x_c) * T_c for the cameras first, first + stride, ...
__device__ __forceinline__ void lg_tail_cameras(const BAWin& W, const LgState* stt, const double* yv, int first, int stride) {
  const int cur = stt->cur, tr = cur ^ 1;
  for (int cc = first; cc < W.Nc; cc += stride) {
    const double* q = W.cam[cur] + (size_t)cc * 7;
    double* qo = W.cam[tr] + (size_t)cc * 7;
    const int cf = W.cam_free[cc];
    if (cf >= 0) {
      double u[6];
#pragma unroll
      for (int e = 0; e < 6; e++) u[e] = yv[cf * 6 + e];
      double qn[4], tn[3];
      se3_oplus(u, q, q + 4, qn, tn);
      qo[0] = qn[0]; qo[1] = qn[1]; qo[2] = qn[2]; qo[3] = qn[3];
      qo[4] = tn[0]; qo[5] = tn[1]; qo[6] = tn[2];
    } else {
#pragma unroll
      for (int e = 0; e < 7; e++) qo[e] = q[e];
    }
    double R[9];
    quat_to_R(qo, R);
    double* o = W.camRt[tr] + (size_t)cc * 12;
#pragma unroll
    for (int e = 0; e < 9; e++) o[e] = R[e];
    o[9] = qo[4]; o[10] = qo[5]; o[11] = qo[6];
  }
}

__device__ __forceinline__ void lg_solve_tail(const BAWin& W, LgState* stt, const double* yv, double* redv,
                                              double lambda) {
  lg_tail_scale(W, stt, yv, redv, lambda);
  lg_tail_cameras(W, stt, yv, threadIdx.x, blockDim.x);
}

}  // namespace urmvo
