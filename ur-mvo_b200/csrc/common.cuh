// csrc/common.cuh — scopes, deterministic reductions and fp64 SE3 math shared by the BA kernels.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg = cooperative_groups;

namespace urmvo {

// ---------------------------------------------------------------------------------------------
// Execution scopes.  The LM solver is written once against a Scope: the set of CTAs that cooperate
// on one problem.  Cross-CTA data always goes through global memory (L2); sync() makes it visible.
//   CtaScope     one CTA            (pose-only frames, single-CTA PCG)
//   ClusterScope one thread-block cluster per BA window (batched windows, hardware cluster barrier)
//   GridScope    the whole cooperative grid on one large problem
// ---------------------------------------------------------------------------------------------
struct CtaScope {
  __device__ __forceinline__ int nblk() const { return 1; }
  __device__ __forceinline__ int blk() const { return 0; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

struct ClusterScope {
  int n, r;
  __device__ ClusterScope() {
    cg::cluster_group c = cg::this_cluster();
    n = (int)c.num_blocks();
    r = (int)c.block_rank();
  }
  __device__ __forceinline__ int nblk() const { return n; }
  __device__ __forceinline__ int blk() const { return r; }
  __device__ __forceinline__ void sync() const {
    __threadfence();
    cg::this_cluster().sync();
  }
};

struct GridScope {
  __device__ __forceinline__ int nblk() const { return (int)gridDim.x; }
  __device__ __forceinline__ int blk() const { return (int)blockIdx.x; }
  __device__ __forceinline__ void sync() const {
    __threadfence();
    cg::this_grid().sync();
  }
};

// ---------------------------------------------------------------------------------------------
// Deterministic reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// CTA-wide sum of NV values per thread; result valid in every thread. red: >= NV*32 doubles of smem.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) red[k * 32 + wid] = v[k];
  }
  __syncthreads();
  if (NV <= 4) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double s = 0.0;
      for (int w = 0; w < nw; w++) s += red[k * 32 + w];
      v[k] = s;
    }
  } else {
    // many values: thread k sums the warp partials of value k once (same warp order), everybody
    // reads the NV totals back — instead of NV * nw dependent additions in every thread
    if (threadIdx.x < NV) {
      double s = 0.0;
      for (int w = 0; w < nw; w++) s += red[threadIdx.x * 32 + w];
      red[NV * 32 + threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = red[NV * 32 + k];
  }
}

constexpr int kScopePartWidth = 8;  // doubles per CTA in each of the two scratch buffers

// Scope-wide reduction of NV sums and NM maxima.  part: 2 * nblk * (NV+NM) doubles of global
// memory (double-buffered by `parity`, which the caller flips after every call).  Fixed order:
// per-CTA tree, then lane-strided partials + butterfly, so the result is run-to-run reproducible.
template <int NV, int NM, class Scope>
__device__ __forceinline__ void scope_reduce(const Scope& sc, double (&sum)[NV], double (&mx)[NM == 0 ? 1 : NM],
                                             double* part, int& parity, double* red) {
  constexpr int NT = NV + NM;
  static_assert(NT <= kScopePartWidth, "scope_reduce: too many values");
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) sum[k] = warp_sum(sum[k]);
#pragma unroll
  for (int k = 0; k < NM; k++) mx[k] = warp_max(mx[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) red[k * 32 + wid] = sum[k];
#pragma unroll
    for (int k = 0; k < NM; k++) red[(NV + k) * 32 + wid] = mx[k];
  }
  __syncthreads();
  const int nb = sc.nblk();
  double* buf = part + (size_t)parity * nb * kScopePartWidth;
  if ((int)threadIdx.x < NT) {
    const int k = threadIdx.x;
    double s = red[k * 32];
    if (k < NV) { for (int w = 1; w < nw; w++) s += red[k * 32 + w]; }
    else        { for (int w = 1; w < nw; w++) s = fmax(s, red[k * 32 + w]); }
    if (nb == 1) red[NT * 32 + k] = s;
    else __stcg(&buf[(size_t)sc.blk() * NT + k], s);
  }
  if (nb > 1) {
    sc.sync();
    if (wid == 0) {
      // lane-strided partial sums; all loads of one stride step are independent (one L2 round trip)
      double acc[NT];
#pragma unroll
      for (int k = 0; k < NT; k++) acc[k] = (k < NV) ? 0.0 : -1.0e300;
#pragma unroll 2
      for (int b = lane; b < nb; b += 32) {
        double x[NT];
#pragma unroll
        for (int k = 0; k < NT; k++) x[k] = __ldcg(&buf[(size_t)b * NT + k]);
#pragma unroll
        for (int k = 0; k < NT; k++) acc[k] = (k < NV) ? acc[k] + x[k] : fmax(acc[k], x[k]);
      }
#pragma unroll
      for (int k = 0; k < NT; k++) {
        const double s = (k < NV) ? warp_sum(acc[k]) : warp_max(acc[k]);
        if (lane == 0) red[NT * 32 + k] = s;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; k++) sum[k] = red[NT * 32 + k];
#pragma unroll
  for (int k = 0; k < NM; k++) mx[k] = red[NT * 32 + NV + k];
  parity ^= 1;
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// fp64 SE3 math with g2o's conventions (g2o/types/slam3d/se3quat.h, upstream).
// Quaternions are (x,y,z,w).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void quat_normalize_w(double* q) {
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
}

__device__ __forceinline__ void quat_to_R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

__device__ __forceinline__ void R_to_quat(const double* m, double* q) {
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0);
    double qq[3];
    qq[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
    qq[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    qq[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2];
  }
}

// T (q[4], t[3]) <- inverse(T)
__device__ __forceinline__ void se3_inverse(const double* qin, const double* tin, double* q, double* t) {
  q[0] = -qin[0]; q[1] = -qin[1]; q[2] = -qin[2]; q[3] = qin[3];
  double R[9];
  quat_to_R(q, R);
  const double a = -tin[0], b = -tin[1], c = -tin[2];
  t[0] = R[0] * a + R[1] * b + R[2] * c;
  t[1] = R[3] * a + R[4] * b + R[5] * c;
  t[2] = R[6] * a + R[7] * b + R[8] * c;
}

// (q,t) <- exp(u) * (q,t), u = (omega, upsilon): VertexSE3Expmap::oplusImpl.
__device__ __forceinline__ void se3_oplus(const double* u, const double* qin, const double* tin,
                                          double* qout, double* tout) {
  const double ox = u[0], oy = u[1], oz = u[2];
  const double theta = sqrt(ox * ox + oy * oy + oz * oz);
  const double O[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
  double O2[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
  double a, b, c, d;
  if (theta < 0.00001) {
    a = 1.0; b = 0.5; c = 0.5; d = 1.0 / 6.0;
  } else {
    const double st = sin(theta), ct = cos(theta);
    a = st / theta;
    b = (1 - ct) / (theta * theta);
    c = b;
    d = (theta - st) / (theta * theta * theta);
  }
  double Re[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    Re[i] = I + a * O[i] + b * O2[i];
    V[i] = I + c * O[i] + d * O2[i];
  }
  double qe[4], te[3];
  R_to_quat(Re, qe);
  quat_normalize_w(qe);
#pragma unroll
  for (int i = 0; i < 3; i++) te[i] = V[i * 3] * u[3] + V[i * 3 + 1] * u[4] + V[i * 3 + 2] * u[5];
  // exp(u) * T
  const double ax = qe[0], ay = qe[1], az = qe[2], aw = qe[3];
  const double bx = qin[0], by = qin[1], bz = qin[2], bw = qin[3];
  double qr[4];
  qr[3] = aw * bw - ax * bx - ay * by - az * bz;
  qr[0] = aw * bx + ax * bw + ay * bz - az * by;
  qr[1] = aw * by + ay * bw + az * bx - ax * bz;
  qr[2] = aw * bz + az * bw + ax * by - ay * bx;
  double Rq[9];
  quat_to_R(qe, Rq);
#pragma unroll
  for (int i = 0; i < 3; i++)
    tout[i] = te[i] + (Rq[i * 3] * tin[0] + Rq[i * 3 + 1] * tin[1] + Rq[i * 3 + 2] * tin[2]);
  quat_normalize_w(qr);
  qout[0] = qr[0]; qout[1] = qr[1]; qout[2] = qr[2]; qout[3] = qr[3];
}

// g2o RobustKernelHuber::robustify: returns rho0, writes rho1.
__device__ __forceinline__ double huber_rho(double e2, double delta, bool robust, double& w) {
  const double dsqr = delta * delta;
  if (!robust || e2 <= dsqr) { w = 1.0; return e2; }
  const double sqrte = sqrt(e2);
  w = delta / sqrte;
  return 2 * sqrte * delta - dsqr;
}

}  // namespace urmvo
