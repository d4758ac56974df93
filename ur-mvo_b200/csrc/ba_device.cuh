// csrc/ba_device.cuh — device helpers shared by the BA translation units (ba_kernels.cu: persistent
// small / medium windows; ba_large.cu: phase kernels of one large problem): per-edge math of
// g2o EdgeSE3ProjectXYZ and the packed-group staging layout.
#pragma once
#include "ba_types.h"
#include "common.cuh"

namespace urmvo {

// ------------------------------------------------------------------------------- edge math

struct EdgeLin {
  double e0, e1, w, rho0;
  double Jx[6];   // 2x3 d e / d X
  double Jp[12];  // 2x6 d e / d xi (rotation first)
};

// pc = R X + t
__device__ __forceinline__ void map_point(const double* __restrict__ Rt, const double* X, double* pc) {
  pc[0] = Rt[0] * X[0] + Rt[1] * X[1] + Rt[2] * X[2] + Rt[9];
  pc[1] = Rt[3] * X[0] + Rt[4] * X[1] + Rt[5] * X[2] + Rt[10];
  pc[2] = Rt[6] * X[0] + Rt[7] * X[1] + Rt[8] * X[2] + Rt[11];
}

// fp64 division costs ~10x a multiply on the SM (software Newton iteration), so every edge takes ONE
// reciprocal iz = 1/z and forms x/z, y/z, x/z^2 ... by multiplication.  The results differ from the
// literal g2o expressions (x*y/z2*fx ...) by a few ulp, far inside the 1e-6 parity tolerance.

// EdgeSE3ProjectXYZ::computeError: e = z - (x/z*fx + cx, y/z*fy + cy); returns chi2. pz[0..2] = x/z, y/z, 1/z.
__device__ __forceinline__ double edge_error(const double* pc, double u, double v, const double* K,
                                             double& e0, double& e1, double* pz) {
  const double iz = 1.0 / pc[2];
  pz[0] = pc[0] * iz; pz[1] = pc[1] * iz; pz[2] = iz;
  e0 = u - (pz[0] * K[0] + K[2]);
  e1 = v - (pz[1] * K[1] + K[3]);
  return e0 * e0 + e1 * e1;
}
__device__ __forceinline__ double edge_error(const double* pc, double u, double v, const double* K,
                                             double& e0, double& e1) {
  double pz[3];
  return edge_error(pc, u, v, K, e0, e1, pz);
}

// EdgeSE3ProjectXYZ::linearizeOplus pose part (2x6), from pz = (x/z, y/z, 1/z).
__device__ __forceinline__ void edge_jac_pose(const double* pz, const double* K, double* Jp) {
  const double xz = pz[0], yz = pz[1], iz = pz[2];
  Jp[0] = xz * yz * K[0];
  Jp[1] = -(1 + xz * xz) * K[0];
  Jp[2] = yz * K[0];
  Jp[3] = -iz * K[0];
  Jp[4] = 0;
  Jp[5] = xz * iz * K[0];
  Jp[6] = (1 + yz * yz) * K[1];
  Jp[7] = -xz * yz * K[1];
  Jp[8] = -xz * K[1];
  Jp[9] = 0;
  Jp[10] = -iz * K[1];
  Jp[11] = yz * iz * K[1];
}

// EdgeSE3ProjectXYZ::linearizeOplus point part (2x3) = -1/z * [[fx,0,-x/z fx],[0,fy,-y/z fy]] * R
__device__ __forceinline__ void edge_jac_point(const double* __restrict__ Rt, const double* pz,
                                               const double* K, double* Jx) {
  const double t02 = -pz[0] * K[0], t12 = -pz[1] * K[1];
  const double miz = -pz[2];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    Jx[c] = miz * (K[0] * Rt[c] + t02 * Rt[6 + c]);
    Jx[3 + c] = miz * (K[1] * Rt[3 + c] + t12 * Rt[6 + c]);
  }
}

// EdgeStereoSE3ProjectXYZ (upstream g2o types_six_dof_expmap.cpp; reference src/g2o_optimization.cc:96-118):
// third residual row u_right - (u_left_projected - bf / z) and the third Jacobian rows, from
// pz = (x/z, y/z, 1/z) and the first rows (row 2 = row 0 plus the bf terms).
__device__ __forceinline__ double edge_error_right(const double* pz, double ur, const double* K, double bf) {
  return ur - (pz[0] * K[0] + K[2] - bf * pz[2]);
}
__device__ __forceinline__ void edge_jac_pose_right(const double* pz, const double* Jp, double bf, double* J2) {
  const double xz = pz[0], yz = pz[1], iz = pz[2];
  J2[0] = Jp[0] - bf * yz * iz;
  J2[1] = Jp[1] + bf * xz * iz;
  J2[2] = Jp[2];
  J2[3] = Jp[3];
  J2[4] = 0;
  J2[5] = Jp[5] - bf * iz * iz;
}
__device__ __forceinline__ void edge_jac_point_right(const double* __restrict__ Rt, const double* pz, const double* Jx,
                                                     double bf, double* J2) {
  const double s = bf * pz[2] * pz[2];
#pragma unroll
  for (int c = 0; c < 3; c++) J2[c] = Jx[c] - s * Rt[6 + c];
}

// Symmetric 3x3 inverse, packed (00 01 02 11 12 22), cofactor formula like Eigen's fixed-size inverse.
__device__ __forceinline__ void sym3_inverse(const double* h, double* r) {
  const double a00 = h[0], a01 = h[1], a02 = h[2], a11 = h[3], a12 = h[4], a22 = h[5];
  const double c00 = a11 * a22 - a12 * a12;
  const double c01 = a12 * a02 - a01 * a22;
  const double c02 = a01 * a12 - a11 * a02;
  const double det = a00 * c00 + a01 * c01 + a02 * c02;
  const double id = 1.0 / det;
  r[0] = c00 * id;
  r[1] = c01 * id;
  r[2] = c02 * id;
  r[3] = (a00 * a22 - a02 * a02) * id;
  r[4] = (a01 * a02 - a00 * a12) * id;
  r[5] = (a00 * a11 - a01 * a01) * id;
}

// ------------------------------------------------------------------------------- packed staging

constexpr int kPackSlots = 33;     // 32 observation slots + one all-zero slot (index kPackZero)
constexpr int kPackZero = 32;      // slot referenced by an absent (point, camera) entry: contributes exactly 0
constexpr int kPackCam = 16;     // width of the (point-in-group, free camera) -> slot table
constexpr int kPackFields = 38;  // Jp[12] | B[6] | A[6] | we[2] | w | h[6] | bl[3] | g[2]

struct PackStage {
  double* f;          // kPackFields x 32
  signed char* slot;  // 32 x kPackCam
  __device__ __forceinline__ double& Jp(int a, int s) { return f[a * kPackSlots + s]; }
  __device__ __forceinline__ double& B(int a, int s) { return f[(12 + a) * kPackSlots + s]; }
  __device__ __forceinline__ double& A(int a, int s) { return f[(18 + a) * kPackSlots + s]; }
  __device__ __forceinline__ double& we(int a, int s) { return f[(24 + a) * kPackSlots + s]; }
  __device__ __forceinline__ double& w(int s) { return f[26 * kPackSlots + s]; }
  __device__ __forceinline__ double& h(int a, int s) { return f[(27 + a) * kPackSlots + s]; }
  __device__ __forceinline__ double& bl(int a, int s) { return f[(33 + a) * kPackSlots + s]; }
  __device__ __forceinline__ double& g(int a, int s) { return f[(36 + a) * kPackSlots + s]; }
  __device__ __forceinline__ double* recbuf() { return f + kPackFields * kPackSlots; }  // 32 x 32 bytes
};

// Packed-mode record access.  The record of the NEXT group is copied global -> shared with cp.async
// (LDGSTS: no destination registers, so the compiler cannot consume it early) into the warp's record buffer
// and read back at the top of the next iteration.  The buffer holds the two 16-byte halves of the 32 records as
// two planes ([half][lane]): every copy and every read-back touches 512 contiguous bytes (a lane stride of 32
// bytes made four lanes share each bank group: twice the shared-memory wavefronts, measured as bank conflicts
// of the LDGSTS path).
struct RecRegs { int4 a, b; };
__device__ __forceinline__ void rec_prefetch(double* buf, const ObsRec* rec, int g, int lane) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(buf) + lane * 16;
  const size_t src = __cvta_generic_to_global(rec + (size_t)g * 32 + lane);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 512), "l"(src + 16) : "memory");
}
__device__ __forceinline__ RecRegs rec_take(const double* buf, int lane) {
  asm volatile("cp.async.wait_all;" ::: "memory");
  const int4* r = reinterpret_cast<const int4*>(buf);
  RecRegs x;
  x.a = r[lane];
  x.b = r[32 + lane];
  return x;
}

}  // namespace urmvo
