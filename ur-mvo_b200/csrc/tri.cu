// csrc/tri.cu — batched multi-view midpoint triangulation on sm_100a and its
// C-ABI entry point.
//
// Reference behaviour reproduced: Mapping::TriangulateMappoint, src/mapping.cc:151-205 (one mappoint
// per call in the reference; here every new mappoint of a keyframe in one launch):
//   bearing b_k = R_k ((u-cx)/fx, (v-cy)/fy, 1)   (Camera::BackProjectMono, src/camera.cc:168-174)
//   A = N I - sum_k b_k b_k^T / |b_k|^2,  rhs = sum_k p_k - sum_k b_k (b_k . p_k) / |b_k|^2
//   Eigen::ColPivHouseholderQR<Matrix3d> with setThreshold(1e-5): rank < 3 -> false, else solve.
// One thread per mappoint: the observers are summed in their stored order (std::map order of the
// reference = ascending frame id), the 3x3 QR with column pivoting follows Eigen's Householder
// conventions operation by operation (restated in the CPU checker tri_oracle.cpp).  A mappoint has
// <= ~35 observers, so the kernel is bound by the 40 B/observer it reads (pose index, keypoint) plus
// the L2-resident pose table; there is nothing to tile.
#include <cstring>
#include <vector>

#include "capi_internal.h"

namespace urmvo {
namespace {

__device__ int colpiv_qr_solve3(double A[3][3], double b[3], double threshold, double* x) {
  int perm[3] = {0, 1, 2};
  double maxpivot = 0.0, diag[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; k++) {
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
    double tail = 0.0;
    for (int i = k + 1; i < 3; i++) tail += A[i][k] * A[i][k];
    const double c0 = A[k][k];
    double beta, tau, v[3] = {0, 0, 0};
    if (tail == 0.0) {
      beta = c0; tau = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail);
      if (c0 >= 0.0) beta = -beta;
      for (int i = k + 1; i < 3; i++) v[i] = A[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    v[k] = 1.0;
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
    diag[k] = fabs(beta);
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

__global__ void __launch_bounds__(128)
tri_midpoint_kernel(int n_pts, const int* __restrict__ obs_off, const int* __restrict__ obs_pose,
                    const double2* __restrict__ obs_uv, const double* __restrict__ Rp, double fx_inv,
                    double fy_inv, double cx, double cy, double* __restrict__ X_out,
                    uint8_t* __restrict__ ok_out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_pts) return;
  const int o0 = obs_off[l], n = obs_off[l + 1] - o0;
  double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, rhs[3] = {0, 0, 0};
  for (int k = 0; k < n; k++) {
    const double* R = Rp + (size_t)obs_pose[o0 + k] * 12;
    const double2 uv = obs_uv[o0 + k];
    const double bx = (uv.x - cx) * fx_inv, by = (uv.y - cy) * fy_inv;
    const double b[3] = {R[0] * bx + R[1] * by + R[2], R[3] * bx + R[4] * by + R[5], R[6] * bx + R[7] * by + R[8]};
    const double inv = 1.0 / (b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    const double bp = b[0] * R[9] + b[1] * R[10] + b[2] * R[11];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = 0; j < 3; j++) A[i][j] -= b[i] * inv * b[j];
      rhs[i] += R[9 + i] - b[i] * inv * bp;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++) A[i][i] += (double)n;
  double X[3] = {0, 0, 0};
  const bool ok = n >= 2 && colpiv_qr_solve3(A, rhs, 1e-5, X) == 3;
  ok_out[l] = ok ? 1 : 0;
  if (ok) { X_out[l * 3] = X[0]; X_out[l * 3 + 1] = X[1]; X_out[l * 3 + 2] = X[2]; }
}

}  // namespace
}  // namespace urmvo

using namespace urmvo;

extern "C" int urmvo_triangulate_batch(urmvo_ctx* ctx, int n_pts, const int32_t* obs_off, const int32_t* obs_pose,
                                       const double* obs_uv, int n_poses, const double* poses_Rp, const double* intr,
                                       double* pts, uint8_t* ok) {
  if (!ctx || n_pts < 0 || !obs_off || !intr || !pts || !ok) return set_error(URMVO_ERR_ARG, "triangulate_batch: null input");
  if (n_pts == 0) return URMVO_OK;
  const int n_obs = obs_off[n_pts];
  if (obs_off[0] != 0 || n_obs < 0 || (n_obs > 0 && (!obs_pose || !obs_uv || !poses_Rp || n_poses <= 0)))
    return set_error(URMVO_ERR_ARG, "triangulate_batch: bad observation arrays");
  for (int l = 0; l < n_pts; l++)
    if (obs_off[l + 1] < obs_off[l]) return set_error(URMVO_ERR_ARG, "triangulate_batch: offsets must ascend");
  for (int o = 0; o < n_obs; o++)
    if (obs_pose[o] < 0 || obs_pose[o] >= n_poses) return set_error(URMVO_ERR_ARG, "triangulate_batch: pose index out of range");
  CU_TRY(cudaSetDevice(ctx->device));
  // layout of the borrowed workspace: [off | pose idx | uv | Rp | X | ok]
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t b_off = up((size_t)(n_pts + 1) * 4), b_idx = up((size_t)n_obs * 4), b_uv = up((size_t)n_obs * 16);
  const size_t b_rp = up((size_t)n_poses * 96), b_x = up((size_t)n_pts * 24), b_ok = up((size_t)n_pts);
  const size_t total = b_off + b_idx + b_uv + b_rp + b_x + b_ok;
  unsigned char* dev = nullptr;
  bool borrowed = false;
  if (!ctx->ws_in_use) {
    if (ctx->ws_bytes < total) {
      if (ctx->ws_dev) cudaFree(ctx->ws_dev);
      ctx->ws_dev = nullptr; ctx->ws_bytes = 0;
      CU_TRY(cudaMalloc(&ctx->ws_dev, total + total / 4));
      ctx->ws_bytes = total + total / 4;
    }
    dev = ctx->ws_dev; borrowed = true; ctx->ws_in_use = true;
  } else {
    CU_TRY(cudaMalloc(&dev, total));
  }
  auto release = [&] { if (borrowed) ctx->ws_in_use = false; else cudaFree(dev); };
  if (ctx->ensure_pinned(total)) { release(); return set_error(URMVO_ERR_CUDA, "cudaMallocHost failed"); }
  unsigned char* H = (unsigned char*)ctx->pinned;
  unsigned char *h_off = H, *h_idx = h_off + b_off, *h_uv = h_idx + b_idx, *h_rp = h_uv + b_uv, *h_x = h_rp + b_rp, *h_ok = h_x + b_x;
  std::memcpy(h_off, obs_off, (size_t)(n_pts + 1) * 4);
  if (n_obs) { std::memcpy(h_idx, obs_pose, (size_t)n_obs * 4); std::memcpy(h_uv, obs_uv, (size_t)n_obs * 16); std::memcpy(h_rp, poses_Rp, (size_t)n_poses * 96); }
  cudaStream_t s = ctx->stream;
  cudaError_t e = cudaMemcpyAsync(dev, H, b_off + b_idx + b_uv + b_rp, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    unsigned char *d_off = dev, *d_idx = d_off + b_off, *d_uv = d_idx + b_idx, *d_rp = d_uv + b_uv, *d_x = d_rp + b_rp, *d_ok = d_x + b_x;
    tri_midpoint_kernel<<<(n_pts + 127) / 128, 128, 0, s>>>(n_pts, (const int*)d_off, (const int*)d_idx, (const double2*)d_uv,
                                                            (const double*)d_rp, 1.0 / intr[0], 1.0 / intr[1], intr[2], intr[3],
                                                            (double*)d_x, d_ok);
    e = cudaGetLastError();
    ctx->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_x, d_x, b_x + b_ok, cudaMemcpyDeviceToHost, s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  release();
  if (e != cudaSuccess) return set_error(URMVO_ERR_CUDA, std::string("triangulate_batch: ") + cudaGetErrorString(e));
  std::memcpy(ok, h_ok, (size_t)n_pts);
  for (int l = 0; l < n_pts; l++)  // like the reference, a failed mappoint keeps its position
    if (ok[l]) std::memcpy(pts + 3 * (size_t)l, h_x + 24 * (size_t)l, 24);
  return URMVO_OK;
}
