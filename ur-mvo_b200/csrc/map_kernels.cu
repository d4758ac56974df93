// csrc/map_kernels.cu — gather / scatter between the device-resident map (reference
// src/mapping.cc:335-535 rebuilds the bundle-adjustment problem from shared_ptr graphs every keyframe) and the
// input / output arrays of a BA plan.  Keyframe poses (T_wc: q, p — 7 doubles), mappoint positions (3) and
// observations (uv: 2) live in slot-addressed arrays in HBM; a window is described by three slot lists.
#include "kernels.h"

namespace urmvo {

namespace {

constexpr int kMapThreads = 256;

// dst[i * W + k] = src[slot[i] * W + k]
template <int W>
__device__ __forceinline__ void gather_rows(double* __restrict__ dst, const double* __restrict__ src,
                                            const int* __restrict__ slot, int n, int gt, int gstride) {
  for (int e = gt; e < n * W; e += gstride) {
    const int i = e / W, k = e - i * W;
    dst[e] = src[(size_t)slot[i] * W + k];
  }
}

__global__ void __launch_bounds__(kMapThreads)
k_map_gather(double* pose_in, double* pts_in, double* uv, const double* __restrict__ d_kf, const double* __restrict__ d_pt,
             const double* __restrict__ d_uv, const int* __restrict__ kf_slot, const int* __restrict__ pt_slot,
             const int* __restrict__ obs_slot, int Nc, int Np, int No) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  gather_rows<7>(pose_in, d_kf, kf_slot, Nc, gt, gstride);
  gather_rows<3>(pts_in, d_pt, pt_slot, Np, gt, gstride);
  gather_rows<2>(uv, d_uv, obs_slot, No, gt, gstride);
}

// results of a window back into the map: poses of the free keyframes, all points of the window
__global__ void __launch_bounds__(kMapThreads)
k_map_scatter(double* d_kf, double* d_pt, const double* __restrict__ pose_out, const double* __restrict__ pts_out,
              const int* __restrict__ cam_free, const int* __restrict__ kf_slot, const int* __restrict__ pt_slot, int Nc, int Np) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  for (int e = gt; e < Nc * 7; e += gstride) {
    const int i = e / 7, k = e - i * 7;
    if (cam_free[i] >= 0) d_kf[(size_t)kf_slot[i] * 7 + k] = pose_out[e];
  }
  for (int e = gt; e < Np * 3; e += gstride) {
    const int i = e / 3, k = e - i * 3;
    d_pt[(size_t)pt_slot[i] * 3 + k] = pts_out[e];
  }
}

// dst[slot[i] * W + k] = vals[i * W + k]  (set / overwrite entries of the map)
__global__ void __launch_bounds__(kMapThreads)
k_map_set(double* dst, const double* __restrict__ vals, const int* __restrict__ slot, int n, int W) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  for (int e = gt; e < n * W; e += gstride) {
    const int i = e / W, k = e - i * W;
    dst[(size_t)slot[i] * W + k] = vals[e];
  }
}

// vals[i * W + k] = src[slot[i] * W + k]
__global__ void __launch_bounds__(kMapThreads)
k_map_get(double* vals, const double* __restrict__ src, const int* __restrict__ slot, int n, int W) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  for (int e = gt; e < n * W; e += gstride) {
    const int i = e / W, k = e - i * W;
    vals[e] = src[(size_t)slot[i] * W + k];
  }
}

inline int map_grid(long long work) {
  const long long g = (work + kMapThreads - 1) / kMapThreads;
  return (int)(g < 1 ? 1 : (g > 592 ? 592 : g));
}

}  // namespace

cudaError_t launch_map_gather(double* pose_in, double* pts_in, double* uv, const double* d_kf, const double* d_pt,
                              const double* d_uv, const int* kf_slot, const int* pt_slot, const int* obs_slot, int Nc,
                              int Np, int No, cudaStream_t s) {
  k_map_gather<<<map_grid((long long)No * 2 + Np * 3 + Nc * 7), kMapThreads, 0, s>>>(pose_in, pts_in, uv, d_kf, d_pt, d_uv, kf_slot,
                                                                                     pt_slot, obs_slot, Nc, Np, No);
  return cudaGetLastError();
}
cudaError_t launch_map_scatter(double* d_kf, double* d_pt, const double* pose_out, const double* pts_out, const int* cam_free,
                               const int* kf_slot, const int* pt_slot, int Nc, int Np, cudaStream_t s) {
  k_map_scatter<<<map_grid((long long)Np * 3 + Nc * 7), kMapThreads, 0, s>>>(d_kf, d_pt, pose_out, pts_out, cam_free, kf_slot, pt_slot, Nc, Np);
  return cudaGetLastError();
}
cudaError_t launch_map_set(double* dst, const double* vals, const int* slot, int n, int W, cudaStream_t s) {
  k_map_set<<<map_grid((long long)n * W), kMapThreads, 0, s>>>(dst, vals, slot, n, W);
  return cudaGetLastError();
}
cudaError_t launch_map_get(double* vals, const double* src, const int* slot, int n, int W, cudaStream_t s) {
  k_map_get<<<map_grid((long long)n * W), kMapThreads, 0, s>>>(vals, src, slot, n, W);
  return cudaGetLastError();
}

}  // namespace urmvo
