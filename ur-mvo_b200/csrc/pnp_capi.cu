// csrc/pnp_capi.cu — C ABI of SolvePnPWithCV (include/urmvo_b200.h; reference src/g2o_optimization.cc:323-377,
// cv::solvePnPRansac(..., false, 100, 20.0, 0.99, inliers)).
//
// Host work (what is inherently sequential in OpenCV's RANSACPointSetRegistrator::run):
//   * the 5-index subsets drawn with cv::RNG(-1) (the PnP callback has no checkSubset, so the sequence depends
//     on the number of points only) — all max_iters subsets of every frame are drawn up front;
//   * the replay of "goodCount > max(maxGoodCount, 4) -> new best, niters = RANSACUpdateNumIters(...)" over the
//     per-iteration inlier counts the device returns: the result is the one OpenCV's sequential loop produces,
//     although every hypothesis of the budget has been evaluated.
// EPnP, the reprojection test, the inlier masks and the refinement over the inliers run in pnp_kernels.cu.
// No CPU fallback.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "capi_internal.h"
#include "kernels.h"

using namespace urmvo;

namespace {

struct CvRng {
  uint64_t state = 0xffffffffffffffffull;
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a + (int)(next() % (unsigned)(b - a)); }
};

int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = std::min(std::max(p, 0.), 1.);
  ep = std::min(std::max(ep, 0.), 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::nearbyint(num / denom);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

extern "C" int urmvo_pnp_ransac_batch(urmvo_ctx* ctx, int B, const int32_t* off, const float* obj, const float* img,
                                      const double* intr, int max_iters, double reproj_thr, double confidence,
                                      uint8_t* inlier, urmvo_pnp_stats* stats) {
  try {
    if (!ctx || !off || !obj || !img || !intr || !inlier || !stats) return set_error(URMVO_ERR_ARG, "pnp_ransac: null argument");
    if (B <= 0 || B > 65535) return set_error(URMVO_ERR_ARG, "pnp_ransac: 1 <= B <= 65535 required");
    if (max_iters <= 0) max_iters = 100;
    if (max_iters > 1000) return set_error(URMVO_ERR_ARG, "pnp_ransac: max_iters <= 1000");
    if (!(reproj_thr > 0)) reproj_thr = 20.0;
    if (!(confidence > 0 && confidence < 1)) confidence = 0.99;
    if (!(intr[0] != 0 && intr[1] != 0)) return set_error(URMVO_ERR_ARG, "pnp_ransac: zero focal length");
    int max_n = 0;
    for (int b = 0; b < B; b++) {
      const int n = off[b + 1] - off[b];
      if (off[b] < 0 || n < 6)
        return set_error(URMVO_ERR_UNSUPPORTED, "pnp_ransac: every problem needs >= 6 points (4 / 5 points are other OpenCV "
                                                "branches; the reference returns 0 below 8 points without calling OpenCV)");
      max_n = std::max(max_n, n);
    }
    const size_t T = (size_t)off[B];
    const int H = B * max_iters, words = (max_n + 31) / 32;
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    // device layout
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t d_off = take((B + 1) * sizeof(int)), d_sets = take((size_t)H * 5 * sizeof(int));
    const size_t d_obj = take(T * 3 * sizeof(float)), d_img = take(T * 2 * sizeof(float)), d_K = take(4 * sizeof(double));
    const size_t in_bytes = o;  // everything up to here is uploaded in one copy
    const size_t d_models = take((size_t)H * 12 * sizeof(double)), d_valid = take((size_t)H * sizeof(int));
    const size_t d_counts = take((size_t)H * sizeof(int)), d_masks = take((size_t)H * words * sizeof(unsigned));
    const size_t d_best = take(B * sizeof(int)), d_out = take((size_t)B * 12 * sizeof(double)), d_inl = take(T);
    const size_t dev_bytes = o;
    unsigned char* D = nullptr;
    bool borrowed = false;
    if (!ctx->ws_in_use) {
      if (ctx->ws_bytes < dev_bytes) {
        if (ctx->ws_dev) cudaFree(ctx->ws_dev);
        ctx->ws_dev = nullptr; ctx->ws_bytes = 0;
        const size_t want = std::max(dev_bytes, (size_t)1 << 22);
        if (cudaMalloc(&ctx->ws_dev, want) != cudaSuccess) return set_error(URMVO_ERR_CUDA, "pnp_ransac: cudaMalloc failed");
        ctx->ws_bytes = want;
      }
      D = ctx->ws_dev; borrowed = true; ctx->ws_in_use = true;
    } else if (cudaMalloc(&D, dev_bytes) != cudaSuccess) {
      return set_error(URMVO_ERR_CUDA, "pnp_ransac: cudaMalloc failed");
    }
    struct Release {
      urmvo_ctx* c; unsigned char* d; bool b;
      ~Release() { if (b) c->ws_in_use = false; else if (d) cudaFree(d); }
    } release{ctx, D, borrowed};
    const size_t pin_bytes = in_bytes + align_up((size_t)H * sizeof(int), 256) + align_up(B * sizeof(int), 256) +
                             align_up((size_t)B * 12 * sizeof(double), 256) + T;
    if (ctx->ensure_pinned(pin_bytes)) return set_error(URMVO_ERR_CUDA, "pnp_ransac: cudaMallocHost failed");
    unsigned char* P = (unsigned char*)ctx->pinned;
    // ---- inputs + subsets
    std::memcpy(P + d_off, off, (B + 1) * sizeof(int));
    int* sets = (int*)(P + d_sets);
    for (int b = 0; b < B; b++) {
      CvRng rng;
      const int n = off[b + 1] - off[b];
      for (int it = 0; it < max_iters; it++) {
        int* idx = sets + ((size_t)b * max_iters + it) * 5;
        for (int i = 0; i < 5; i++) {
          int v = rng.uniform(0, n);
          for (;;) {
            bool dup = false;
            for (int j = 0; j < i; j++) dup |= idx[j] == v;
            if (!dup) break;
            v = rng.uniform(0, n);
          }
          idx[i] = v;
        }
      }
    }
    std::memcpy(P + d_obj, obj, T * 3 * sizeof(float));
    std::memcpy(P + d_img, img, T * 2 * sizeof(float));
    std::memcpy(P + d_K, intr, 4 * sizeof(double));
    CU_TRY(cudaMemcpyAsync(D, P, in_bytes, cudaMemcpyHostToDevice, s));
    const float thr2 = (float)(reproj_thr * reproj_thr);
    CU_TRY(launch_pnp_hypotheses(H, max_iters, words, (const int*)(D + d_off), (const int*)(D + d_sets), (const float*)(D + d_obj),
                                 (const float*)(D + d_img), (const double*)(D + d_K), (double*)(D + d_models), (int*)(D + d_valid),
                                 thr2, (unsigned*)(D + d_masks), (int*)(D + d_counts), s));
    ctx->launches += 2;
    int* h_counts = (int*)(P + in_bytes);
    int* h_best = (int*)((unsigned char*)h_counts + align_up((size_t)H * sizeof(int), 256));
    double* h_out = (double*)((unsigned char*)h_best + align_up(B * sizeof(int), 256));
    uint8_t* h_inl = (uint8_t*)((unsigned char*)h_out + align_up((size_t)B * 12 * sizeof(double), 256));
    CU_TRY(cudaMemcpyAsync(h_counts, D + d_counts, (size_t)H * sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    // ---- OpenCV's sequential bookkeeping over the counts
    for (int b = 0; b < B; b++) {
      const int n = off[b + 1] - off[b];
      int niters = max_iters, max_good = 0, best = -1, models = 0, iter = 0;
      for (iter = 0; iter < niters; iter++) {
        const int good = h_counts[(size_t)b * max_iters + iter];
        if (good < 0) continue;  // runKernel returned no model
        models++;
        if (good > std::max(max_good, 4)) {
          max_good = good;
          best = iter;
          niters = ransac_update_num_iters(confidence, (double)(n - good) / n, 5, niters);
        }
      }
      h_best[b] = best;
      stats[b].found = best >= 0 ? 1 : 0;
      stats[b].iters = iter;
      stats[b].n_inliers = max_good;
      stats[b].n_models = models;
    }
    CU_TRY(cudaMemcpyAsync(D + d_best, h_best, B * sizeof(int), cudaMemcpyHostToDevice, s));
    CU_TRY(launch_pnp_refine(B, max_iters, words, (const int*)(D + d_off), (const float*)(D + d_obj), (const float*)(D + d_img),
                             (const double*)(D + d_K), (const double*)(D + d_models), (const unsigned*)(D + d_masks),
                             (const int*)(D + d_best), (double*)(D + d_out), (uint8_t*)(D + d_inl), s));
    ctx->launches += 1;
    CU_TRY(cudaMemcpyAsync(h_out, D + d_out, (size_t)B * 12 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(h_inl, D + d_inl, T, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    std::memcpy(inlier, h_inl, T);
    for (int b = 0; b < B; b++) {
      if (stats[b].found) {
        std::memcpy(stats[b].R, h_out + (size_t)b * 12, 9 * sizeof(double));
        std::memcpy(stats[b].t, h_out + (size_t)b * 12 + 9, 3 * sizeof(double));
      } else {
        const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        std::memcpy(stats[b].R, I, sizeof(I));
        stats[b].t[0] = stats[b].t[1] = stats[b].t[2] = 0.0;
      }
    }
    return URMVO_OK;
  } catch (const std::exception& e) {
    return set_error(URMVO_ERR_ARG, std::string("pnp_ransac: ") + e.what());
  }
}

extern "C" int urmvo_pnp_ransac(urmvo_ctx* ctx, int N, const float* obj, const float* img, const double* intr,
                                int max_iters, double reproj_thr, double confidence, uint8_t* inlier,
                                urmvo_pnp_stats* stats) {
  const int32_t off[2] = {0, N};
  return urmvo_pnp_ransac_batch(ctx, 1, off, obj, img, intr, max_iters, reproj_thr, confidence, inlier, stats);
}
