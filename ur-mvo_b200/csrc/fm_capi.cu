// csrc/fm_capi.cu — C ABI of the per-frame fundamental-matrix RANSAC (include/urmvo_b200.h,
// "per-frame outlier rejection of the matcher"; reference src/point_matching.cc:44-58).
//
// Host work (everything that is inherently sequential in OpenCV's RANSACPointSetRegistrator::run):
//   * drawing the 7-index subsets with cv::RNG(-1) + FMEstimatorCallback::checkSubset re-draws — the
//     sequence does not depend on any model, so all max_iters subsets are drawn up front, one host
//     thread per problem of a batch;
//   * replaying "goodCount > max(maxGoodCount, 6) -> new best, niters = RANSACUpdateNumIters(...)"
//     over the per-(iteration, model) inlier counts the device returns.
// The models, the errors and the inlier flags are computed by fm_kernels.cu.  No CPU fallback.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "capi_internal.h"
#include "kernels.h"

using namespace urmvo;

namespace {

// cv::RNG: multiply-with-carry, state 0xffffffffffffffff in RANSACPointSetRegistrator::run
struct CvRng {
  uint64_t state = 0xffffffffffffffffull;
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a + (int)(next() % (unsigned)(b - a)); }
};

// haveCollinearPoints: only the last point of the subset is tested against all earlier pairs
bool last_point_collinear(const float* m, const int* idx) {
  const float* pi = m + 2 * idx[6];
  for (int j = 0; j < 6; j++) {
    const double dx1 = m[2 * idx[j]] - pi[0], dy1 = m[2 * idx[j] + 1] - pi[1];
    for (int k = 0; k < j; k++) {
      const double dx2 = m[2 * idx[k]] - pi[0], dy2 = m[2 * idx[k] + 1] - pi[1];
      if (std::fabs(dx2 * dy1 - dy2 * dx1) <=
          FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2)))
        return true;
    }
  }
  return false;
}

// getSubset for iterations 0..max_iters-1; returns the number of subsets (an exhausted attempt
// budget ends the RANSAC loop).  out: local indices + base.
int draw_subsets(const float* p0, const float* p1, int N, int max_iters, int base, int* out) {
  CvRng rng;
  for (int it = 0; it < max_iters; it++) {
    int idx[7];
    bool found = false;
    for (int attempt = 0; attempt < 10000 && !found; attempt++) {
      for (int i = 0; i < 7; i++) {
        int v = rng.uniform(0, N);
        for (;;) {
          bool dup = false;
          for (int j = 0; j < i; j++) dup |= idx[j] == v;
          if (!dup) break;
          v = rng.uniform(0, N);
        }
        idx[i] = v;
      }
      found = !last_point_collinear(p0, idx) && !last_point_collinear(p1, idx);
    }
    if (!found) return it;
    for (int i = 0; i < 7; i++) out[(size_t)it * 7 + i] = base + idx[i];
  }
  return max_iters;
}

// RANSACUpdateNumIters
int update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = std::min(std::max(p, 0.), 1.);
  ep = std::min(std::max(ep, 0.), 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::nearbyint(num / denom);
}

}  // namespace

struct urmvo_fm_plan {
  urmvo_ctx* ctx = nullptr;
  int B = 0, total_n = 0, total_hyp = 0, max_n = 0, max_iters = 0;
  double confidence = 0.99;
  float thr2 = 9.f;
  std::vector<int> off, hyp_off;  // B+1 each
  unsigned char* dev = nullptr;
  size_t o_pts = 0, o_off = 0, o_sets = 0, o_prob = 0, o_models = 0, o_nmod = 0, o_counts = 0, o_win = 0,
         o_found = 0, o_mask = 0, bytes = 0;
  // pinned host mirrors of what comes back / goes up per finish
  int* h_counts = nullptr;   // total_hyp*3
  int* h_nmod = nullptr;     // total_hyp
  double* h_models = nullptr;  // total_hyp*27 (winner lookup)
  double* h_win = nullptr;   // B*9
  int* h_found = nullptr;    // B
  uint8_t* h_mask = nullptr; // total_n
};

extern "C" void urmvo_fm_plan_destroy(urmvo_fm_plan* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  if (p->dev) cudaFree(p->dev);
  if (p->h_counts) cudaFreeHost(p->h_counts);
  delete p;
}

extern "C" int urmvo_fm_plan_hypotheses(const urmvo_fm_plan* p) { return p ? p->total_hyp : 0; }

extern "C" int urmvo_fm_plan_create(urmvo_ctx* ctx, urmvo_fm_plan** out, int B, const int32_t* off,
                                    const float* pts0, const float* pts1, double thresh, double confidence,
                                    int max_iters) {
  if (!ctx || !out || B <= 0 || !off || !pts0 || !pts1)
    return set_error(URMVO_ERR_ARG, "fm_plan_create: null or empty input");
  *out = nullptr;
  if (max_iters <= 0) max_iters = 1;
  if (thresh <= 0) thresh = 3;  // cv::findFundamentalMat defaults
  if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
  int max_n = 0;
  for (int b = 0; b < B; b++) {
    const int n = off[b + 1] - off[b];
    if (off[0] != 0 || n < 0) return set_error(URMVO_ERR_ARG, "fm_plan_create: offsets must start at 0 and ascend");
    if (n < 15)
      return set_error(URMVO_ERR_UNSUPPORTED,
                       "fm_ransac: fewer than 15 correspondences (OpenCV's 7-point / LMedS branch; keep the "
                       "reference's own cv::findFundamentalMat call for these)");
    max_n = std::max(max_n, n);
  }
  CU_TRY(cudaSetDevice(ctx->device));
  urmvo_fm_plan* p = new urmvo_fm_plan();
  p->ctx = ctx; p->B = B; p->max_iters = max_iters; p->confidence = confidence;
  p->thr2 = (float)(thresh * thresh);
  p->off.assign(off, off + B + 1);
  p->total_n = off[B];
  p->max_n = max_n;
  // ---- host: subsets of every problem (threads over problems)
  std::vector<int> sets((size_t)B * max_iters * 7);
  std::vector<int> n_sets(B, 0);
  {
    const int hw = (int)std::thread::hardware_concurrency();
    const int nt = std::max(1, std::min(B, hw > 0 ? hw : 1));
    auto work = [&](int t) {
      for (int b = t; b < B; b += nt)
        n_sets[b] = draw_subsets(pts0 + 2 * (size_t)off[b], pts1 + 2 * (size_t)off[b], off[b + 1] - off[b], max_iters,
                                 off[b], sets.data() + (size_t)b * max_iters * 7);
    };
    if (nt == 1) {
      work(0);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) th.emplace_back(work, t);
      for (auto& t : th) t.join();
    }
  }
  p->hyp_off.assign(B + 1, 0);
  for (int b = 0; b < B; b++) p->hyp_off[b + 1] = p->hyp_off[b] + n_sets[b];
  p->total_hyp = p->hyp_off[B];
  const size_t H = (size_t)std::max(p->total_hyp, 1), T = (size_t)p->total_n;
  auto take = [&](size_t bytes) { size_t o = p->bytes; p->bytes = (p->bytes + bytes + 255) / 256 * 256; return o; };
  p->o_pts = take(T * sizeof(float4));
  p->o_off = take((size_t)(B + 1) * sizeof(int));
  p->o_sets = take(H * 7 * sizeof(int));
  p->o_prob = take(H * sizeof(int));
  p->o_models = take(H * 27 * sizeof(double));
  p->o_nmod = take(H * sizeof(int));
  p->o_counts = take(H * 3 * sizeof(int));
  p->o_win = take((size_t)B * 9 * sizeof(double));
  p->o_found = take((size_t)B * sizeof(int));
  p->o_mask = take(T);
  cudaError_t e = cudaMalloc(&p->dev, p->bytes);
  if (e != cudaSuccess) { delete p; return set_error(URMVO_ERR_CUDA, std::string("cudaMalloc fm plan: ") + cudaGetErrorString(e)); }
  // one pinned block: [counts | nmod | models | win | found | mask] + upload staging [pts | sets | prob]
  const size_t hb_counts = H * 3 * sizeof(int), hb_nmod = H * sizeof(int), hb_models = H * 27 * sizeof(double);
  const size_t hb_win = (size_t)B * 9 * sizeof(double), hb_found = (size_t)B * sizeof(int), hb_mask = (T + 15) / 16 * 16;
  const size_t hb_pts = T * sizeof(float4), hb_sets = H * 7 * sizeof(int), hb_prob = H * sizeof(int);
  unsigned char* hb = nullptr;
  e = cudaMallocHost(&hb, hb_models + hb_win + hb_counts + hb_nmod + hb_found + hb_mask + hb_pts + hb_sets + hb_prob + 64);  // + alignment padding
  if (e != cudaSuccess) { cudaFree(p->dev); delete p; return set_error(URMVO_ERR_CUDA, "cudaMallocHost fm plan failed"); }
  p->h_counts = (int*)hb;  // first member: the block is freed through h_counts
  unsigned char* q = hb + hb_counts;
  p->h_nmod = (int*)q; q += hb_nmod;
  // keep 8-byte alignment for the doubles
  q = (unsigned char*)(((uintptr_t)q + 7) & ~(uintptr_t)7);
  p->h_models = (double*)q; q += hb_models;
  p->h_win = (double*)q; q += hb_win;
  p->h_found = (int*)q; q += hb_found;
  p->h_mask = (uint8_t*)q; q += hb_mask;
  q = (unsigned char*)(((uintptr_t)q + 15) & ~(uintptr_t)15);
  float4* h_pts = (float4*)q; q += hb_pts;
  int* h_sets = (int*)q; q += hb_sets;
  int* h_prob = (int*)q;
  for (size_t i = 0; i < T; i++) h_pts[i] = make_float4(pts0[2 * i], pts0[2 * i + 1], pts1[2 * i], pts1[2 * i + 1]);
  for (int b = 0; b < B; b++) {
    std::memcpy(h_sets + (size_t)p->hyp_off[b] * 7, sets.data() + (size_t)b * max_iters * 7, (size_t)n_sets[b] * 7 * sizeof(int));
    for (int h = p->hyp_off[b]; h < p->hyp_off[b + 1]; h++) h_prob[h] = b;
  }
  cudaStream_t s = ctx->stream;
  cudaError_t e1 = cudaMemcpyAsync(p->dev + p->o_pts, h_pts, hb_pts, cudaMemcpyHostToDevice, s);
  cudaError_t e2 = cudaMemcpyAsync(p->dev + p->o_off, p->off.data(), (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice, s);
  cudaError_t e3 = p->total_hyp ? cudaMemcpyAsync(p->dev + p->o_sets, h_sets, (size_t)p->total_hyp * 7 * sizeof(int), cudaMemcpyHostToDevice, s) : cudaSuccess;
  cudaError_t e4 = p->total_hyp ? cudaMemcpyAsync(p->dev + p->o_prob, h_prob, (size_t)p->total_hyp * sizeof(int), cudaMemcpyHostToDevice, s) : cudaSuccess;
  cudaError_t e5 = cudaStreamSynchronize(s);  // p->off is pageable, the staging is reused
  for (cudaError_t ee : {e1, e2, e3, e4, e5})
    if (ee != cudaSuccess) {
      urmvo_fm_plan_destroy(p);
      return set_error(URMVO_ERR_CUDA, std::string("fm_plan_create upload: ") + cudaGetErrorString(ee));
    }
  *out = p;
  return URMVO_OK;
}

extern "C" int urmvo_fm_plan_run(urmvo_fm_plan* p) {
  if (!p) return set_error(URMVO_ERR_ARG, "fm_plan_run: null plan");
  CU_TRY(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  unsigned char* D = p->dev;
  if (p->total_hyp == 0) return URMVO_OK;
  CU_TRY(launch_fm_solve(p->total_hyp, (const int*)(D + p->o_sets), (const float4*)(D + p->o_pts),
                         (double*)(D + p->o_models), (int*)(D + p->o_nmod), s));
  CU_TRY(launch_fm_score(p->total_hyp, (const int*)(D + p->o_prob), (const int*)(D + p->o_off),
                         (const float4*)(D + p->o_pts), (const double*)(D + p->o_models),
                         (const int*)(D + p->o_nmod), p->thr2, (int*)(D + p->o_counts), p->ctx->n_sm, s));
  p->ctx->launches += 2;
  return URMVO_OK;
}

extern "C" int urmvo_fm_plan_finish(urmvo_fm_plan* p, uint8_t* inlier, urmvo_fm_stats* stats) {
  if (!p) return set_error(URMVO_ERR_ARG, "fm_plan_finish: null plan");
  CU_TRY(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  unsigned char* D = p->dev;
  const size_t H = (size_t)p->total_hyp;
  if (H) {
    CU_TRY(cudaMemcpyAsync(p->h_counts, D + p->o_counts, H * 3 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(p->h_nmod, D + p->o_nmod, H * sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(p->h_models, D + p->o_models, H * 27 * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  CU_TRY(cudaStreamSynchronize(s));
  // ---- replay of RANSACPointSetRegistrator::run per problem
  for (int b = 0; b < p->B; b++) {
    const int N = p->off[b + 1] - p->off[b];
    const int h0 = p->hyp_off[b], avail = p->hyp_off[b + 1] - h0;
    int niters = std::max(p->max_iters, 1), max_good = 0, iter = 0, models = 0, best_h = -1, best_k = 0;
    for (iter = 0; iter < niters; iter++) {
      if (iter >= avail) break;  // getSubset exhausted its attempts (iter == 0: no model at all)
      const int h = h0 + iter;
      const int nm = p->h_nmod[h];
      for (int k = 0; k < nm; k++) {
        models++;
        const int good = p->h_counts[(size_t)h * 3 + k];
        if (good > std::max(max_good, 6)) {
          max_good = good;
          best_h = h; best_k = k;
          niters = update_num_iters(p->confidence, (double)(N - good) / N, 7, niters);
        }
      }
    }
    p->h_found[b] = max_good > 0 ? 1 : 0;
    for (int i = 0; i < 9; i++) p->h_win[(size_t)b * 9 + i] = best_h >= 0 ? p->h_models[(size_t)best_h * 27 + 9 * best_k + i] : 0.0;
    if (stats) {
      stats[b].found = p->h_found[b];
      stats[b].iters = iter;
      stats[b].n_inliers = max_good;
      stats[b].n_models = models;
      for (int i = 0; i < 9; i++) stats[b].F[i] = p->h_win[(size_t)b * 9 + i];
    }
  }
  CU_TRY(cudaMemcpyAsync(D + p->o_win, p->h_win, (size_t)p->B * 9 * sizeof(double), cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(D + p->o_found, p->h_found, (size_t)p->B * sizeof(int), cudaMemcpyHostToDevice, s));
  CU_TRY(launch_fm_mask(p->B, p->max_n, (const int*)(D + p->o_off), (const float4*)(D + p->o_pts),
                        (const double*)(D + p->o_win), (const int*)(D + p->o_found), p->thr2, D + p->o_mask, s));
  p->ctx->launches += 1;
  if (inlier) {
    CU_TRY(cudaMemcpyAsync(p->h_mask, D + p->o_mask, (size_t)p->total_n, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    std::memcpy(inlier, p->h_mask, (size_t)p->total_n);
  } else {
    CU_TRY(cudaStreamSynchronize(s));
  }
  return URMVO_OK;
}

extern "C" int urmvo_fm_ransac_batch(urmvo_ctx* ctx, int B, const int32_t* off, const float* pts0, const float* pts1,
                                     double thresh, double confidence, int max_iters, uint8_t* inlier,
                                     urmvo_fm_stats* stats) {
  if (!inlier) return set_error(URMVO_ERR_ARG, "fm_ransac: null inlier output");
  urmvo_fm_plan* p = nullptr;
  int rc = urmvo_fm_plan_create(ctx, &p, B, off, pts0, pts1, thresh, confidence, max_iters);
  if (rc != URMVO_OK) return rc;
  rc = urmvo_fm_plan_run(p);
  if (rc == URMVO_OK) rc = urmvo_fm_plan_finish(p, inlier, stats);
  urmvo_fm_plan_destroy(p);
  return rc;
}

extern "C" int urmvo_fm_ransac(urmvo_ctx* ctx, int N, const float* pts0, const float* pts1, double thresh,
                               double confidence, int max_iters, uint8_t* inlier, urmvo_fm_stats* stats) {
  const int32_t off[2] = {0, N};
  return urmvo_fm_ransac_batch(ctx, 1, off, pts0, pts1, thresh, confidence, max_iters, inlier, stats);
}
