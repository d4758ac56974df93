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
//
// Fewer than 15 matches (cv::findFundamentalMat leaves the RANSAC branch, fundam.cpp): N == 7 solves the seven
// points once and flags every match; 8 <= N <= 14 is LMeDSPointSetRegistrator::run — a fixed budget
// (RANSACUpdateNumIters(confidence, 0.45, 7, max_iters), at least 3), the score of a model is its median error
// (computed on the device), the first strictly smaller median wins, and the inliers are the matches within
// sigma = 2.5 * 1.4826 * (1 + 5 / (N - 7)) * sqrt(median).  N == 7 and N == 14 reproduce the real OpenCV bit
// for bit; for 8 <= N <= 13 the median is the rounding noise of an exactly-fitted sample point and no two
// implementations agree — the call still returns a valid LMedS answer.
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

// getSubset for the next `want` iterations of one problem (the RNG state persists between rounds);
// returns the number of subsets drawn (< want: the attempt budget ran out, which ends the RANSAC
// loop of that problem).  out: indices + base (position of the problem in the concatenated arrays).
int draw_subsets(CvRng& rng, const float* p0, const float* p1, int N, int want, int base, int* out) {
  for (int it = 0; it < want; it++) {
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
  return want;
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

constexpr int kFirstRound = 256;  // iterations evaluated before the budget is known

// Per-problem state of RANSACPointSetRegistrator::run between rounds.
struct FmProblem {
  CvRng rng;
  int N = 0, base = 0;
  int mode = 0;               // 0: RANSAC (N >= 15), 1: LMedS (8..14), 2: direct 7-point (N == 7)
  double min_median = DBL_MAX;
  int niters = 0, iter = 0, max_good = 0, models = 0, win = -1;  // win = hyp_id*3 + slot
  bool done = false;
  int pos = 0, cnt = 0;  // slice of the current round
};

template <class F>
void parallel_over(int n, F&& fn) {
  const int hw = urmvo::host_threads();
  const int nt = std::max(1, std::min(n / 4, hw > 0 ? hw : 1));  // a thread is worth >= 4 problems
  if (nt <= 1) {
    for (int i = 0; i < n; i++) fn(i);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++)
    th.emplace_back([&, t] { for (int i = t; i < n; i += nt) fn(i); });
  for (auto& t : th) t.join();
}

}  // namespace

struct urmvo_fm_plan {
  urmvo_ctx* ctx = nullptr;
  int B = 0, total_n = 0, max_n = 0, max_iters = 0, evaluated = 0;
  double confidence = 0.99;
  float thr2 = 9.f;
  std::vector<int> off;
  std::vector<float> p0, p1;  // host copies: the subset draws of later rounds read the coordinates
  std::vector<FmProblem> prob;
  bool borrowed = false;
  unsigned char* dev = nullptr;
  size_t o_pts = 0, o_off = 0, o_models = 0, o_ids = 0, o_sets = 0, o_nmod = 0, o_counts = 0, o_win = 0,
         o_winF = 0, o_mask = 0, bytes = 0;
  // pinned host block
  unsigned char* hb = nullptr;
  int *h_ids = nullptr, *h_sets = nullptr, *h_nmod = nullptr, *h_counts = nullptr, *h_win = nullptr;  // h_win: [win | thr bits]
  bool any_small = false;  // a problem with fewer than 15 matches
  double* h_winF = nullptr;
  uint8_t* h_mask = nullptr;
  size_t round_cap = 0;  // hypotheses one round can hold
};

extern "C" void urmvo_fm_plan_destroy(urmvo_fm_plan* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  if (p->borrowed) p->ctx->ws_in_use = false;
  else if (p->dev) cudaFree(p->dev);
  if (p->hb) cudaFreeHost(p->hb);
  delete p;
}

extern "C" int urmvo_fm_plan_hypotheses(const urmvo_fm_plan* p) { return p ? p->evaluated : 0; }

static int fm_plan_create_impl(urmvo_ctx* ctx, urmvo_fm_plan** out, int B, const int32_t* off, const float* pts0,
                               const float* pts1, double thresh, double confidence, int max_iters, bool borrow) try {
  if (!ctx || !out || B <= 0 || !off || !pts0 || !pts1)
    return set_error(URMVO_ERR_ARG, "fm_plan_create: null or empty input");
  *out = nullptr;
  if (max_iters <= 0) max_iters = 1;
  if (thresh <= 0) thresh = 3;  // cv::findFundamentalMat defaults
  if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
  if (off[0] != 0) return set_error(URMVO_ERR_ARG, "fm_plan_create: offsets must start at 0");
  int max_n = 0;
  for (int b = 0; b < B; b++) {
    const int n = off[b + 1] - off[b];
    if (n < 0) return set_error(URMVO_ERR_ARG, "fm_plan_create: offsets must ascend");
    if (n < 7)
      return set_error(URMVO_ERR_UNSUPPORTED,
                       "fm_ransac: fewer than 7 correspondences (cv::findFundamentalMat returns an empty matrix and "
                       "leaves the mask untouched)");
    max_n = std::max(max_n, n);
  }
  if ((long long)B * max_iters > (1ll << 24)) return set_error(URMVO_ERR_ARG, "fm_plan_create: B * max_iters exceeds 2^24 hypotheses (3.6 GB of models); split the batch");
  CU_TRY(cudaSetDevice(ctx->device));
  urmvo_fm_plan* p = new urmvo_fm_plan();
  p->ctx = ctx; p->B = B; p->max_iters = max_iters; p->confidence = confidence;
  p->thr2 = (float)(thresh * thresh);
  p->off.assign(off, off + B + 1);
  p->total_n = off[B];
  p->max_n = max_n;
  p->p0.assign(pts0, pts0 + 2 * (size_t)p->total_n);
  p->p1.assign(pts1, pts1 + 2 * (size_t)p->total_n);
  p->prob.resize(B);
  for (int b = 0; b < B; b++) {
    p->prob[b].N = off[b + 1] - off[b]; p->prob[b].base = off[b];
    p->prob[b].mode = p->prob[b].N >= 15 ? 0 : (p->prob[b].N == 7 ? 2 : 1);
    p->any_small = p->any_small || p->prob[b].mode != 0;
  }
  const size_t T = (size_t)p->total_n, H = (size_t)B * max_iters;
  p->round_cap = H;  // a round never holds more than every remaining iteration of every problem
  auto take = [&](size_t bytes) { size_t o = p->bytes; p->bytes = (p->bytes + bytes + 255) / 256 * 256; return o; };
  p->o_pts = take(T * sizeof(float4));
  p->o_off = take((size_t)(B + 1) * sizeof(int));
  p->o_models = take(H * 27 * sizeof(double));
  p->o_ids = take(H * sizeof(int));
  p->o_sets = take(H * 7 * sizeof(int));
  p->o_nmod = take(H * sizeof(int));
  p->o_counts = take(H * 3 * sizeof(int));
  p->o_win = take((size_t)2 * B * sizeof(int));  // [winning model | the problem's own threshold]
  p->o_winF = take((size_t)B * 9 * sizeof(double));
  p->o_mask = take(T);
  cudaError_t e = cudaSuccess;
  if (borrow && !ctx->ws_in_use) {
    if (ctx->ws_bytes < p->bytes) {
      if (ctx->ws_dev) cudaFree(ctx->ws_dev);
      ctx->ws_dev = nullptr; ctx->ws_bytes = 0;
      e = cudaMalloc(&ctx->ws_dev, p->bytes + p->bytes / 4);
      if (e == cudaSuccess) ctx->ws_bytes = p->bytes + p->bytes / 4;
    }
    if (e == cudaSuccess) { p->dev = ctx->ws_dev; p->borrowed = true; ctx->ws_in_use = true; }
  } else {
    e = cudaMalloc(&p->dev, p->bytes);
  }
  if (e != cudaSuccess) { delete p; return set_error(URMVO_ERR_CUDA, std::string("cudaMalloc fm plan: ") + cudaGetErrorString(e)); }
  // pinned: [winF | ids | sets | nmod | counts | win | mask | pts staging]
  const size_t b_winF = (size_t)B * 9 * sizeof(double), b_ids = H * sizeof(int), b_sets = H * 7 * sizeof(int);
  const size_t b_nmod = H * sizeof(int), b_counts = H * 3 * sizeof(int), b_win = (size_t)2 * B * sizeof(int);
  const size_t b_mask = (T + 15) / 16 * 16, b_pts = T * sizeof(float4);
  const size_t hb_bytes = b_winF + b_ids + b_sets + b_nmod + b_counts + b_win + b_mask + b_pts + 64;
  unsigned char* hb = nullptr;
  if (borrow) {
    if (ctx->ensure_pinned(hb_bytes)) { urmvo_fm_plan_destroy(p); return set_error(URMVO_ERR_CUDA, "cudaMallocHost failed"); }
    hb = (unsigned char*)ctx->pinned;
  } else {
    e = cudaMallocHost(&hb, hb_bytes);
    if (e != cudaSuccess) { urmvo_fm_plan_destroy(p); return set_error(URMVO_ERR_CUDA, "cudaMallocHost fm plan failed"); }
    p->hb = hb;
  }
  unsigned char* q = hb;
  p->h_winF = (double*)q; q += b_winF;
  p->h_ids = (int*)q; q += b_ids;
  p->h_sets = (int*)q; q += b_sets;
  p->h_nmod = (int*)q; q += b_nmod;
  p->h_counts = (int*)q; q += b_counts;
  p->h_win = (int*)q; q += b_win;
  p->h_mask = (uint8_t*)q; q += b_mask;
  q = (unsigned char*)(((uintptr_t)q + 15) & ~(uintptr_t)15);
  float4* h_pts = (float4*)q;
  for (size_t i = 0; i < T; i++) h_pts[i] = make_float4(pts0[2 * i], pts0[2 * i + 1], pts1[2 * i], pts1[2 * i + 1]);
  cudaStream_t s = ctx->stream;
  cudaError_t e1 = cudaMemcpyAsync(p->dev + p->o_pts, h_pts, b_pts, cudaMemcpyHostToDevice, s);
  cudaError_t e2 = cudaMemcpyAsync(p->dev + p->o_off, p->off.data(), (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice, s);
  cudaError_t e3 = cudaStreamSynchronize(s);  // p->off is pageable
  for (cudaError_t ee : {e1, e2, e3})
    if (ee != cudaSuccess) {
      urmvo_fm_plan_destroy(p);
      return set_error(URMVO_ERR_CUDA, std::string("fm_plan_create upload: ") + cudaGetErrorString(ee));
    }
  *out = p;
  return URMVO_OK;
} catch (const std::exception& e) {  // no exception crosses the C ABI
  return set_error(URMVO_ERR_ARG, std::string("fm_plan_create_impl: ") + e.what());
}

extern "C" int urmvo_fm_plan_create(urmvo_ctx* ctx, urmvo_fm_plan** out, int B, const int32_t* off,
                                    const float* pts0, const float* pts1, double thresh, double confidence,
                                    int max_iters) {
  return fm_plan_create_impl(ctx, out, B, off, pts0, pts1, thresh, confidence, max_iters, false);
}

// The RANSAC loop of every problem, in rounds: round 1 evaluates the first kFirstRound iterations of
// each problem, later rounds everything that is left of each problem's (shrinking) budget.  After
// every round the host replays the sequential "better model -> new budget" logic on the counts.
extern "C" int urmvo_fm_plan_run(urmvo_fm_plan* p) try {
  if (!p) return set_error(URMVO_ERR_ARG, "fm_plan_run: null plan");
  CU_TRY(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  unsigned char* D = p->dev;
  for (auto& pr : p->prob) {
    pr.rng = CvRng();
    pr.niters = std::max(p->max_iters, 1);
    pr.iter = 0; pr.max_good = 0; pr.models = 0; pr.win = -1; pr.done = false;
    pr.min_median = DBL_MAX;
    if (pr.mode == 1)  // OpenCV: at least 3; never more than the hypothesis store of the plan holds per problem
      pr.niters = std::min(std::max(update_num_iters(p->confidence, 0.45, 7, std::max(p->max_iters, 1)), 3), std::max(p->max_iters, 1));
    if (pr.mode == 2) pr.niters = 1;
  }
  p->evaluated = 0;
  for (int round = 0;; round++) {
    // ---- plan the round
    int n = 0;
    std::vector<int> active;
    for (int b = 0; b < p->B; b++) {
      FmProblem& pr = p->prob[b];
      if (pr.done) continue;
      const int want = round == 0 ? (pr.mode ? pr.niters : std::min(pr.niters, kFirstRound)) : pr.niters - pr.iter;
      if (want <= 0) { pr.done = true; continue; }
      pr.pos = n; pr.cnt = want;
      n += want;
      active.push_back(b);
    }
    if (n == 0) break;
    // ---- host: draw the subsets of the round (cv::RNG chains, one per problem)
    parallel_over((int)active.size(), [&](int a) {
      FmProblem& pr = p->prob[active[a]];
      int got = 1;
      if (pr.mode == 2) {  // the seven matches themselves, no draw
        for (int k = 0; k < 7; k++) p->h_sets[(size_t)pr.pos * 7 + k] = pr.base + k;
      } else {
        got = draw_subsets(pr.rng, p->p0.data() + 2 * (size_t)pr.base, p->p1.data() + 2 * (size_t)pr.base, pr.N,
                           pr.cnt, pr.base, p->h_sets + (size_t)pr.pos * 7);
      }
      for (int i = 0; i < pr.cnt; i++) p->h_ids[pr.pos + i] = active[a] * p->max_iters + pr.iter + i;
      // subsets past an exhausted attempt budget are never evaluated: mark them with the first one
      for (int i = got; i < pr.cnt; i++)
        for (int k = 0; k < 7; k++) p->h_sets[(size_t)(pr.pos + i) * 7 + k] = pr.base + k;
      if (got < pr.cnt) pr.cnt = -got - 1;  // remember where the draws ended
    });
    CU_TRY(cudaMemcpyAsync(D + p->o_ids, p->h_ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(D + p->o_sets, p->h_sets, (size_t)n * 7 * sizeof(int), cudaMemcpyHostToDevice, s));
    CU_TRY(launch_fm_solve(n, (const int*)(D + p->o_ids), (const int*)(D + p->o_sets), (const float4*)(D + p->o_pts),
                           (double*)(D + p->o_models), (int*)(D + p->o_nmod), s));
    CU_TRY(launch_fm_score(n, (const int*)(D + p->o_ids), p->max_iters, (const int*)(D + p->o_off),
                           (const float4*)(D + p->o_pts), (const double*)(D + p->o_models),
                           (const int*)(D + p->o_nmod), p->thr2, (int*)(D + p->o_counts), p->ctx->n_sm, s));
    p->ctx->launches += 2;
    p->evaluated += n;
    CU_TRY(cudaMemcpyAsync(p->h_nmod, D + p->o_nmod, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(p->h_counts, D + p->o_counts, (size_t)n * 3 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    // ---- replay of RANSACPointSetRegistrator::run on the round's counts
    for (int b : active) {
      FmProblem& pr = p->prob[b];
      int avail = pr.cnt;
      bool exhausted = false;
      if (avail < 0) { avail = -avail - 1; exhausted = true; }
      int i = 0;
      if (pr.mode == 2) {  // direct solution: the first of the <= 3 models, every match flagged
        const int nm = p->h_nmod[pr.pos];
        pr.models = nm; pr.iter = 1; pr.max_good = nm > 0 ? 7 : 0;
        pr.win = nm > 0 ? p->h_ids[pr.pos] * 3 : -1;
        pr.done = true;
        continue;
      }
      if (pr.mode == 1) {  // LMeDSPointSetRegistrator::run: the first strictly smaller median wins
        for (; i < avail && pr.iter < pr.niters; i++, pr.iter++) {
          const int nm = p->h_nmod[pr.pos + i];
          for (int k = 0; k < nm; k++) {
            pr.models++;
            float med;
            std::memcpy(&med, &p->h_counts[(size_t)(pr.pos + i) * 3 + k], sizeof(float));
            if ((double)med < pr.min_median) { pr.min_median = (double)med; pr.win = p->h_ids[pr.pos + i] * 3 + k; }
          }
        }
        pr.done = true;  // the whole budget was one round (or the draws ran out)
        continue;
      }
      for (; i < avail && pr.iter < pr.niters; i++, pr.iter++) {
        const int nm = p->h_nmod[pr.pos + i];
        for (int k = 0; k < nm; k++) {
          pr.models++;
          const int good = p->h_counts[(size_t)(pr.pos + i) * 3 + k];
          if (good > std::max(pr.max_good, 6)) {
            pr.max_good = good;
            pr.win = p->h_ids[pr.pos + i] * 3 + k;
            pr.niters = update_num_iters(p->confidence, (double)(pr.N - good) / pr.N, 7, pr.niters);
          }
        }
      }
      if (pr.iter >= pr.niters || (exhausted && i >= avail)) pr.done = true;
    }
  }
  return URMVO_OK;
} catch (const std::exception& e) {  // no exception crosses the C ABI
  return set_error(URMVO_ERR_ARG, std::string("urmvo_fm_plan_run: ") + e.what());
}

extern "C" int urmvo_fm_plan_finish(urmvo_fm_plan* p, uint8_t* inlier, urmvo_fm_stats* stats) try {
  if (!p) return set_error(URMVO_ERR_ARG, "fm_plan_finish: null plan");
  CU_TRY(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  unsigned char* D = p->dev;
  float* h_thr = (float*)(p->h_win + p->B);
  for (int b = 0; b < p->B; b++) {
    const FmProblem& pr = p->prob[b];
    h_thr[b] = p->thr2;
    if (pr.mode == 0) {
      p->h_win[b] = pr.max_good > 0 ? pr.win : -1;
    } else if (pr.mode == 2) {
      p->h_win[b] = pr.win;
      h_thr[b] = -1.f;  // mask.setTo(1), whether or not the solver found a model
    } else {
      p->h_win[b] = pr.min_median < DBL_MAX ? pr.win : -1;
      if (pr.min_median < DBL_MAX) {
        double sigma = 2.5 * 1.4826 * (1 + 5. / (pr.N - 7)) * std::sqrt(pr.min_median);
        sigma = std::max(sigma, 0.001);
        h_thr[b] = (float)(sigma * sigma);
      }
    }
  }
  CU_TRY(cudaMemcpyAsync(D + p->o_win, p->h_win, (size_t)2 * p->B * sizeof(int), cudaMemcpyHostToDevice, s));
  CU_TRY(launch_fm_mask(p->B, p->max_n, (const int*)(D + p->o_off), (const float4*)(D + p->o_pts),
                        (const double*)(D + p->o_models), (const int*)(D + p->o_win), p->thr2,
                        p->any_small ? (const float*)(D + p->o_win) + p->B : nullptr, D + p->o_mask,
                        (double*)(D + p->o_winF), s));
  p->ctx->launches += 1;
  if (inlier || p->any_small) CU_TRY(cudaMemcpyAsync(p->h_mask, D + p->o_mask, (size_t)p->total_n, cudaMemcpyDeviceToHost, s));
  if (stats) CU_TRY(cudaMemcpyAsync(p->h_winF, D + p->o_winF, (size_t)p->B * 9 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  if (inlier) std::memcpy(inlier, p->h_mask, (size_t)p->total_n);
  if (stats)
    for (int b = 0; b < p->B; b++) {
      const FmProblem& pr = p->prob[b];
      int good = pr.max_good;
      if (pr.mode == 1) {  // LMedS: inliers of the winning model under its own sigma; found = at least 7 of them
        good = 0;
        for (int i = 0; i < pr.N; i++) good += p->h_mask[pr.base + i];
      }
      stats[b].found = pr.mode == 1 ? (pr.min_median < DBL_MAX && good >= 7 ? 1 : 0) : (pr.max_good > 0 ? 1 : 0);
      stats[b].iters = pr.iter;
      stats[b].n_inliers = pr.mode == 2 ? 7 : good;
      stats[b].n_models = pr.models;
      for (int i = 0; i < 9; i++) stats[b].F[i] = p->h_winF[(size_t)b * 9 + i];
    }
  return URMVO_OK;
} catch (const std::exception& e) {  // no exception crosses the C ABI
  return set_error(URMVO_ERR_ARG, std::string("urmvo_fm_plan_finish: ") + e.what());
}

extern "C" int urmvo_fm_ransac_batch(urmvo_ctx* ctx, int B, const int32_t* off, const float* pts0, const float* pts1,
                                     double thresh, double confidence, int max_iters, uint8_t* inlier,
                                     urmvo_fm_stats* stats) {
  if (!inlier) return set_error(URMVO_ERR_ARG, "fm_ransac: null inlier output");
  urmvo_fm_plan* p = nullptr;
  int rc = fm_plan_create_impl(ctx, &p, B, off, pts0, pts1, thresh, confidence, max_iters, true);
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
