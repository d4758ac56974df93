// adapter/point_matching_outliers.cc — see point_matching_outliers.h.  Only marshals the two
// std::vector<cv::Point2f> (already contiguous x,y floats) into urmvo_fm_ransac; the RANSAC itself
// runs in csrc/fm_kernels.cu.  No CPU fallback and no abort: a GPU failure is reported to the caller
// (FindFundamentalInliersStatus), who decides whether to keep the reference's own OpenCV call for that
// frame or to stop.
#include "point_matching_outliers.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "urmvo_b200.h"

namespace {
urmvo_ctx* fm_context() {
  static urmvo_ctx* ctx = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (urmvo_create(&ctx, 0) != URMVO_OK) {
      std::fprintf(stderr, "[urmvo_b200] %s\n", urmvo_last_error());
      ctx = nullptr;
    }
  });
  return ctx;
}
std::mutex g_fm_mutex;  // one context, one call at a time (the reference calls MatchingPoints from the tracking thread)
thread_local int g_fm_status = URMVO_OK;
}  // namespace

int FindFundamentalInliersStatus() { return g_fm_status; }

bool FindFundamentalInliersGPU(const std::vector<cv::Point2f>& points0, const std::vector<cv::Point2f>& points1,
                               std::vector<unsigned char>& inliers) {
  static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f must be two packed floats");
  const int n = (int)points0.size();
  g_fm_status = URMVO_ERR_UNSUPPORTED;
  if (n < 7 || points1.size() != points0.size()) return false;  // OpenCV: empty matrix, mask never created
  urmvo_ctx* ctx = fm_context();
  if (!ctx) {
    g_fm_status = URMVO_ERR_NO_DEVICE;
    std::fprintf(stderr, "[urmvo_b200] FindFundamentalInliersGPU: no usable B200 (there is no CPU path in this library)\n");
    return false;
  }
  std::lock_guard<std::mutex> lock(g_fm_mutex);
  std::vector<unsigned char> out(n, 0);
  const int rc = urmvo_fm_ransac(ctx, n, &points0[0].x, &points1[0].x, 3.0, 0.99, 1000, out.data(), nullptr);
  g_fm_status = rc;
  if (rc != URMVO_OK) {  // e.g. a failed cudaMalloc: `inliers` stays untouched, the caller sees false + the status
    std::fprintf(stderr, "[urmvo_b200] urmvo_fm_ransac: %s\n", urmvo_last_error());
    return false;
  }
  inliers.swap(out);
  return true;
}
