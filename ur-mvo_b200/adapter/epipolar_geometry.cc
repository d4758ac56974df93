// adapter/epipolar_geometry.cc — drop-in replacement for the reference's src/epipolar_geometry.cc.
// EpipolarGeometry keeps its header (include/epipolar_geometry.h:20-48); reconstruct() draws the
// 8-point sets with the reference's own scheme (glibc rand(), :53-71, :100-112) on the host and runs
// everything else — normalisation, H and F RANSAC, model selection, motion recovery, triangulation —
// in the sm_100a kernels through the C ABI.
#include "epipolar_geometry.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "urmvo_b200.h"

bool EpipolarGeometry::Random::already_seeded = false;

void EpipolarGeometry::Random::seed_rand(int seed) { srand(seed); }

void EpipolarGeometry::Random::seed_rand_once(int seed) {
  if (!already_seeded) {
    seed_rand(seed);
    already_seeded = true;
  }
}

int EpipolarGeometry::Random::RandomInt(int min, int max) {
  int d = max - min + 1;
  return int(((double)rand() / ((double)RAND_MAX + 1.0)) * d) + min;
}

EpipolarGeometry::EpipolarGeometry(const Eigen::Matrix3f& k, float sigma, int iterations)
    : _K(k), _Sigma(sigma), _Sigma2(sigma * sigma), _MaxIterations(iterations) {}

namespace {
urmvo_ctx* tv_context() {  // its own context: reconstruct() runs on the feature thread
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
}  // namespace

bool EpipolarGeometry::reconstruct(const std::vector<cv::KeyPoint>& vKeys1, const std::vector<cv::KeyPoint>& vKeys2,
                                   const std::vector<int> vMatches12, Eigen::Matrix4f& T21,
                                   std::vector<cv::Point3f>& vP3D, std::vector<bool>& vbTriangulated) {
  urmvo_ctx* ctx = tv_context();
  if (!ctx) return false;
  const int n1 = (int)vKeys1.size(), n2 = (int)vKeys2.size();
  std::vector<float> k1((size_t)n1 * 2), k2((size_t)n2 * 2);
  for (int i = 0; i < n1; i++) { k1[(size_t)i * 2] = vKeys1[i].pt.x; k1[(size_t)i * 2 + 1] = vKeys1[i].pt.y; }
  for (int i = 0; i < n2; i++) { k2[(size_t)i * 2] = vKeys2[i].pt.x; k2[(size_t)i * 2 + 1] = vKeys2[i].pt.y; }
  std::vector<int32_t> m12(n1, -1);
  int N = 0;
  for (int i = 0; i < n1 && i < (int)vMatches12.size(); i++) {
    m12[i] = vMatches12[i];
    if (vMatches12[i] >= 0) N++;
  }
  if (N < 8) return false;
  // minimum sets, drawn exactly like the reference (swap-with-back over the available indices)
  std::vector<int32_t> sets((size_t)_MaxIterations * 8);
  std::vector<size_t> all(N), avail;
  for (int i = 0; i < N; i++) all[i] = i;
  Random::seed_rand_once(0);
  for (int it = 0; it < _MaxIterations; it++) {
    avail = all;
    for (size_t j = 0; j < 8; j++) {
      int randi = Random::RandomInt(0, (int)avail.size() - 1);
      sets[(size_t)it * 8 + j] = (int32_t)avail[randi];
      avail[randi] = avail.back();
      avail.pop_back();
    }
  }
  float K[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) K[r * 3 + c] = _K(r, c);
  float T[16];
  std::vector<float> P3D((size_t)n1 * 3);
  std::vector<uint8_t> tri(n1);
  int ok = 0;
  const int rc = urmvo_two_view(ctx, n1, k1.data(), n2, k2.data(), m12.data(), K, _Sigma, _MaxIterations, sets.data(),
                                T, P3D.data(), tri.data(), nullptr, nullptr, nullptr, &ok);
  if (rc != URMVO_OK) {
    std::fprintf(stderr, "[urmvo_b200] EpipolarGeometry::reconstruct: %s\n", urmvo_last_error());
    return false;
  }
  if (!ok) return false;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) T21(r, c) = T[r * 4 + c];
  // the reference's homography branch forgets to assign vP3D; we always return it
  vP3D.resize(n1);
  vbTriangulated.assign(n1, false);
  for (int i = 0; i < n1; i++) {
    vP3D[i] = cv::Point3f(P3D[(size_t)i * 3], P3D[(size_t)i * 3 + 1], P3D[(size_t)i * 3 + 2]);
    vbTriangulated[i] = tri[i] != 0;
  }
  return true;
}
