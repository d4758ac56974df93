// adapter/g2o_optimization.cc — drop-in replacement for the reference's src/g2o_optimization.cc
// (LocalmapOptimization :20-177, FrameOptimization :179-321).  Same signatures, same in-place result
// convention, so src/mapping.cc:471 and src/tracking.cc:883 call it unchanged.  It only flattens the
// reference's containers into the SoA arrays of include/urmvo_b200.h; all arithmetic runs in the
// sm_100a kernels.  SolvePnPWithCV (:323-377, an OpenCV call) is NOT part of this path: keep the
// reference's own definition of it in its own translation unit.
#include "g2o_optimization.h"

#include <cstdio>
#include <map>
#include <mutex>
#include <vector>

#include "urmvo_b200.h"

namespace {

urmvo_ctx* ba_context() {
  static urmvo_ctx* ctx = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (urmvo_create(&ctx, 0) != URMVO_OK) {
      std::fprintf(stderr, "[urmvo_b200] %s\n", urmvo_last_error());
      ctx = nullptr;  // no CPU fallback: the calls below fail loudly
    }
  });
  return ctx;
}

void put_pose(const Pose3d& p, double* out) {
  out[0] = p.q.x(); out[1] = p.q.y(); out[2] = p.q.z(); out[3] = p.q.w();
  out[4] = p.p(0); out[5] = p.p(1); out[6] = p.p(2);
}

void get_pose(const double* in, Pose3d& p) {
  p.q.x() = in[0]; p.q.y() = in[1]; p.q.z() = in[2]; p.q.w() = in[3];
  p.p(0) = in[4]; p.p(1) = in[5]; p.p(2) = in[6];
}

}  // namespace

void LocalmapOptimization(MapOfPoses& poses, MapOfPoints3d& points, std::vector<CameraPtr>& camera_list,
                          VectorOfMonoPointConstraints& mono_point_constraints,
                          VectorOfStereoPointConstraints& stereo_point_constraints,
                          const OptimizationConfig& cfg) {
  (void)stereo_point_constraints;  // mono camera: the reference ignores them too (:96)
  urmvo_ctx* ctx = ba_context();
  if (!ctx || poses.empty() || camera_list.empty()) return;
  // dense indices in ascending-id order (std::map order), like g2o's vertex ordering
  std::map<int, int> pose_idx, point_idx;
  std::vector<double> P(poses.size() * 7), X(points.size() * 3);
  std::vector<uint8_t> fixed(poses.size());
  int n = 0;
  for (auto& kv : poses) {
    pose_idx[kv.first] = n;
    put_pose(kv.second, &P[(size_t)n * 7]);
    fixed[n] = kv.second.fixed ? 1 : 0;
    n++;
  }
  n = 0;
  for (auto& kv : points) {
    point_idx[kv.first] = n;
    for (int k = 0; k < 3; k++) X[(size_t)n * 3 + k] = kv.second.p(k);
    n++;
  }
  // observations: constraints whose vertices exist (g2o drops edges with a missing vertex)
  std::vector<double> uv;
  std::vector<int32_t> cam, pt;
  std::vector<size_t> src;
  uv.reserve(mono_point_constraints.size() * 2);
  for (size_t i = 0; i < mono_point_constraints.size(); i++) {
    const MonoPointConstraintPtr& c = mono_point_constraints[i];
    auto pi = pose_idx.find(c->id_pose);
    auto li = point_idx.find(c->id_point);
    if (pi == pose_idx.end() || li == point_idx.end()) continue;
    uv.push_back(c->keypoint(0)); uv.push_back(c->keypoint(1));
    cam.push_back(pi->second); pt.push_back(li->second);
    src.push_back(i);
  }
  CameraPtr& camera = camera_list[mono_point_constraints.empty() ? 0 : mono_point_constraints[0]->id_camera];
  const double intr[4] = {camera->Fx(), camera->Fy(), camera->Cx(), camera->Cy()};
  std::vector<uint8_t> inlier(src.size(), 0);
  const int rc = urmvo_local_ba(ctx, (int)poses.size(), P.data(), fixed.data(), (int)points.size(), X.data(),
                                (int)src.size(), uv.data(), cam.data(), pt.data(), intr, cfg.mono_point,
                                /*it0=*/10, /*it1=*/5, inlier.data(), nullptr, nullptr);
  if (rc != URMVO_OK) {
    std::fprintf(stderr, "[urmvo_b200] LocalmapOptimization: %s\n", urmvo_last_error());
    return;  // inputs untouched, like a g2o optimize() that did nothing
  }
  for (size_t k = 0; k < src.size(); k++) mono_point_constraints[src[k]]->inlier = inlier[k] != 0;
  n = 0;
  for (auto& kv : poses) { get_pose(&P[(size_t)n * 7], kv.second); n++; }
  n = 0;
  for (auto& kv : points) {
    for (int k = 0; k < 3; k++) kv.second.p(k) = X[(size_t)n * 3 + k];
    n++;
  }
}

int FrameOptimization(MapOfPoses& poses, MapOfPoints3d& points, std::vector<CameraPtr>& camera_list,
                      VectorOfMonoPointConstraints& mono_point_constraints,
                      VectorOfStereoPointConstraints& stereo_point_constraints,
                      const OptimizationConfig& cfg) {
  urmvo_ctx* ctx = ba_context();
  const int total = (int)(mono_point_constraints.size() + stereo_point_constraints.size());
  if (!ctx || poses.size() != 1 || camera_list.empty()) return 0;
  MapOfPoses::iterator pose_it = poses.begin();
  double P[7];
  put_pose(pose_it->second, P);
  const int No = (int)mono_point_constraints.size();
  std::vector<double> uv((size_t)No * 2), Xw((size_t)No * 3);
  std::vector<uint8_t> inlier(No);
  for (int i = 0; i < No; i++) {
    const MonoPointConstraintPtr& c = mono_point_constraints[i];
    const Position3d& point = points[c->id_point];  // operator[] like the reference (:214)
    uv[(size_t)i * 2] = c->keypoint(0); uv[(size_t)i * 2 + 1] = c->keypoint(1);
    for (int k = 0; k < 3; k++) Xw[(size_t)i * 3 + k] = point.p(k);
    inlier[i] = c->inlier ? 1 : 0;
  }
  CameraPtr& camera = camera_list[No ? mono_point_constraints[0]->id_camera : 0];
  const double intr[4] = {camera->Fx(), camera->Fy(), camera->Cx(), camera->Cy()};
  const int32_t off[2] = {0, No};
  int32_t n_inlier = 0;
  const int rc = urmvo_pose_only_batch(ctx, 1, off, P, uv.data(), Xw.data(), intr, cfg.mono_point, /*rounds=*/4,
                                       /*its=*/10, inlier.data(), &n_inlier);
  if (rc != URMVO_OK) {
    std::fprintf(stderr, "[urmvo_b200] FrameOptimization: %s\n", urmvo_last_error());
    return 0;
  }
  for (int i = 0; i < No; i++) mono_point_constraints[i]->inlier = inlier[i] != 0;
  get_pose(P, pose_it->second);
  // :319-320 returns #mono + #stereo - #outliers; stereo edges do not exist for a mono camera
  return n_inlier + (total - No);
}
