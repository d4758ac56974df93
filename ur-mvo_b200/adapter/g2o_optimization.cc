// adapter/g2o_optimization.cc — drop-in replacement for the reference's src/g2o_optimization.cc
// (LocalmapOptimization :20-177, FrameOptimization :179-321, SolvePnPWithCV :323-377).
// Same signatures, same in-place result convention, so src/mapping.cc:471 and src/tracking.cc:883 call
// it unchanged.  It only flattens the reference's containers into the SoA arrays of
// include/urmvo_b200.h; all arithmetic runs in the sm_100a kernels.
//
// Behaviour kept from the reference:
//   * constraints whose pose / point vertex does not exist are dropped (g2o refuses such edges);
//   * the camera type is the type of the camera of the LAST mono constraint (:90, :228): stereo edges
//     are only added when that camera is STEREO (:96, :235) — with no mono constraint at all the type
//     stays MONO and the stereo vector is ignored, exactly like the reference;
//   * an empty graph is a silent no-op (g2o: "0 vertices to optimize");
//   * nothing throws.  On a GPU failure the inputs are left untouched and the failure is recorded:
//     urmvo_adapter_last_status() returns the urmvo_status of the last call on this thread (0 = the
//     map was optimised), urmvo_last_error() the message — the caller can tell an optimised map from
//     an untouched one (include/urmvo_b200.h).
// The reference reads fx, fy, cx, cy (and BF) per edge from camera_list[id_camera] (:86-89, :106-113).  The adapter
// collects the cameras the constraints of a call refer to: when they all have the same intrinsics (always true
// for the reference's configurations: camera_list has one entry) the single-camera entry points run, otherwise
// the camera-model table and a per-edge model index go through urmvo_local_ba_multicam /
// urmvo_pose_only_batch_multicam (more than 128 distinct cameras in one call: URMVO_ERR_UNSUPPORTED).
// Both entry points are called from the tracking thread in the reference; a mutex guards the shared
// context so that a caller with a separate mapping thread stays safe.
#include "g2o_optimization.h"

#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <vector>

#include "urmvo_b200.h"

namespace {

std::mutex g_ba_mutex;
thread_local int g_last_status = URMVO_OK;

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

// the cameras one call refers to: one (fx fy cx cy bf) row per id_camera, in order of first use
struct CameraModels {
  std::vector<int> ids;
  std::vector<double> rows;
  int index(int id_camera, Camera& c) {
    for (size_t m = 0; m < ids.size(); m++) if (ids[m] == id_camera) return (int)m;
    ids.push_back(id_camera);
    const double w[5] = {c.Fx(), c.Fy(), c.Cx(), c.Cy(), c.BF()};
    rows.insert(rows.end(), w, w + 5);
    return (int)ids.size() - 1;
  }
  bool empty() const { return ids.empty(); }
  // all constraints see the same pinhole model (BF only matters when a stereo edge exists)
  bool single(bool with_bf) const {
    for (size_t m = 1; m < ids.size(); m++)
      for (int k = 0; k < (with_bf ? 5 : 4); k++) if (rows[m * 5 + k] != rows[k]) return false;
    return true;
  }
};

int fail_status(int rc, const char* who) {
  g_last_status = rc;
  std::fprintf(stderr, "[urmvo_b200] %s: %s\n", who, rc == URMVO_ERR_UNSUPPORTED && !*urmvo_last_error() ?
               "constraints refer to more than 128 distinct cameras" : urmvo_last_error());
  return rc;
}

}  // namespace

extern "C" int urmvo_adapter_last_status(void) { return g_last_status; }

void LocalmapOptimization(MapOfPoses& poses, MapOfPoints3d& points, std::vector<CameraPtr>& camera_list,
                          VectorOfMonoPointConstraints& mono_point_constraints,
                          VectorOfStereoPointConstraints& stereo_point_constraints,
                          const OptimizationConfig& cfg) {
  g_last_status = URMVO_OK;
  if (poses.empty() || points.empty() || camera_list.empty()) return;  // empty graph: g2o does nothing
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
  CameraType camType = MONO;
  CameraModels K;
  std::vector<double> uv3;
  std::vector<uint8_t> kind, model;
  std::vector<int32_t> cam, pt;
  std::vector<size_t> src_mono, src_stereo;
  uv3.reserve((mono_point_constraints.size() + stereo_point_constraints.size()) * 3);
  for (size_t i = 0; i < mono_point_constraints.size(); i++) {
    const MonoPointConstraintPtr& c = mono_point_constraints[i];
    Camera& camera = *camera_list[c->id_camera];
    camType = camera.GetCameraType();  // :90 — the type of the last mono constraint's camera decides
    auto pi = pose_idx.find(c->id_pose);
    auto li = point_idx.find(c->id_point);
    if (pi == pose_idx.end() || li == point_idx.end()) continue;
    model.push_back((uint8_t)std::min(K.index(c->id_camera, camera), 255));
    uv3.push_back(c->keypoint(0)); uv3.push_back(c->keypoint(1)); uv3.push_back(0.0);
    kind.push_back(0);
    cam.push_back(pi->second); pt.push_back(li->second);
    src_mono.push_back(i);
  }
  if (camType == STEREO) {  // :96
    for (size_t i = 0; i < stereo_point_constraints.size(); i++) {
      const StereoPointConstraintPtr& c = stereo_point_constraints[i];
      auto pi = pose_idx.find(c->id_pose);
      auto li = point_idx.find(c->id_point);
      if (pi == pose_idx.end() || li == point_idx.end()) continue;
      model.push_back((uint8_t)std::min(K.index(c->id_camera, *camera_list[c->id_camera]), 255));
      uv3.push_back(c->keypoint(0)); uv3.push_back(c->keypoint(1)); uv3.push_back(c->keypoint(2));
      kind.push_back(1);
      cam.push_back(pi->second); pt.push_back(li->second);
      src_stereo.push_back(i);
    }
  }
  const size_t No = kind.size();
  if (No == 0) return;  // no edge: a silent no-op in the reference too
  if (K.ids.size() > 128) { fail_status(URMVO_ERR_UNSUPPORTED, "LocalmapOptimization"); return; }
  std::vector<uint8_t> inlier(No, 0);
  int rc;
  {
    std::lock_guard<std::mutex> lock(g_ba_mutex);
    urmvo_ctx* ctx = ba_context();
    if (!ctx) { g_last_status = URMVO_ERR_NO_DEVICE; return; }
    if (!K.single(!src_stereo.empty())) {  // several camera models: per-edge intrinsics (:86-89, :106-113)
      for (size_t o = 0; o < No; o++) kind[o] = (uint8_t)(kind[o] | (model[o] << 1));
      rc = urmvo_local_ba_multicam(ctx, (int)poses.size(), P.data(), fixed.data(), (int)points.size(), X.data(), (int)No,
                                   uv3.data(), kind.data(), cam.data(), pt.data(), (int)K.ids.size(), K.rows.data(),
                                   cfg.mono_point, src_stereo.empty() ? cfg.mono_point : cfg.stereo_point,
                                   /*it0=*/10, /*it1=*/5, inlier.data(), nullptr, nullptr);
    } else if (src_stereo.empty()) {
      std::vector<double> uv(No * 2);
      for (size_t o = 0; o < No; o++) { uv[o * 2] = uv3[o * 3]; uv[o * 2 + 1] = uv3[o * 3 + 1]; }
      rc = urmvo_local_ba(ctx, (int)poses.size(), P.data(), fixed.data(), (int)points.size(), X.data(), (int)No,
                          uv.data(), cam.data(), pt.data(), K.rows.data(), cfg.mono_point, /*it0=*/10, /*it1=*/5,
                          inlier.data(), nullptr, nullptr);
    } else {
      rc = urmvo_local_ba_stereo(ctx, (int)poses.size(), P.data(), fixed.data(), (int)points.size(), X.data(), (int)No,
                                 uv3.data(), kind.data(), cam.data(), pt.data(), K.rows.data(), cfg.mono_point,
                                 cfg.stereo_point, /*it0=*/10, /*it1=*/5, inlier.data(), nullptr, nullptr);
    }
  }
  if (rc != URMVO_OK) { fail_status(rc, "LocalmapOptimization"); return; }  // inputs untouched
  for (size_t k = 0; k < src_mono.size(); k++) mono_point_constraints[src_mono[k]]->inlier = inlier[k] != 0;
  for (size_t k = 0; k < src_stereo.size(); k++)
    stereo_point_constraints[src_stereo[k]]->inlier = inlier[src_mono.size() + k] != 0;
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
  g_last_status = URMVO_OK;
  const int total = (int)(mono_point_constraints.size() + stereo_point_constraints.size());
  if (poses.size() != 1 || camera_list.empty()) { g_last_status = URMVO_ERR_ARG; return 0; }  // :184 assert(poses.size() == 1)
  MapOfPoses::iterator pose_it = poses.begin();
  double P[7];
  put_pose(pose_it->second, P);
  const int Nm = (int)mono_point_constraints.size();
  CameraType camType = MONO;
  CameraModels K;
  std::vector<double> uv3, Xw;
  std::vector<uint8_t> kind, model, inlier;
  uv3.reserve((size_t)total * 3); Xw.reserve((size_t)total * 3);
  for (int i = 0; i < Nm; i++) {
    const MonoPointConstraintPtr& c = mono_point_constraints[i];
    const Position3d& point = points[c->id_point];  // operator[] like the reference (:214)
    Camera& camera = *camera_list[c->id_camera];
    camType = camera.GetCameraType();  // :228
    model.push_back((uint8_t)std::min(K.index(c->id_camera, camera), 255));
    uv3.push_back(c->keypoint(0)); uv3.push_back(c->keypoint(1)); uv3.push_back(0.0);
    for (int k = 0; k < 3; k++) Xw.push_back(point.p(k));
    kind.push_back(0);
    inlier.push_back(c->inlier ? 1 : 0);
  }
  int Ns = 0;
  if (camType == STEREO) {  // :235
    Ns = (int)stereo_point_constraints.size();
    for (int i = 0; i < Ns; i++) {
      const StereoPointConstraintPtr& c = stereo_point_constraints[i];
      const Position3d& point = points[c->id_point];
      model.push_back((uint8_t)std::min(K.index(c->id_camera, *camera_list[c->id_camera]), 255));
      uv3.push_back(c->keypoint(0)); uv3.push_back(c->keypoint(1)); uv3.push_back(c->keypoint(2));
      for (int k = 0; k < 3; k++) Xw.push_back(point.p(k));
      kind.push_back(1);
      inlier.push_back(c->inlier ? 1 : 0);
    }
  }
  const int No = Nm + Ns;
  if (K.empty()) {  // no edge at all: g2o optimises nothing, every round counts zero outliers
    return total;
  }
  if (K.ids.size() > 128) { fail_status(URMVO_ERR_UNSUPPORTED, "FrameOptimization"); return 0; }
  const int32_t off[2] = {0, No};
  int32_t n_inlier = 0;
  int rc;
  {
    std::lock_guard<std::mutex> lock(g_ba_mutex);
    urmvo_ctx* ctx = ba_context();
    if (!ctx) { g_last_status = URMVO_ERR_NO_DEVICE; return 0; }
    if (!K.single(Ns > 0)) {  // several camera models: per-edge intrinsics (:221-224, :243-250)
      for (int o = 0; o < No; o++) kind[o] = (uint8_t)(kind[o] | (model[o] << 1));
      rc = urmvo_pose_only_batch_multicam(ctx, 1, off, P, uv3.data(), kind.data(), Xw.data(), (int)K.ids.size(),
                                          K.rows.data(), cfg.mono_point, Ns > 0 ? cfg.stereo_point : cfg.mono_point,
                                          /*rounds=*/4, /*its=*/10, inlier.data(), &n_inlier);
    } else if (Ns == 0) {
      std::vector<double> uv((size_t)No * 2);
      for (int o = 0; o < No; o++) { uv[(size_t)o * 2] = uv3[(size_t)o * 3]; uv[(size_t)o * 2 + 1] = uv3[(size_t)o * 3 + 1]; }
      rc = urmvo_pose_only_batch(ctx, 1, off, P, uv.data(), Xw.data(), K.rows.data(), cfg.mono_point, /*rounds=*/4,
                                 /*its=*/10, inlier.data(), &n_inlier);
    } else {
      rc = urmvo_pose_only_batch_stereo(ctx, 1, off, P, uv3.data(), kind.data(), Xw.data(), K.rows.data(), cfg.mono_point,
                                        cfg.stereo_point, /*rounds=*/4, /*its=*/10, inlier.data(), &n_inlier);
    }
  }
  if (rc != URMVO_OK) { fail_status(rc, "FrameOptimization"); return 0; }
  for (int i = 0; i < Nm; i++) mono_point_constraints[i]->inlier = inlier[i] != 0;
  for (int i = 0; i < Ns; i++) stereo_point_constraints[i]->inlier = inlier[Nm + i] != 0;
  get_pose(P, pose_it->second);
  // :319-320 returns #mono + #stereo - #outliers (stereo constraints of a MONO camera are never edges)
  return n_inlier + (total - No);
}


// SolvePnPWithCV (reference src/g2o_optimization.cc:323-377): same gather (valid mappoints with a keypoint, float
// positions), same "< 8 correspondences -> 0" guard, the OpenCV call replaced by urmvo_pnp_ransac, the same
// conversion of T_cw to pose = T_wc and of the inlier list to mappoint ids.  A GPU failure returns 0 (the
// reference's `catch (...) return 0`) and is recorded in urmvo_adapter_last_status().
int SolvePnPWithCV(FramePtr frame, std::vector<MappointPtr>& mappoints, Eigen::Matrix4d& pose, std::vector<int>& inliers) {
  g_last_status = URMVO_OK;
  std::vector<float> object_points, image_points;
  std::vector<int> point_indexes;
  CameraPtr camera = frame->GetCamera();
  for (size_t i = 0; i < mappoints.size(); i++) {
    MappointPtr mpt = mappoints[i];
    if (mpt == nullptr || !mpt->IsValid()) continue;
    Eigen::Vector2d keypoint;
    if (!frame->GetKeypointPosition(i, keypoint)) continue;
    const Eigen::Vector3d& point_position = mpt->GetPosition();
    for (int k = 0; k < 3; k++) object_points.push_back((float)point_position(k));
    image_points.push_back((float)keypoint(0));
    image_points.push_back((float)keypoint(1));
    point_indexes.push_back((int)i);
  }
  const int n = (int)point_indexes.size();
  if (n < 8) return 0;
  urmvo_ctx* ctx = ba_context();
  if (!ctx) { fail_status(URMVO_ERR_NO_DEVICE, "SolvePnPWithCV"); return 0; }
  const double intr[4] = {camera->Fx(), camera->Fy(), camera->Cx(), camera->Cy()};
  std::vector<uint8_t> flags(n);
  urmvo_pnp_stats st;
  int rc;
  {
    std::lock_guard<std::mutex> lock(g_ba_mutex);
    rc = urmvo_pnp_ransac(ctx, n, object_points.data(), image_points.data(), intr, 100, 20.0, 0.99, flags.data(), &st);
  }
  if (rc != URMVO_OK) { fail_status(rc, "SolvePnPWithCV"); return 0; }
  // pose = [R_wc | -R_wc t_cw] (:357-366); the other entries of `pose` are the caller's (identity in tracking.cc:797)
  for (int r = 0; r < 3; r++) {
    double tw = 0;
    for (int c = 0; c < 3; c++) {
      pose(r, c) = st.R[c * 3 + r];
      tw -= st.R[c * 3 + r] * st.t[c];
    }
    pose(r, 3) = tw;
  }
  inliers = std::vector<int>(mappoints.size(), -1);
  int n_in = 0;
  for (int i = 0; i < n; i++)
    if (flags[i]) {
      const int point_idx = point_indexes[i];
      inliers[point_idx] = mappoints[point_idx]->GetId();
      n_in++;
    }
  return st.found ? n_in : 0;
}
