// tests/shim/adapter_driver.cc — drives the drop-in adapters through the reference-facing API.
// Reads a little binary problem file written by tests/test_adapter.py, writes the results back.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "epipolar_geometry.h"
#include "g2o_optimization.h"
#include "point_matching_outliers.h"
extern "C" int urmvo_adapter_last_status(void);
template <class T> static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { perror("read"); exit(2); } return v; }
template <class T> static void wr(FILE* f, const std::vector<T>& v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }
int main(int argc, char** argv) {
  if (argc < 4) return 1;
  const std::string mode = argv[1];
  FILE* in = fopen(argv[2], "rb"); FILE* out = fopen(argv[3], "wb");
  if (!in || !out) return 1;
  std::vector<CameraPtr> cams;
  OptimizationConfig cfg{10.0, 75.0, 0.5};
  if (mode == "ba") {
    auto hdr = rd<int>(in, 3); int Nc = hdr[0], Np = hdr[1], No = hdr[2];
    auto intr = rd<double>(in, 4); auto ids = rd<int>(in, Nc); auto P = rd<double>(in, (size_t)Nc * 7); auto fx = rd<unsigned char>(in, Nc);
    auto pids = rd<int>(in, Np); auto X = rd<double>(in, (size_t)Np * 3);
    auto uv = rd<double>(in, (size_t)No * 2); auto oc = rd<int>(in, No); auto op = rd<int>(in, No);
    cams.emplace_back(new Camera(intr[0], intr[1], intr[2], intr[3]));
    MapOfPoses poses; MapOfPoints3d points; VectorOfMonoPointConstraints mono; VectorOfStereoPointConstraints stereo;
    for (int c = 0; c < Nc; c++) { Pose3d p; p.fixed = fx[c]; p.q.x() = P[c*7]; p.q.y() = P[c*7+1]; p.q.z() = P[c*7+2]; p.q.w() = P[c*7+3]; for (int k = 0; k < 3; k++) p.p(k) = P[c*7+4+k]; poses[ids[c]] = p; }
    for (int l = 0; l < Np; l++) { Position3d q; for (int k = 0; k < 3; k++) q.p(k) = X[l*3+k]; points[pids[l]] = q; }
    for (int o = 0; o < No; o++) { auto m = std::make_shared<MonoPointConstraint>(); m->id_pose = ids[oc[o]]; m->id_point = pids[op[o]]; m->id_camera = 0; m->inlier = true; m->keypoint(0) = uv[o*2]; m->keypoint(1) = uv[o*2+1]; m->pixel_sigma = 0.8; mono.push_back(m); }
    LocalmapOptimization(poses, points, cams, mono, stereo, cfg);
    std::vector<double> Po, Xo; std::vector<unsigned char> inl;
    for (auto& kv : poses) { Po.push_back(kv.second.q.x()); Po.push_back(kv.second.q.y()); Po.push_back(kv.second.q.z()); Po.push_back(kv.second.q.w()); for (int k = 0; k < 3; k++) Po.push_back(kv.second.p(k)); }
    for (auto& kv : points) for (int k = 0; k < 3; k++) Xo.push_back(kv.second.p(k));
    for (auto& m : mono) inl.push_back(m->inlier);
    wr(out, Po); wr(out, Xo); wr(out, inl);
  } else if (mode == "ba_stereo") {
    // mono + stereo constraints of a STEREO camera: kind[o] = 1 -> StereoPointConstraint (u, v, u_right)
    auto hdr = rd<int>(in, 3); int Nc = hdr[0], Np = hdr[1], No = hdr[2];
    auto intr = rd<double>(in, 5); auto ids = rd<int>(in, Nc); auto P = rd<double>(in, (size_t)Nc * 7); auto fx = rd<unsigned char>(in, Nc);
    auto pids = rd<int>(in, Np); auto X = rd<double>(in, (size_t)Np * 3);
    auto uv3 = rd<double>(in, (size_t)No * 3); auto kind = rd<unsigned char>(in, No); auto oc = rd<int>(in, No); auto op = rd<int>(in, No);
    cams.emplace_back(new Camera(intr[0], intr[1], intr[2], intr[3], STEREO, intr[4]));
    MapOfPoses poses; MapOfPoints3d points; VectorOfMonoPointConstraints mono; VectorOfStereoPointConstraints stereo;
    for (int c = 0; c < Nc; c++) { Pose3d p; p.fixed = fx[c]; p.q.x() = P[c*7]; p.q.y() = P[c*7+1]; p.q.z() = P[c*7+2]; p.q.w() = P[c*7+3]; for (int k = 0; k < 3; k++) p.p(k) = P[c*7+4+k]; poses[ids[c]] = p; }
    for (int l = 0; l < Np; l++) { Position3d q; for (int k = 0; k < 3; k++) q.p(k) = X[l*3+k]; points[pids[l]] = q; }
    std::vector<int> where(No);  // position in the mono / stereo vector
    for (int o = 0; o < No; o++) {
      if (kind[o]) { auto m = std::make_shared<StereoPointConstraint>(); m->id_pose = ids[oc[o]]; m->id_point = pids[op[o]]; m->id_camera = 0; m->inlier = true; for (int k = 0; k < 3; k++) m->keypoint(k) = uv3[o*3+k]; m->pixel_sigma = 0.8; where[o] = (int)stereo.size(); stereo.push_back(m); }
      else { auto m = std::make_shared<MonoPointConstraint>(); m->id_pose = ids[oc[o]]; m->id_point = pids[op[o]]; m->id_camera = 0; m->inlier = true; m->keypoint(0) = uv3[o*3]; m->keypoint(1) = uv3[o*3+1]; m->pixel_sigma = 0.8; where[o] = (int)mono.size(); mono.push_back(m); }
    }
    LocalmapOptimization(poses, points, cams, mono, stereo, cfg);
    std::vector<double> Po, Xo; std::vector<unsigned char> inl;
    for (auto& kv : poses) { Po.push_back(kv.second.q.x()); Po.push_back(kv.second.q.y()); Po.push_back(kv.second.q.z()); Po.push_back(kv.second.q.w()); for (int k = 0; k < 3; k++) Po.push_back(kv.second.p(k)); }
    for (auto& kv : points) for (int k = 0; k < 3; k++) Xo.push_back(kv.second.p(k));
    for (int o = 0; o < No; o++) inl.push_back(kind[o] ? stereo[where[o]]->inlier : mono[where[o]]->inlier);
    std::vector<int> status = {urmvo_adapter_last_status()};
    wr(out, Po); wr(out, Xo); wr(out, inl); wr(out, status);
  } else if (mode == "pose_stereo") {
    auto hdr = rd<int>(in, 1); int No = hdr[0];
    auto intr = rd<double>(in, 5); auto P = rd<double>(in, 7); auto uv3 = rd<double>(in, (size_t)No * 3); auto kind = rd<unsigned char>(in, No); auto X = rd<double>(in, (size_t)No * 3);
    cams.emplace_back(new Camera(intr[0], intr[1], intr[2], intr[3], STEREO, intr[4]));
    MapOfPoses poses; MapOfPoints3d points; VectorOfMonoPointConstraints mono; VectorOfStereoPointConstraints stereo;
    Pose3d p; p.q.x() = P[0]; p.q.y() = P[1]; p.q.z() = P[2]; p.q.w() = P[3]; for (int k = 0; k < 3; k++) p.p(k) = P[4+k]; poses[42] = p;
    std::vector<int> where(No);
    for (int o = 0; o < No; o++) { Position3d q; q.fixed = true; for (int k = 0; k < 3; k++) q.p(k) = X[o*3+k]; points[1000 + o] = q;
      if (kind[o]) { auto m = std::make_shared<StereoPointConstraint>(); m->id_pose = 42; m->id_point = 1000 + o; m->id_camera = 0; m->inlier = true; for (int k = 0; k < 3; k++) m->keypoint(k) = uv3[o*3+k]; m->pixel_sigma = 0.8; where[o] = (int)stereo.size(); stereo.push_back(m); }
      else { auto m = std::make_shared<MonoPointConstraint>(); m->id_pose = 42; m->id_point = 1000 + o; m->id_camera = 0; m->inlier = true; m->keypoint(0) = uv3[o*3]; m->keypoint(1) = uv3[o*3+1]; m->pixel_sigma = 0.8; where[o] = (int)mono.size(); mono.push_back(m); } }
    int n = FrameOptimization(poses, points, cams, mono, stereo, cfg);
    Pose3d& r = poses.begin()->second;
    std::vector<double> Po = {r.q.x(), r.q.y(), r.q.z(), r.q.w(), r.p(0), r.p(1), r.p(2)};
    std::vector<unsigned char> inl; for (int o = 0; o < No; o++) inl.push_back(kind[o] ? stereo[where[o]]->inlier : mono[where[o]]->inlier);
    std::vector<int> ni = {n};
    wr(out, Po); wr(out, inl); wr(out, ni);
  } else if (mode == "ba_multicam") {
    // several cameras in camera_list: kind_model[o] = stereo bit | id_camera << 1 (src/g2o_optimization.cc:86-89)
    auto hdr = rd<int>(in, 4); int Nc = hdr[0], Np = hdr[1], No = hdr[2], Nm = hdr[3];
    auto tab = rd<double>(in, (size_t)Nm * 5); auto ids = rd<int>(in, Nc); auto P = rd<double>(in, (size_t)Nc * 7); auto fx = rd<unsigned char>(in, Nc);
    auto pids = rd<int>(in, Np); auto X = rd<double>(in, (size_t)Np * 3);
    auto uv3 = rd<double>(in, (size_t)No * 3); auto km = rd<unsigned char>(in, No); auto oc = rd<int>(in, No); auto op = rd<int>(in, No);
    for (int m = 0; m < Nm; m++) cams.emplace_back(new Camera(tab[m*5], tab[m*5+1], tab[m*5+2], tab[m*5+3], STEREO, tab[m*5+4]));
    MapOfPoses poses; MapOfPoints3d points; VectorOfMonoPointConstraints mono; VectorOfStereoPointConstraints stereo;
    for (int c = 0; c < Nc; c++) { Pose3d p; p.fixed = fx[c]; p.q.x() = P[c*7]; p.q.y() = P[c*7+1]; p.q.z() = P[c*7+2]; p.q.w() = P[c*7+3]; for (int k = 0; k < 3; k++) p.p(k) = P[c*7+4+k]; poses[ids[c]] = p; }
    for (int l = 0; l < Np; l++) { Position3d q; for (int k = 0; k < 3; k++) q.p(k) = X[l*3+k]; points[pids[l]] = q; }
    std::vector<int> where(No);
    for (int o = 0; o < No; o++) {
      if (km[o] & 1) { auto m = std::make_shared<StereoPointConstraint>(); m->id_pose = ids[oc[o]]; m->id_point = pids[op[o]]; m->id_camera = km[o] >> 1; m->inlier = true; for (int k = 0; k < 3; k++) m->keypoint(k) = uv3[o*3+k]; m->pixel_sigma = 0.8; where[o] = (int)stereo.size(); stereo.push_back(m); }
      else { auto m = std::make_shared<MonoPointConstraint>(); m->id_pose = ids[oc[o]]; m->id_point = pids[op[o]]; m->id_camera = km[o] >> 1; m->inlier = true; m->keypoint(0) = uv3[o*3]; m->keypoint(1) = uv3[o*3+1]; m->pixel_sigma = 0.8; where[o] = (int)mono.size(); mono.push_back(m); }
    }
    LocalmapOptimization(poses, points, cams, mono, stereo, cfg);
    std::vector<double> Po, Xo; std::vector<unsigned char> inl;
    for (auto& kv : poses) { Po.push_back(kv.second.q.x()); Po.push_back(kv.second.q.y()); Po.push_back(kv.second.q.z()); Po.push_back(kv.second.q.w()); for (int k = 0; k < 3; k++) Po.push_back(kv.second.p(k)); }
    for (auto& kv : points) for (int k = 0; k < 3; k++) Xo.push_back(kv.second.p(k));
    for (int o = 0; o < No; o++) inl.push_back((km[o] & 1) ? stereo[where[o]]->inlier : mono[where[o]]->inlier);
    std::vector<int> status = {urmvo_adapter_last_status()};
    wr(out, Po); wr(out, Xo); wr(out, inl); wr(out, status);
  } else if (mode == "pose_multicam") {
    auto hdr = rd<int>(in, 2); int No = hdr[0], Nm = hdr[1];
    auto tab = rd<double>(in, (size_t)Nm * 5); auto P = rd<double>(in, 7); auto uv3 = rd<double>(in, (size_t)No * 3); auto km = rd<unsigned char>(in, No); auto X = rd<double>(in, (size_t)No * 3);
    for (int m = 0; m < Nm; m++) cams.emplace_back(new Camera(tab[m*5], tab[m*5+1], tab[m*5+2], tab[m*5+3], STEREO, tab[m*5+4]));
    MapOfPoses poses; MapOfPoints3d points; VectorOfMonoPointConstraints mono; VectorOfStereoPointConstraints stereo;
    Pose3d p; p.q.x() = P[0]; p.q.y() = P[1]; p.q.z() = P[2]; p.q.w() = P[3]; for (int k = 0; k < 3; k++) p.p(k) = P[4+k]; poses[42] = p;
    std::vector<int> where(No);
    for (int o = 0; o < No; o++) { Position3d q; q.fixed = true; for (int k = 0; k < 3; k++) q.p(k) = X[o*3+k]; points[1000 + o] = q;
      if (km[o] & 1) { auto m = std::make_shared<StereoPointConstraint>(); m->id_pose = 42; m->id_point = 1000 + o; m->id_camera = km[o] >> 1; m->inlier = true; for (int k = 0; k < 3; k++) m->keypoint(k) = uv3[o*3+k]; m->pixel_sigma = 0.8; where[o] = (int)stereo.size(); stereo.push_back(m); }
      else { auto m = std::make_shared<MonoPointConstraint>(); m->id_pose = 42; m->id_point = 1000 + o; m->id_camera = km[o] >> 1; m->inlier = true; m->keypoint(0) = uv3[o*3]; m->keypoint(1) = uv3[o*3+1]; m->pixel_sigma = 0.8; where[o] = (int)mono.size(); mono.push_back(m); } }
    int n = FrameOptimization(poses, points, cams, mono, stereo, cfg);
    Pose3d& r = poses.begin()->second;
    std::vector<double> Po = {r.q.x(), r.q.y(), r.q.z(), r.q.w(), r.p(0), r.p(1), r.p(2)};
    std::vector<unsigned char> inl; for (int o = 0; o < No; o++) inl.push_back((km[o] & 1) ? stereo[where[o]]->inlier : mono[where[o]]->inlier);
    std::vector<int> ni = {n, urmvo_adapter_last_status()};
    wr(out, Po); wr(out, inl); wr(out, ni);
  } else if (mode == "pose") {
    auto hdr = rd<int>(in, 1); int No = hdr[0];
    auto intr = rd<double>(in, 4); auto P = rd<double>(in, 7); auto uv = rd<double>(in, (size_t)No * 2); auto X = rd<double>(in, (size_t)No * 3);
    cams.emplace_back(new Camera(intr[0], intr[1], intr[2], intr[3]));
    MapOfPoses poses; MapOfPoints3d points; VectorOfMonoPointConstraints mono; VectorOfStereoPointConstraints stereo;
    Pose3d p; p.q.x() = P[0]; p.q.y() = P[1]; p.q.z() = P[2]; p.q.w() = P[3]; for (int k = 0; k < 3; k++) p.p(k) = P[4+k]; poses[42] = p;
    for (int o = 0; o < No; o++) { Position3d q; q.fixed = true; for (int k = 0; k < 3; k++) q.p(k) = X[o*3+k]; points[1000 + o] = q;
      auto m = std::make_shared<MonoPointConstraint>(); m->id_pose = 42; m->id_point = 1000 + o; m->id_camera = 0; m->inlier = true; m->keypoint(0) = uv[o*2]; m->keypoint(1) = uv[o*2+1]; m->pixel_sigma = 0.8; mono.push_back(m); }
    int n = FrameOptimization(poses, points, cams, mono, stereo, cfg);
    Pose3d& r = poses.begin()->second;
    std::vector<double> Po = {r.q.x(), r.q.y(), r.q.z(), r.q.w(), r.p(0), r.p(1), r.p(2)};
    std::vector<unsigned char> inl; for (auto& m : mono) inl.push_back(m->inlier);
    std::vector<int> ni = {n};
    wr(out, Po); wr(out, inl); wr(out, ni);
  } else if (mode == "tv") {
    auto hdr = rd<int>(in, 3); int n1 = hdr[0], n2 = hdr[1], its = hdr[2];
    auto K = rd<float>(in, 9); auto k1 = rd<float>(in, (size_t)n1 * 2); auto k2 = rd<float>(in, (size_t)n2 * 2); auto m = rd<int>(in, n1);
    Eigen::Matrix3f Km; for (int i = 0; i < 9; i++) Km.d[i] = K[i];
    std::vector<cv::KeyPoint> a(n1), b(n2);
    for (int i = 0; i < n1; i++) { a[i].pt.x = k1[i*2]; a[i].pt.y = k1[i*2+1]; }
    for (int i = 0; i < n2; i++) { b[i].pt.x = k2[i*2]; b[i].pt.y = k2[i*2+1]; }
    EpipolarGeometry eg(Km, 1.0f, its);
    Eigen::Matrix4f T21; std::vector<cv::Point3f> P3D; std::vector<bool> tri;
    bool ok = eg.reconstruct(a, b, std::vector<int>(m.begin(), m.end()), T21, P3D, tri);
    std::vector<int> okv = {ok ? 1 : 0}; std::vector<float> T(T21.d, T21.d + 16), P; std::vector<unsigned char> tv;
    for (auto& p : P3D) { P.push_back(p.x); P.push_back(p.y); P.push_back(p.z); }
    for (bool t : tri) tv.push_back(t);
    if (!ok) { P.assign((size_t)n1 * 3, 0.f); tv.assign(n1, 0); }
    wr(out, okv); wr(out, T); wr(out, P); wr(out, tv);
  } else if (mode == "pnp") {
    // n slots of the frame: valid[i] = 0 -> nullptr mappoint, 2 -> invalid mappoint, 1 -> usable
    auto hdr = rd<int>(in, 1); int n = hdr[0];
    auto intr = rd<double>(in, 4); auto X = rd<double>(in, (size_t)n * 3); auto uv = rd<double>(in, (size_t)n * 2); auto valid = rd<unsigned char>(in, n);
    CameraPtr cam(new Camera(intr[0], intr[1], intr[2], intr[3]));
    std::vector<Eigen::Vector2d> kps(n);
    std::vector<MappointPtr> mps(n);
    for (int i = 0; i < n; i++) {
      kps[i](0) = uv[i*2]; kps[i](1) = uv[i*2+1];
      Eigen::Vector3d p; for (int k = 0; k < 3; k++) p(k) = X[i*3+k];
      if (valid[i]) mps[i] = std::make_shared<Mappoint>(5000 + 3 * i, p, valid[i] == 1);
    }
    FramePtr frame(new Frame(cam, kps));
    Eigen::Matrix4d Twc; for (int i = 0; i < 4; i++) Twc(i, i) = 1.0;
    std::vector<int> inl;
    const int cnt = SolvePnPWithCV(frame, mps, Twc, inl);
    inl.resize(n, -1);
    std::vector<int> c = {cnt, urmvo_adapter_last_status()}; std::vector<double> T(Twc.d, Twc.d + 16);
    wr(out, c); wr(out, T); wr(out, inl);
  } else if (mode == "fm") {
    auto hdr = rd<int>(in, 1); int n = hdr[0];
    auto a = rd<float>(in, (size_t)n * 2); auto b = rd<float>(in, (size_t)n * 2);
    std::vector<cv::Point2f> p0(n), p1(n);
    for (int i = 0; i < n; i++) { p0[i].x = a[i*2]; p0[i].y = a[i*2+1]; p1[i].x = b[i*2]; p1[i].y = b[i*2+1]; }
    std::vector<unsigned char> inl(n, 7);
    const bool handled = FindFundamentalInliersGPU(p0, p1, inl);
    std::vector<int> h = {handled ? 1 : 0};
    wr(out, h); wr(out, inl);
  }
  fclose(in); fclose(out);
  return 0;
}
