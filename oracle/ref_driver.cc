// oracle/ref_driver.cc — drives the UNMODIFIED reference (be2rlab/UR-MVO) through its own entry points and
// dumps the results, so that the CPU oracle can be pinned against real g2o / Eigen output the day a
// toolchain with g2o + Eigen3 + OpenCV + yaml-cpp exists (none of them is installed in this image:
// DESIGN.md §2).  Built by oracle/build_ref.sh together with the reference's own source files where they
// lie (never copied); outputs go to oracle/_ref/.  UNVERIFIED: it has never been compiled here.
//
//   ref_driver ba   camera.yaml in.bin out.bin     LocalmapOptimization  (src/g2o_optimization.cc:20-177)
//   ref_driver pose camera.yaml in.bin out.bin     FrameOptimization     (:179-321)
//   ref_driver tv   camera.yaml in.bin out.bin     EpipolarGeometry::reconstruct (src/epipolar_geometry.cc:18-98)
// The binary layouts are the ones tests/shim/adapter_driver.cc reads and writes (tests/test_adapter.py
// builds them), so oracle/make_ref_goldens.py can feed the same problems to both.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "epipolar_geometry.h"
#include "g2o_optimization.h"

template <class T> static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { perror("read"); exit(2); } return v; }
template <class T> static void wr(FILE* f, const std::vector<T>& v) { if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }

int main(int argc, char** argv) {
  if (argc < 5) return 1;
  const std::string mode = argv[1];
  // the reference's Camera reads fx, fy, cx, cy from its yaml (src/camera.cc:35-38): the goldens use
  // configs/camera_settings/aqua.yaml, the file the synthetic generator copies its intrinsics from
  CameraPtr camera = std::shared_ptr<Camera>(new Camera(argv[2], CameraType::MONO));
  std::vector<CameraPtr> cams{camera};
  FILE* in = fopen(argv[3], "rb"); FILE* out = fopen(argv[4], "wb");
  if (!in || !out) return 1;
  OptimizationConfig cfg;
  cfg.mono_point = 10.0; cfg.stereo_point = 75.0; cfg.rate = 0.5;  // configs/configs_aqua.yaml:40-48
  if (mode == "ba") {
    auto hdr = rd<int>(in, 3); int Nc = hdr[0], Np = hdr[1], No = hdr[2];
    rd<double>(in, 4);  // intrinsics of the problem file (must equal the yaml's)
    auto ids = rd<int>(in, Nc); auto P = rd<double>(in, (size_t)Nc * 7); auto fx = rd<unsigned char>(in, Nc);
    auto pids = rd<int>(in, Np); auto X = rd<double>(in, (size_t)Np * 3);
    auto uv = rd<double>(in, (size_t)No * 2); auto oc = rd<int>(in, No); auto op = rd<int>(in, No);
    MapOfPoses poses; MapOfPoints3d points; VectorOfMonoPointConstraints mono; VectorOfStereoPointConstraints stereo;
    for (int c = 0; c < Nc; c++) {
      Pose3d p; p.fixed = fx[c];
      p.q = Eigen::Quaterniond(P[c * 7 + 3], P[c * 7], P[c * 7 + 1], P[c * 7 + 2]);
      p.p = Eigen::Vector3d(P[c * 7 + 4], P[c * 7 + 5], P[c * 7 + 6]);
      poses.insert(std::pair<int, Pose3d>(ids[c], p));
    }
    for (int l = 0; l < Np; l++) {
      Position3d q; q.fixed = false; q.p = Eigen::Vector3d(X[l * 3], X[l * 3 + 1], X[l * 3 + 2]);
      points.insert(std::pair<int, Position3d>(pids[l], q));
    }
    for (int o = 0; o < No; o++) {
      MonoPointConstraintPtr m(new MonoPointConstraint());
      m->id_pose = ids[oc[o]]; m->id_point = pids[op[o]]; m->id_camera = 0; m->inlier = true;
      m->keypoint = Eigen::Vector2d(uv[o * 2], uv[o * 2 + 1]); m->pixel_sigma = 0.8;
      mono.push_back(m);
    }
    LocalmapOptimization(poses, points, cams, mono, stereo, cfg);
    std::vector<double> Po, Xo; std::vector<unsigned char> inl;
    for (auto& kv : poses) {
      Po.push_back(kv.second.q.x()); Po.push_back(kv.second.q.y()); Po.push_back(kv.second.q.z()); Po.push_back(kv.second.q.w());
      for (int k = 0; k < 3; k++) Po.push_back(kv.second.p(k));
    }
    for (auto& kv : points) for (int k = 0; k < 3; k++) Xo.push_back(kv.second.p(k));
    for (auto& m : mono) inl.push_back(m->inlier);
    wr(out, Po); wr(out, Xo); wr(out, inl);
  } else if (mode == "pose") {
    auto hdr = rd<int>(in, 1); int No = hdr[0];
    rd<double>(in, 4);
    auto P = rd<double>(in, 7); auto uv = rd<double>(in, (size_t)No * 2); auto X = rd<double>(in, (size_t)No * 3);
    MapOfPoses poses; MapOfPoints3d points; VectorOfMonoPointConstraints mono; VectorOfStereoPointConstraints stereo;
    Pose3d p; p.q = Eigen::Quaterniond(P[3], P[0], P[1], P[2]); p.p = Eigen::Vector3d(P[4], P[5], P[6]);
    poses.insert(std::pair<int, Pose3d>(42, p));
    for (int o = 0; o < No; o++) {
      Position3d q; q.fixed = true; q.p = Eigen::Vector3d(X[o * 3], X[o * 3 + 1], X[o * 3 + 2]);
      points.insert(std::pair<int, Position3d>(1000 + o, q));
      MonoPointConstraintPtr m(new MonoPointConstraint());
      m->id_pose = 42; m->id_point = 1000 + o; m->id_camera = 0; m->inlier = true;
      m->keypoint = Eigen::Vector2d(uv[o * 2], uv[o * 2 + 1]); m->pixel_sigma = 0.8;
      mono.push_back(m);
    }
    int n = FrameOptimization(poses, points, cams, mono, stereo, cfg);
    Pose3d& r = poses.begin()->second;
    std::vector<double> Po = {r.q.x(), r.q.y(), r.q.z(), r.q.w(), r.p(0), r.p(1), r.p(2)};
    std::vector<unsigned char> inl; for (auto& m : mono) inl.push_back(m->inlier);
    std::vector<int> ni = {n};
    wr(out, Po); wr(out, inl); wr(out, ni);
  } else if (mode == "tv") {
    auto hdr = rd<int>(in, 3); int n1 = hdr[0], n2 = hdr[1], its = hdr[2];
    auto K = rd<float>(in, 9); auto k1 = rd<float>(in, (size_t)n1 * 2); auto k2 = rd<float>(in, (size_t)n2 * 2); auto m = rd<int>(in, n1);
    Eigen::Matrix3f Km;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Km(i, j) = K[i * 3 + j];
    std::vector<cv::KeyPoint> a(n1), b(n2);
    for (int i = 0; i < n1; i++) { a[i].pt.x = k1[i * 2]; a[i].pt.y = k1[i * 2 + 1]; }
    for (int i = 0; i < n2; i++) { b[i].pt.x = k2[i * 2]; b[i].pt.y = k2[i * 2 + 1]; }
    EpipolarGeometry eg(Km, 1.0f, its);
    Eigen::Matrix4f T21 = Eigen::Matrix4f::Identity(); std::vector<cv::Point3f> P3D; std::vector<bool> tri;
    bool ok = eg.reconstruct(a, b, std::vector<int>(m.begin(), m.end()), T21, P3D, tri);
    std::vector<int> okv = {ok ? 1 : 0}; std::vector<float> T, P; std::vector<unsigned char> tv;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T.push_back(T21(i, j));
    P3D.resize(n1); tri.resize(n1);
    for (auto& p : P3D) { P.push_back(p.x); P.push_back(p.y); P.push_back(p.z); }
    for (bool t : tri) tv.push_back(t);
    if (!ok) { P.assign((size_t)n1 * 3, 0.f); tv.assign(n1, 0); }
    wr(out, okv); wr(out, T); wr(out, P); wr(out, tv);
  }
  fclose(in); fclose(out);
  return 0;
}
