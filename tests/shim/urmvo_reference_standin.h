// tests/shim/urmvo_reference_standin.h — NOT the reference's headers.
//
// The drop-in adapters (ur-mvo_b200/adapter/) include the reference's own headers by name
// ("g2o_optimization.h", "epipolar_geometry.h", "types.h").  Eigen, OpenCV and g2o are not installed
// in this image, so the adapter tests compile against this stand-in instead: it declares the
// interface the adapters implement (same names, same signatures, same fields — what
// include/g2o_optimization.h:13-21, include/epipolar_geometry.h:9-48 and include/types.h:18-104 of
// the reference expose) on top of the tiny Eigen / OpenCV stand-ins next to it.  In the reference's
// tree the real headers are used and this file plays no part.
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <opencv2/opencv.hpp>

#include <cstdlib>
#include <map>
#include <memory>
#include <vector>

// ---- types.h
struct Pose3d {
  bool fixed = false;
  Eigen::Vector3d p;
  Eigen::Quaterniond q;
};
struct Position3d {
  bool fixed = false;
  Eigen::Vector3d p;
};
struct MonoPointConstraint {
  int id_pose, id_point, id_camera;
  bool inlier;
  Eigen::Vector2d keypoint;
  double pixel_sigma;
};
struct StereoPointConstraint {
  int id_pose, id_point, id_camera;
  bool inlier;
  Eigen::Vector3d keypoint;
  double pixel_sigma;
};
using MapOfPoses = std::map<int, Pose3d, std::less<int>, Eigen::aligned_allocator<std::pair<const int, Pose3d>>>;
using MapOfPoints3d = std::map<int, Position3d, std::less<int>, Eigen::aligned_allocator<std::pair<const int, Position3d>>>;
using MonoPointConstraintPtr = std::shared_ptr<MonoPointConstraint>;
using StereoPointConstraintPtr = std::shared_ptr<StereoPointConstraint>;
using VectorOfMonoPointConstraints = std::vector<MonoPointConstraintPtr>;
using VectorOfStereoPointConstraints = std::vector<StereoPointConstraintPtr>;

// ---- what g2o_optimization.h pulls in from read_configs.h / camera.h (only what the optimiser reads)
struct OptimizationConfig {
  double mono_point, stereo_point, rate;
};
enum CameraType { MONO = 0, STEREO = 1 };
class Camera {
 public:
  Camera(double fx, double fy, double cx, double cy, CameraType type = MONO, double bf = 0)
      : f_{fx, fy}, c_{cx, cy}, type_(type), bf_(bf) {}
  CameraType GetCameraType() { return type_; }
  double BF() { return bf_; }
  double Fx() { return f_[0]; }
  double Fy() { return f_[1]; }
  double Cx() { return c_[0]; }
  double Cy() { return c_[1]; }

 private:
  double f_[2], c_[2];
  CameraType type_;
  double bf_;
};
using CameraPtr = std::shared_ptr<Camera>;

// ---- what g2o_optimization.h pulls in from frame.h / mappoint.h (only what SolvePnPWithCV reads:
// include/frame.h:57,83, include/mappoint.h:27,33,36)
class Mappoint {
 public:
  Mappoint(int id, const Eigen::Vector3d& p, bool valid = true) : id_(id), p_(p), valid_(valid) {}
  int GetId() { return id_; }
  bool IsValid() { return valid_; }
  Eigen::Vector3d& GetPosition() { return p_; }

 private:
  int id_;
  Eigen::Vector3d p_;
  bool valid_;
};
using MappointPtr = std::shared_ptr<Mappoint>;
class Frame {
 public:
  Frame(CameraPtr camera, std::vector<Eigen::Vector2d> keypoints) : camera_(camera), kps_(std::move(keypoints)) {}
  CameraPtr GetCamera() { return camera_; }
  bool GetKeypointPosition(size_t idx, Eigen::Vector2d& keypoint_pos) {
    if (idx >= kps_.size()) return false;
    keypoint_pos = kps_[idx];
    return true;
  }

 private:
  CameraPtr camera_;
  std::vector<Eigen::Vector2d> kps_;
};
using FramePtr = std::shared_ptr<Frame>;

// ---- g2o_optimization.h: the entry points the adapter defines
void LocalmapOptimization(MapOfPoses&, MapOfPoints3d&, std::vector<CameraPtr>&, VectorOfMonoPointConstraints&,
                          VectorOfStereoPointConstraints&, const OptimizationConfig&);
int FrameOptimization(MapOfPoses&, MapOfPoints3d&, std::vector<CameraPtr>&, VectorOfMonoPointConstraints&,
                      VectorOfStereoPointConstraints&, const OptimizationConfig&);
int SolvePnPWithCV(FramePtr frame, std::vector<MappointPtr>& mappoints, Eigen::Matrix4d& pose, std::vector<int>& inliers);

// ---- epipolar_geometry.h: public interface + the members the constructor initialises
class EpipolarGeometry {
 public:
  EpipolarGeometry(const Eigen::Matrix3f& K, float sigma = 1.0, int iterations = 200);
  bool reconstruct(const std::vector<cv::KeyPoint>&, const std::vector<cv::KeyPoint>&, const std::vector<int>,
                   Eigen::Matrix4f&, std::vector<cv::Point3f>&, std::vector<bool>&);
  class Random {
   public:
    static bool already_seeded;
    static void seed_rand(int);
    static void seed_rand_once(int);
    static int RandomInt(int, int);
  };

 private:
  Eigen::Matrix3f _K;
  float _Sigma, _Sigma2;
  int _MaxIterations;
};
