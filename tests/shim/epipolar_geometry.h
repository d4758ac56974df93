// tests/shim/epipolar_geometry.h — stand-in for the public part of the reference's
// include/epipolar_geometry.h:9-48 and the members its constructor initialises (:64-70).
#pragma once
#include <Eigen/Core>
#include <cstdlib>
#include <opencv2/opencv.hpp>
#include <vector>
class EpipolarGeometry {
 public:
  EpipolarGeometry(const Eigen::Matrix3f& k, float sigma = 1.0, int iterations = 200);
  bool reconstruct(const std::vector<cv::KeyPoint>& vKeys1, const std::vector<cv::KeyPoint>& vKeys2,
                   const std::vector<int> vMatches12, Eigen::Matrix4f& T21, std::vector<cv::Point3f>& vP3D,
                   std::vector<bool>& vbTriangulated);
  class Random {
   public:
    static bool already_seeded;
    static void seed_rand(int seed);
    static void seed_rand_once(int seed);
    static int RandomInt(int min, int max);
  };
 private:
  Eigen::Matrix3f _K;
  float _Sigma, _Sigma2;
  int _MaxIterations;
};
