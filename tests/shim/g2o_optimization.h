// tests/shim/g2o_optimization.h — stand-in for the reference's include/g2o_optimization.h:13-21
// (identical declarations) plus the few types it pulls in from camera.h / read_configs.h.
#pragma once
#include <memory>
#include <vector>
#include "types.h"
struct OptimizationConfig { double mono_point; double stereo_point; double rate; };  // read_configs.h:39-43
enum CameraType { MONO = 0, STEREO = 1 };
class Camera {  // camera.h: only the getters the optimiser reads
 public:
  Camera(double fx, double fy, double cx, double cy) : _fx(fx), _fy(fy), _cx(cx), _cy(cy) {}
  CameraType GetCameraType() { return MONO; }
  double BF() { return 0; }
  double Fx() { return _fx; }
  double Fy() { return _fy; }
  double Cx() { return _cx; }
  double Cy() { return _cy; }
 private:
  double _fx, _fy, _cx, _cy;
};
typedef std::shared_ptr<Camera> CameraPtr;
void LocalmapOptimization(MapOfPoses& poses, MapOfPoints3d& points, std::vector<CameraPtr>& camera_list,
                          VectorOfMonoPointConstraints& mono_point_constraints,
                          VectorOfStereoPointConstraints& stereo_point_constraints, const OptimizationConfig& cfg);
int FrameOptimization(MapOfPoses& poses, MapOfPoints3d& points, std::vector<CameraPtr>& camera_list,
                      VectorOfMonoPointConstraints& mono_point_constraints,
                      VectorOfStereoPointConstraints& stereo_point_constraints, const OptimizationConfig& cfg);
