// tests/shim/types.h — stand-in for the reference's include/types.h:18-104 (same names and fields;
// the reference's own header is used in its tree).
#pragma once
#include <map>
#include <memory>
#include <vector>
#include <Eigen/Core>
#include <Eigen/Geometry>
struct Pose3d { bool fixed = false; Eigen::Vector3d p; Eigen::Quaterniond q; };
typedef std::map<int, Pose3d, std::less<int>, Eigen::aligned_allocator<std::pair<const int, Pose3d>>> MapOfPoses;
struct Position3d { bool fixed = false; Eigen::Vector3d p; };
typedef std::map<int, Position3d, std::less<int>, Eigen::aligned_allocator<std::pair<const int, Position3d>>> MapOfPoints3d;
struct MonoPointConstraint { int id_pose, id_point, id_camera; bool inlier; Eigen::Vector2d keypoint; double pixel_sigma; };
typedef std::shared_ptr<MonoPointConstraint> MonoPointConstraintPtr;
typedef std::vector<MonoPointConstraintPtr> VectorOfMonoPointConstraints;
struct StereoPointConstraint { int id_pose, id_point, id_camera; bool inlier; Eigen::Vector3d keypoint; double pixel_sigma; };
typedef std::shared_ptr<StereoPointConstraint> StereoPointConstraintPtr;
typedef std::vector<StereoPointConstraintPtr> VectorOfStereoPointConstraints;
