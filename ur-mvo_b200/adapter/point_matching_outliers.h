// adapter/point_matching_outliers.h — drop-in for the outlier rejection inside
// PointMatching::MatchingPoints (reference src/point_matching.cc:50-60).
#pragma once
#include <vector>

#include <opencv2/opencv.hpp>

// Fills `inliers` (one byte per correspondence, 1 = keep) exactly as
//     cv::findFundamentalMat(points0, points1, cv::FM_RANSAC, 3, 0.99, inliers);
// does, on the GPU.  Returns true when the GPU path handled the call.  Returns false — leaving
// `inliers` untouched — for fewer than 15 correspondences (OpenCV's direct 7-point / LMedS branches):
// the caller then makes the original OpenCV call, see INTEGRATION.md.
bool FindFundamentalInliersGPU(const std::vector<cv::Point2f>& points0, const std::vector<cv::Point2f>& points1,
                               std::vector<unsigned char>& inliers);
