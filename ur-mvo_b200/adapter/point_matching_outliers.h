// adapter/point_matching_outliers.h — drop-in for the outlier rejection inside
// PointMatching::MatchingPoints (reference src/point_matching.cc:50-60).
#pragma once
#include <vector>

#include <opencv2/opencv.hpp>

// Fills `inliers` (one byte per correspondence, 1 = keep) exactly as
//     cv::findFundamentalMat(points0, points1, cv::FM_RANSAC, 3, 0.99, inliers);
// does, on the GPU.  Returns true when the GPU path handled the call.  Returns false — leaving
// `inliers` untouched — in two cases the caller can tell apart with FindFundamentalInliersStatus():
//   URMVO_ERR_UNSUPPORTED  fewer than 7 correspondences (OpenCV returns an empty matrix and does not create the mask;
//                          the reference then reads inliers[i] of an empty vector): keep the original call;
//   any other status       the GPU call failed (no B200, a CUDA error such as a failed allocation): the message is
//                          in urmvo_last_error(); nothing aborts — keep the OpenCV call for this frame or stop.
bool FindFundamentalInliersGPU(const std::vector<cv::Point2f>& points0, const std::vector<cv::Point2f>& points1,
                               std::vector<unsigned char>& inliers);
// urmvo_status of the last FindFundamentalInliersGPU call on this thread (0 after a handled call).
int FindFundamentalInliersStatus();
