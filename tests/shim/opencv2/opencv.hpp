// tests/shim/opencv2/opencv.hpp — NOT OpenCV: the two value types the adapter's signature names.
#pragma once
namespace cv {
struct Point2f { float x = 0, y = 0; };
struct Point3f { float x = 0, y = 0, z = 0; Point3f() = default; Point3f(float a, float b, float c) : x(a), y(b), z(c) {} };
struct KeyPoint { Point2f pt; };
}  // namespace cv
