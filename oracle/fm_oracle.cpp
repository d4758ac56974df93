// oracle/fm_oracle.cpp — CPU restatement of the per-frame outlier rejection of
// PointMatching::MatchingPoints (reference src/point_matching.cc:44-58):
//     cv::findFundamentalMat(points0, points1, cv::FM_RANSAC, 3, 0.99, inliers);
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// PARITY PINNED: the algorithm lives in OpenCV (calib3d: fundam.cpp / ptsetreg.cpp), which the
// reference links but does not vendor.  The Python build of the same library (cv2 4.13) is importable
// in the build container, so tests/golden/make_golden_fm.py runs the REAL cv2.findFundamentalMat on
// seeded inputs and commits inputs + inlier masks + F as tests/golden/golden_fm_r01.npz;
// tests/test_golden_fm.py requires this restatement to reproduce every mask bit-for-bit.
//
// Restated pieces (OpenCV 4.x):
//   cv::RNG (multiply-with-carry, state 0xffffffffffffffff, uniform(a,b) = a + next() % (b-a))
//   RANSACPointSetRegistrator::run / getSubset (7 distinct indices, <= 10000 attempts, checkSubset)
//   FMEstimatorCallback::checkSubset -> haveCollinearPoints (last point against all earlier pairs)
//   run7Point (isotropic normalisation, 2-D null space of the 7x9 system, cubic det = 0, <= 3 models,
//              de-normalisation, F33 = 1)
//   cv::solveCubic
//   FMEstimatorCallback::computeError (max of the two squared point-line distances, float result)
//   findInliers (err <= (float)(thr*thr)), best = strictly more inliers, RANSACUpdateNumIters
// Below 15 points cv::findFundamentalMat leaves the RANSAC branch (fundam.cpp): N == 7 runs the 7-point solver once
// and sets every mask byte to 1; 8 <= N <= 14 runs LMeDSPointSetRegistrator::run (outlier ratio 0.45 -> 300
// iterations, median = the count/2-th smallest float error, sigma = 2.5 * 1.4826 * (1 + 5 / (N - 7)) * sqrt(median)).
// urmvo_oracle_find_fundamental restates the whole dispatch.  For 8 <= N <= 13 the count/2-th smallest error is
// one of the seven exactly-fitted sample points (~1e-27): the winner is decided by rounding noise of the solver —
// the restatement is the same algorithm, but no two builds of it (OpenCV's own included) agree bit for bit there;
// N == 7 and N == 14 are pinned against the real cv2 (tests/golden/make_golden_fm_small.py).
// The null space comes from a Householder QR of the transposed 7x9 system instead of the SVD OpenCV
// calls (LAPACK): any basis of the null space gives the same <= 3 fundamental matrices.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "oracle.h"

namespace {

struct CvRng {
  uint64_t state = 0xffffffffffffffffull;
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a + (int)(next() % (unsigned)(b - a)); }
};

bool have_collinear(const float* m, const int* idx, int count) {
  const int i = count - 1;
  const float* pi = m + 2 * idx[i];
  for (int j = 0; j < i; j++) {
    const double dx1 = m[2 * idx[j]] - pi[0], dy1 = m[2 * idx[j] + 1] - pi[1];
    for (int k = 0; k < j; k++) {
      const double dx2 = m[2 * idx[k]] - pi[0], dy2 = m[2 * idx[k] + 1] - pi[1];
      if (std::fabs(dx2 * dy1 - dy2 * dx1) <=
          FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2)))
        return true;
    }
  }
  return false;
}

bool get_subset(const float* m1, const float* m2, int count, CvRng& rng, int max_attempts, int* idx) {
  for (int it = 0; it < max_attempts; it++) {
    for (int i = 0; i < 7; i++) {
      int v = rng.uniform(0, count);
      while (std::find(idx, idx + i, v) != idx + i) v = rng.uniform(0, count);
      idx[i] = v;
    }
    if (!have_collinear(m1, idx, 7) && !have_collinear(m2, idx, 7)) return true;
  }
  return false;
}

// cv::solveCubic: c[0] x^3 + c[1] x^2 + c[2] x + c[3] = 0.  Returns the number of roots (-1: all x).
int solve_cubic(const double* c, double* r) {
  double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
  int n = 0;
  double x0 = 0, x1 = 0, x2 = 0;
  if (a0 == 0) {
    if (a1 == 0) {
      if (a2 == 0) {
        n = a3 == 0 ? -1 : 0;
      } else {
        x0 = -a3 / a2;
        n = 1;
      }
    } else {
      double d = a2 * a2 - 4 * a1 * a3;
      if (d >= 0) {
        d = std::sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (std::fabs(q1) > std::fabs(q2)) {
          x0 = q1 / a1;
          x1 = a3 / q1;
        } else {
          x0 = q2 / a1;
          x1 = a3 / q2;
        }
        n = d > 0 ? 2 : 1;
      }
    }
  } else {
    a0 = 1. / a0;
    a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d > 0) {
      const double theta = std::acos(R / std::sqrt(Qcubed));
      const double sqrtQ = std::sqrt(Q);
      const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
      x0 = t0 * std::cos(t1) - t2;
      x1 = t0 * std::cos(t1 + (2. * M_PI / 3)) - t2;
      x2 = t0 * std::cos(t1 + (4. * M_PI / 3)) - t2;
      n = 3;
    } else if (d == 0) {
      if (R >= 0) {
        x0 = -2 * std::pow(R, 1. / 3) - a1 / 3;
        x1 = std::pow(R, 1. / 3) - a1 / 3;
      } else {
        x0 = 2 * std::pow(-R, 1. / 3) - a1 / 3;
        x1 = -std::pow(-R, 1. / 3) - a1 / 3;
      }
      x2 = 0;
      n = x0 == x1 ? 1 : 2;
      x1 = x0 == x1 ? 0 : x1;
    } else {
      double e;
      d = std::sqrt(-d);
      e = std::pow(d + std::fabs(R), 1. / 3);
      if (R > 0) e = -e;
      x0 = (e + Q / e) - a1 * (1. / 3);
      n = 1;
    }
  }
  r[0] = x0; r[1] = x1; r[2] = x2;
  return n;
}

// Sum of 16 values in the order of a 16-lane xor-butterfly (what the CUDA kernel's shuffles do):
// pairwise tree ((v0+v1)+(v2+v3))+... ; every lane of the butterfly ends with this same value.
double tree16(const double* v) {
  double a[16];
  for (int i = 0; i < 16; i++) a[i] = v[i];
  for (int w = 1; w < 16; w <<= 1)
    for (int i = 0; i < 16; i += 2 * w) a[i] = a[i] + a[i + w];
  return a[0];
}

// 2-D null space of the 7x9 system A (rows = equations): Householder QR of M = A^T (9x7),
// Q = H_0 H_1 ... H_6, and the last two columns of Q span null(A).  f1 = Q e_8, f2 = Q e_7.
// Spec shared with the CUDA kernel (csrc/fm_kernels.cu), operation by operation: reflector k uses
// x = M[k.., k], sigma = tree16(x^2), norm = sqrt(sigma), alpha = -sign(x_k) norm, v = x - alpha e_k,
// beta = 1 / (norm (norm + |x_k|)) (0 when norm == 0), M[:, j] -= (beta * tree16(v . M[:, j])) * v.
// OpenCV takes the last two right singular vectors of an SVD instead; any basis of the null space
// gives the same <= 3 fundamental matrices (tests/test_golden_fm.py pins the result against cv2).
void null_space_7x9(double* A, double* f1, double* f2) {
  double M[16][7];  // row r of A^T, rows 9..15 are the idle lanes (zero)
  for (int r = 0; r < 16; r++)
    for (int k = 0; k < 7; k++) M[r][k] = r < 9 ? A[k * 9 + r] : 0.0;
  double V[7][16], beta[7];
  for (int k = 0; k < 7; k++) {
    double x[16], sq[16];
    for (int r = 0; r < 16; r++) { x[r] = r >= k ? M[r][k] : 0.0; sq[r] = x[r] * x[r]; }
    const double sigma = tree16(sq);
    const double xkk = M[k][k];
    const double norm = std::sqrt(sigma);
    const double alpha = xkk >= 0.0 ? -norm : norm;
    for (int r = 0; r < 16; r++) V[k][r] = r == k ? x[r] - alpha : x[r];
    beta[k] = norm > 0.0 ? 1.0 / (norm * (norm + std::fabs(xkk))) : 0.0;
    for (int j = k + 1; j < 7; j++) {
      double pr[16];
      for (int r = 0; r < 16; r++) pr[r] = V[k][r] * M[r][j];
      const double w = beta[k] * tree16(pr);
      for (int r = 0; r < 16; r++) M[r][j] = M[r][j] - w * V[k][r];
    }
  }
  double y1[16], y2[16];
  for (int r = 0; r < 16; r++) { y1[r] = r == 8 ? 1.0 : 0.0; y2[r] = r == 7 ? 1.0 : 0.0; }
  for (int k = 6; k >= 0; k--) {
    double p1[16], p2[16];
    for (int r = 0; r < 16; r++) { p1[r] = V[k][r] * y1[r]; p2[r] = V[k][r] * y2[r]; }
    const double w1 = beta[k] * tree16(p1), w2 = beta[k] * tree16(p2);
    for (int r = 0; r < 16; r++) { y1[r] = y1[r] - w1 * V[k][r]; y2[r] = y2[r] - w2 * V[k][r]; }
  }
  for (int r = 0; r < 9; r++) { f1[r] = y1[r]; f2[r] = y2[r]; }
}

// run7Point on the 7 selected correspondences; F: up to 3 row-major 3x3 matrices.
int run7point(const float* m1, const float* m2, const int* idx, double* F) {
  double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
  for (int i = 0; i < 7; i++) {
    c1x += m1[2 * idx[i]]; c1y += m1[2 * idx[i] + 1];
    c2x += m2[2 * idx[i]]; c2y += m2[2 * idx[i] + 1];
  }
  const double t = 1. / 7;
  c1x *= t; c1y *= t; c2x *= t; c2y *= t;
  double scale1 = 0, scale2 = 0;
  for (int i = 0; i < 7; i++) {
    const double dx1 = m1[2 * idx[i]] - c1x, dy1 = m1[2 * idx[i] + 1] - c1y;
    const double dx2 = m2[2 * idx[i]] - c2x, dy2 = m2[2 * idx[i] + 1] - c2y;
    scale1 += std::sqrt(dx1 * dx1 + dy1 * dy1);
    scale2 += std::sqrt(dx2 * dx2 + dy2 * dy2);
  }
  scale1 *= t; scale2 *= t;
  if (scale1 < FLT_EPSILON || scale2 < FLT_EPSILON) return 0;
  scale1 = std::sqrt(2.) / scale1;
  scale2 = std::sqrt(2.) / scale2;
  double A[63];
  for (int i = 0; i < 7; i++) {
    const double x0 = (m1[2 * idx[i]] - c1x) * scale1, y0 = (m1[2 * idx[i] + 1] - c1y) * scale1;
    const double x1 = (m2[2 * idx[i]] - c2x) * scale2, y1 = (m2[2 * idx[i] + 1] - c2y) * scale2;
    double* a = A + i * 9;
    a[0] = x1 * x0; a[1] = x1 * y0; a[2] = x1;
    a[3] = y1 * x0; a[4] = y1 * y0; a[5] = y1;
    a[6] = x0; a[7] = y0; a[8] = 1;
  }
  double f1[9], f2[9];
  null_space_7x9(A, f1, f2);
  for (int i = 0; i < 9; i++) f1[i] -= f2[i];
  double c[4], r[3];
  double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
  c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
  c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
         f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
         f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
         f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
         f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
  const int n = solve_cubic(c, r);
  if (n < 1 || n > 3) return n < 0 ? 0 : n;
  const double T1[9] = {scale1, 0, -scale1 * c1x, 0, scale1, -scale1 * c1y, 0, 0, 1};
  const double T2[9] = {scale2, 0, -scale2 * c2x, 0, scale2, -scale2 * c2y, 0, 0, 1};
  for (int k = 0; k < n; k++) {
    double* fm = F + 9 * k;
    double lambda = r[k], mu = 1.;
    const double s = f1[8] * r[k] + f2[8];
    double Fn[9];
    if (std::fabs(s) > DBL_EPSILON) {
      mu = 1. / s;
      lambda *= mu;
      Fn[8] = 1.;
    } else {
      Fn[8] = 0.;
    }
    for (int i = 0; i < 8; i++) Fn[i] = f1[i] * lambda + f2[i] * mu;
    // F = T2^T * Fn * T1
    double tmp[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double a = 0;
        for (int l = 0; l < 3; l++) a += T2[l * 3 + i] * Fn[l * 3 + j];
        tmp[i * 3 + j] = a;
      }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double a = 0;
        for (int l = 0; l < 3; l++) a += tmp[i * 3 + l] * T1[l * 3 + j];
        fm[i * 3 + j] = a;
      }
    if (std::fabs(fm[8]) > FLT_EPSILON) {
      const double inv = 1. / fm[8];
      for (int i = 0; i < 9; i++) fm[i] *= inv;
    }
  }
  return n;
}

inline float sampson_max(const double* F, const float* a, const float* b) {
  const double x1 = a[0], y1 = a[1], x2 = b[0], y2 = b[1];
  double A = F[0] * x1 + F[1] * y1 + F[2];
  double B = F[3] * x1 + F[4] * y1 + F[5];
  double C = F[6] * x1 + F[7] * y1 + F[8];
  const double s2 = 1. / (A * A + B * B);
  const double d2 = x2 * A + y2 * B + C;
  A = F[0] * x2 + F[3] * y2 + F[6];
  B = F[1] * x2 + F[4] * y2 + F[7];
  C = F[2] * x2 + F[5] * y2 + F[8];
  const double s1 = 1. / (A * A + B * B);
  const double d1 = x1 * A + y1 * B + C;
  return (float)std::max(d1 * d1 * s1, d2 * d2 * s2);
}

int update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = std::min(std::max(p, 0.), 1.);
  ep = std::min(std::max(ep, 0.), 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::nearbyint(num / denom);
}

}  // namespace

extern "C" int urmvo_oracle_fm_subsets(int N, const float* p0, const float* p1, int max_iters, int32_t* idx) {
  CvRng rng;
  int n = 0;
  for (; n < max_iters; n++)
    if (!get_subset(p0, p1, N, rng, 10000, idx + 7 * n)) break;
  return n;
}

extern "C" int urmvo_oracle_fm_run7(const float* p0, const float* p1, double* F27) {
  const int idx[7] = {0, 1, 2, 3, 4, 5, 6};
  return run7point(p0, p1, idx, F27);
}

extern "C" void urmvo_oracle_fm_errors(int N, const float* p0, const float* p1, const double* F, float* err) {
  for (int i = 0; i < N; i++) err[i] = sampson_max(F, p0 + 2 * i, p1 + 2 * i);
}

extern "C" int urmvo_oracle_fm_ransac(int N, const float* p0, const float* p1, double thresh, double confidence,
                                      int max_iters, uint8_t* mask, double* F9, int32_t* stats3) {
  if (N < 15) return -1;  // OpenCV: < 7 nothing, 7 direct, 8..14 LMedS — not this path
  if (thresh <= 0) thresh = 3;
  if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
  CvRng rng;
  int niters = std::max(max_iters, 1), max_good = 0, iter = 0, models = 0;
  const float t = (float)(thresh * thresh);
  double best[9] = {0};
  std::memset(mask, 0, N);
  uint8_t* cur = new uint8_t[N];
  for (iter = 0; iter < niters; iter++) {
    int idx[7];
    if (!get_subset(p0, p1, N, rng, 10000, idx)) {
      if (iter == 0) { delete[] cur; return 0; }
      break;
    }
    double F[27];
    const int nm = run7point(p0, p1, idx, F);
    if (nm <= 0) continue;
    for (int k = 0; k < nm; k++) {
      models++;
      int good = 0;
      for (int i = 0; i < N; i++) {
        const int f = sampson_max(F + 9 * k, p0 + 2 * i, p1 + 2 * i) <= t;
        cur[i] = (uint8_t)f;
        good += f;
      }
      if (good > std::max(max_good, 6)) {
        std::memcpy(mask, cur, N);
        std::memcpy(best, F + 9 * k, sizeof(best));
        max_good = good;
        niters = update_num_iters(confidence, (double)(N - good) / N, 7, niters);
      }
    }
  }
  delete[] cur;
  if (F9) std::memcpy(F9, best, sizeof(best));
  if (stats3) { stats3[0] = iter; stats3[1] = max_good; stats3[2] = models; }
  return max_good > 0 ? 1 : 0;
}

// LMeDSPointSetRegistrator::run (ptsetreg.cpp) with FMEstimatorCallback, 8 <= N <= 14 in findFundamentalMat.
static int fm_lmeds(int N, const float* p0, const float* p1, double confidence, int max_iters, uint8_t* mask,
                    double* F9, int32_t* stats3) {
  CvRng rng;
  int niters = std::max(update_num_iters(confidence, 0.45, 7, max_iters), 3), iter = 0, models = 0;
  double min_median = DBL_MAX, best[9] = {0};
  std::memset(mask, 0, N);
  float err[16];
  for (iter = 0; iter < niters; iter++) {
    int idx[7];
    if (!get_subset(p0, p1, N, rng, 10000, idx)) {
      if (iter == 0) return 0;
      break;
    }
    double F[27];
    const int nm = run7point(p0, p1, idx, F);
    if (nm <= 0) continue;
    for (int k = 0; k < nm; k++) {
      models++;
      for (int i = 0; i < N; i++) err[i] = sampson_max(F + 9 * k, p0 + 2 * i, p1 + 2 * i);
      int32_t bits[16];  // std::nth_element(errf.ptr<int>(), ... + count / 2, ...): ordered as integers
      std::memcpy(bits, err, sizeof(float) * N);
      std::sort(bits, bits + N);
      float med;
      std::memcpy(&med, &bits[N / 2], sizeof(float));
      const double median = med;
      if (median < min_median) { min_median = median; std::memcpy(best, F + 9 * k, sizeof(best)); }
    }
  }
  int count = 0;
  if (min_median < DBL_MAX) {
    double sigma = 2.5 * 1.4826 * (1 + 5. / (N - 7)) * std::sqrt(min_median);
    sigma = std::max(sigma, 0.001);
    const float t = (float)(sigma * sigma);
    for (int i = 0; i < N; i++) {
      mask[i] = sampson_max(best, p0 + 2 * i, p1 + 2 * i) <= t ? 1 : 0;
      count += mask[i];
    }
  }
  if (F9) std::memcpy(F9, best, sizeof(best));
  if (stats3) { stats3[0] = iter; stats3[1] = count; stats3[2] = models; }
  return (min_median < DBL_MAX && count >= 7) ? 1 : 0;
}

// cv::findFundamentalMat(points0, points1, FM_RANSAC, thresh, confidence, mask) for any N (fundam.cpp):
// returns 1 if OpenCV returns a matrix, 0 if it returns an empty one (the mask is filled as OpenCV leaves it),
// -1 for N < 7 (OpenCV returns before it creates the mask).
extern "C" int urmvo_oracle_find_fundamental(int N, const float* p0, const float* p1, double thresh,
                                             double confidence, int max_iters, uint8_t* mask, double* F9,
                                             int32_t* stats3) {
  if (N < 7) return -1;
  if (N >= 15) return urmvo_oracle_fm_ransac(N, p0, p1, thresh, confidence, max_iters, mask, F9, stats3);
  if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
  if (N == 7) {  // direct 7-point solution, every point flagged (mask.setTo(1))
    const int idx[7] = {0, 1, 2, 3, 4, 5, 6};
    double F[27];
    const int nm = run7point(p0, p1, idx, F);
    std::memset(mask, 1, 7);
    if (F9) { std::memset(F9, 0, 9 * sizeof(double)); if (nm > 0) std::memcpy(F9, F, 9 * sizeof(double)); }
    if (stats3) { stats3[0] = 1; stats3[1] = 7; stats3[2] = std::max(nm, 0); }
    return nm > 0 ? 1 : 0;
  }
  return fm_lmeds(N, p0, p1, confidence, std::max(max_iters, 1), mask, F9, stats3);
}

