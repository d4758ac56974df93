#!/usr/bin/env bash
# oracle/build_ref.sh — compiles the reference's OWN sources for the hot path (never copied: they are
# compiled where they lie under $REFERENCE, default /root/reference) into oracle/_ref/ref_driver, when the
# third-party dependencies exist.  In the image this repo was built in NONE of them does (no Eigen3, no
# OpenCV C++ headers, no g2o, no yaml-cpp, no network) — the script then says so and exits 3; DESIGN.md §2
# records that the BA / pose-only / two-view oracles are therefore "parity unpinned".
#
#   EIGEN3_INCLUDE=/usr/include/eigen3 OPENCV_PREFIX=/usr G2O_PREFIX=/usr/local YAMLCPP_PREFIX=/usr \
#       oracle/build_ref.sh && python oracle/make_ref_goldens.py
# make_ref_goldens.py then feeds the committed synthetic problems to ref_driver and writes
# tests/golden/golden_ref.npz, which tests/test_golden_oracle.py::test_oracle_against_reference_goldens
# picks up (skipped while the file does not exist).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REFERENCE="${REFERENCE:-/root/reference}"
EIGEN3_INCLUDE="${EIGEN3_INCLUDE:-/usr/include/eigen3}"
OPENCV_PREFIX="${OPENCV_PREFIX:-/usr}"
G2O_PREFIX="${G2O_PREFIX:-/usr/local}"
YAMLCPP_PREFIX="${YAMLCPP_PREFIX:-/usr}"
missing=()
[ -f "$EIGEN3_INCLUDE/Eigen/Dense" ] || missing+=("Eigen3 ($EIGEN3_INCLUDE/Eigen/Dense)")
[ -f "$OPENCV_PREFIX/include/opencv4/opencv2/opencv.hpp" ] || [ -f "$OPENCV_PREFIX/include/opencv2/opencv.hpp" ] || missing+=("OpenCV C++ headers under $OPENCV_PREFIX/include")
[ -f "$G2O_PREFIX/include/g2o/core/block_solver.h" ] || missing+=("g2o ($G2O_PREFIX/include/g2o)")
[ -f "$YAMLCPP_PREFIX/include/yaml-cpp/yaml.h" ] || missing+=("yaml-cpp ($YAMLCPP_PREFIX/include/yaml-cpp)")
[ -f "$REFERENCE/src/g2o_optimization.cc" ] || missing+=("the reference sources ($REFERENCE/src)")
if [ ${#missing[@]} -gt 0 ]; then
  echo "oracle/build_ref.sh: cannot build the reference here, missing:" >&2
  for m in "${missing[@]}"; do echo "   - $m" >&2; done
  exit 3
fi
OCV_INC="$OPENCV_PREFIX/include/opencv4"; [ -d "$OCV_INC" ] || OCV_INC="$OPENCV_PREFIX/include"
mkdir -p "$HERE/_ref"
# the files the path needs: the two translation units of the hot path plus what their headers pull in
# (camera / frame / mappoint / utils — SolvePnPWithCV in g2o_optimization.cc refers to Frame and Mappoint)
SRC=(g2o_optimization.cc epipolar_geometry.cc camera.cc frame.cc mappoint.cc utils.cc)
OBJ=()
for s in "${SRC[@]}"; do
  o="$HERE/_ref/${s%.*}.o"
  g++ -O2 -std=c++17 -fPIC -I"$REFERENCE/include" -I"$EIGEN3_INCLUDE" -I"$OCV_INC" -I"$G2O_PREFIX/include" \
      -I"$YAMLCPP_PREFIX/include" -c "$REFERENCE/src/$s" -o "$o"
  OBJ+=("$o")
done
g++ -O2 -std=c++17 -I"$REFERENCE/include" -I"$EIGEN3_INCLUDE" -I"$OCV_INC" -I"$G2O_PREFIX/include" -I"$YAMLCPP_PREFIX/include" \
    "$HERE/ref_driver.cc" "${OBJ[@]}" -o "$HERE/_ref/ref_driver" \
    -L"$G2O_PREFIX/lib" -lg2o_core -lg2o_stuff -lg2o_types_sba -lg2o_types_slam3d -lg2o_types_sim3 -lg2o_solver_eigen -lg2o_solver_dense \
    -L"$OPENCV_PREFIX/lib" -lopencv_core -lopencv_imgproc -lopencv_calib3d -lopencv_features2d -lopencv_highgui -lopencv_imgcodecs \
    -L"$YAMLCPP_PREFIX/lib" -lyaml-cpp -lpthread
echo "built $HERE/_ref/ref_driver"
