"""Generates tests/golden/golden_fm_small_r02.npz: outputs of the REAL cv2.findFundamentalMat(FM_RANSAC, 3, 0.99)
(the OpenCV call of reference src/point_matching.cc:50) for the two match counts below 15 whose result is
reproducible: N == 7 (direct 7-point solution, every mask byte 1) and N == 14 (LMedS: the median is the smallest
error OUTSIDE the 7-point sample).  For 8 <= N <= 13 the median is the error of an exactly-fitted sample point
(~1e-27): the script also records how often the restatement in oracle/fm_oracle.cpp agrees with cv2 there
(`agree_8_13`, informational — it is rounding noise of the LAPACK SVD inside OpenCV).

    python tests/golden/make_golden_fm_small.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from urmvo_b200 import synth  # noqa: E402

# (seed, matches, inlier fraction, pixel noise, rotation deg)
CASES = [(5000 + 13 * k, n, inl, sig, rot) for k, (n, inl, sig, rot) in enumerate(
    [(7, 1.0, 0.3, 2.0), (7, 0.7, 0.5, 4.0), (7, 1.0, 0.0, 6.0), (7, 0.4, 1.0, 3.0), (7, 0.85, 0.7, 5.0), (7, 1.0, 1.5, 1.0),
     (14, 1.0, 0.3, 2.0), (14, 0.8, 0.5, 4.0), (14, 0.6, 0.7, 6.0), (14, 0.5, 1.0, 3.0), (14, 0.9, 0.0, 5.0),
     (14, 0.7, 0.7, 1.0), (14, 1.0, 1.2, 5.0), (14, 0.65, 0.4, 8.0), (14, 0.75, 0.9, 2.5), (14, 0.55, 0.6, 4.5)])]


def main():
    out = {"cv2_version": np.array(cv2.__version__), "n_cases": np.array(len(CASES))}
    for k, (seed, n, inl, sig, rot) in enumerate(CASES):
        p0, p1 = synth.make_fm(seed, n, inl, sig, rot)
        F, mask = cv2.findFundamentalMat(p0, p1, cv2.FM_RANSAC, 3, 0.99)
        out[f"p0_{k}"] = p0
        out[f"p1_{k}"] = p1
        out[f"mask_{k}"] = mask.ravel().astype(np.uint8)
        out[f"found_{k}"] = np.array(0 if F is None else 1)
        Fp = np.zeros((9, 3))  # N == 7: up to three stacked solutions (their order depends on the null-space basis)
        if F is not None:
            Fp[:F.shape[0]] = F
        out[f"F_{k}"] = Fp
        print(f"case {k}: seed {seed} N {n} found {F is not None} inliers {int(mask.sum())}")
    import pyoracle
    agree = np.zeros((6, 2), dtype=np.int32)
    for n in range(8, 14):
        for seed in range(60):
            p0, p1 = synth.make_fm(3000 + seed * 17 + n, n, [0.6, 0.8, 1.0][seed % 3], 0.5, 4.0)
            _, mask = cv2.findFundamentalMat(p0, p1, cv2.FM_RANSAC, 3, 0.99)
            agree[n - 8, 0] += int(np.array_equal(mask.ravel().astype(np.uint8), pyoracle.find_fundamental(p0, p1)["mask"]))
            agree[n - 8, 1] += 1
    out["agree_8_13"] = agree
    print("agreement 8..13 (informational):", agree[:, 0].tolist(), "of", agree[:, 1].tolist())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_fm_small_r02.npz"), **out)


if __name__ == "__main__":
    main()
