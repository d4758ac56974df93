"""Generates tests/golden/golden_pnp_r02.npz: outputs of the REAL cv2.solvePnPRansac (the OpenCV call of
reference src/g2o_optimization.cc:353-355, `cv::solvePnPRansac(..., false, 100, 20.0, 0.99, inliers)`) on seeded
frames.  Run in the build container (cv2 is importable there); the .npz travels with the repo, cv2 is not needed
to consume it.

    python tests/golden/make_golden_pnp.py

The golden is NOT regenerated when the specification of the EPnP hypotheses (oracle/pnp_oracle.cpp) changes: it
is the pin of that specification against the real library.
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
from urmvo_b200 import synth  # noqa: E402

# (points, outlier fraction, pixel noise) x seeds
SHAPES = [(8, 0.0, 0.5), (20, 0.1, 0.5), (50, 0.2, 0.7), (100, 0.3, 0.7), (200, 0.1, 0.7), (300, 0.2, 0.7), (500, 0.4, 1.0),
          (1000, 0.1, 0.7), (1000, 0.3, 0.7), (1000, 0.5, 0.7), (2000, 0.2, 0.5), (300, 0.6, 0.7)]
SEEDS_PER_SHAPE = 10


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    k = 0
    for si, (n, frac, sig) in enumerate(SHAPES):
        for j in range(SEEDS_PER_SHAPE):
            seed = 3000 + 100 * si + j
            p = synth.make_pnp(seed, n, frac, sig)
            K = np.array([[p["intr"][0], 0, p["intr"][2]], [0, p["intr"][1], p["intr"][3]], [0, 0, 1.0]])
            ok, rvec, tvec, inl = cv2.solvePnPRansac(p["obj"], p["img"], K, np.zeros(5), iterationsCount=100,
                                                     reprojectionError=20.0, confidence=0.99)
            mask = np.zeros(n, dtype=np.uint8)
            if ok and inl is not None:
                mask[inl.ravel()] = 1
            R = cv2.Rodrigues(rvec)[0] if ok else np.eye(3)
            out[f"case_{k}"] = np.array([seed, n, frac, sig])
            out[f"ok_{k}"] = np.array(int(ok))
            out[f"mask_{k}"] = mask
            out[f"R_{k}"] = R
            out[f"t_{k}"] = (tvec.ravel() if ok else np.zeros(3))
            k += 1
    out["n_cases"] = np.array(k)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_pnp_r02.npz"), **out)
    print(f"{k} cases, cv2 {cv2.__version__}")


if __name__ == "__main__":
    main()
