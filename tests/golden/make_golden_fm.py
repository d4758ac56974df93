"""Generates tests/golden/golden_fm_r01.npz: outputs of the REAL cv2.findFundamentalMat (the OpenCV
call of reference src/point_matching.cc:53) on seeded frame pairs.  Run in the build container
(cv2 is importable there); the .npz travels with the repo, cv2 is not needed to consume it.

    python tests/golden/make_golden_fm.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
from urmvo_b200 import synth  # noqa: E402

# (seed, matches, inlier fraction, pixel noise, rotation deg)
CASES = [
    (2001, 15, 0.9, 0.3, 2.0), (2002, 16, 0.8, 0.5, 3.0), (2003, 31, 0.7, 0.5, 5.0), (2004, 64, 0.6, 0.7, 5.0),
    (2005, 100, 0.5, 0.7, 4.0), (2006, 200, 0.45, 1.0, 6.0), (2007, 257, 0.8, 0.7, 1.0), (2008, 400, 0.35, 0.7, 5.0),
    (2009, 500, 0.9, 0.4, 8.0), (2010, 640, 0.7, 1.2, 5.0), (2011, 777, 0.6, 0.7, 2.5), (2012, 1000, 0.7, 0.7, 5.0),
    (2013, 1000, 0.3, 0.7, 5.0), (2014, 1000, 0.95, 0.2, 5.0), (2015, 1500, 0.5, 0.7, 5.0), (2016, 2048, 0.65, 0.9, 3.0),
]


def main():
    out = {"cv2_version": np.array(cv2.__version__), "n_cases": np.array(len(CASES))}
    for k, (seed, n, inl, sig, rot) in enumerate(CASES):
        p0, p1 = synth.make_fm(seed, n, inl, sig, rot)
        F, mask = cv2.findFundamentalMat(p0, p1, cv2.FM_RANSAC, 3, 0.99)
        assert F is not None and F.shape == (3, 3), (seed, None if F is None else F.shape)
        out[f"p0_{k}"] = p0
        out[f"p1_{k}"] = p1
        out[f"mask_{k}"] = mask.ravel().astype(np.uint8)
        out[f"F_{k}"] = F
        print(f"case {k}: seed {seed} N {n} inliers {int(mask.sum())}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_fm_r01.npz"), **out)


if __name__ == "__main__":
    main()
