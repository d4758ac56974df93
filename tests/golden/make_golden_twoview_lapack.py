"""Generates tests/golden/golden_twoview_lapack.npz: EpipolarGeometry::reconstruct over 210 seeded scenes
computed by the float64 numpy.linalg.svd (LAPACK) pipeline of tests/lapack_twoview.py — an implementation
that is INDEPENDENT of the repo's fp32 specification (oracle / CUDA kernels).  The file pins how far that
specification may drift from an Eigen-JacobiSVD-class implementation; it is NOT to be regenerated when the
specification (Jacobi / Householder details, summation orders) changes — only if the scene list changes.
    python tests/golden/make_golden_twoview_lapack.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from urmvo_b200 import synth  # noqa: E402
import lapack_twoview as L  # noqa: E402

N_HYP = 200  # the reference's iteration count (include/epipolar_geometry.h:20)


def scene_list():
    """(seed, n_keys, inlier_frac, planar, rot_deg, t) — general, planar and low-inlier scenes."""
    out = []
    for s in range(120):
        r = np.random.default_rng(s)
        out.append((5000 + s, int(r.integers(300, 1001)), float(r.uniform(0.5, 0.9)), 0, float(r.uniform(1, 8)),
                    float(r.uniform(0.1, 0.5)), float(r.uniform(-0.05, 0.05)), float(r.uniform(-0.1, 0.1))))
    for s in range(50):
        r = np.random.default_rng(1000 + s)
        out.append((6000 + s, int(r.integers(300, 801)), float(r.uniform(0.6, 0.9)), 1, float(r.uniform(2, 8)),
                    float(r.uniform(0.2, 0.5)), 0.02, 0.05))
    for s in range(40):
        r = np.random.default_rng(2000 + s)
        out.append((7000 + s, int(r.integers(400, 1001)), float(r.uniform(0.25, 0.45)), 0, 5.0, 0.3, 0.02, 0.05))
    return np.array(out, dtype=np.float64)


def make_scene(row):
    seed, n_keys, frac, planar, rot, tx, ty, tz = row
    tv = synth.make_two_view(int(seed), n_keys=int(n_keys), inlier_frac=float(frac), planar=bool(planar), rot_deg=float(rot),
                             t=(float(tx), float(ty), float(tz)))
    tv["sets"] = synth.draw_sets(int((tv["matches12"] >= 0).sum()), N_HYP, int(seed))
    return tv


if __name__ == "__main__":
    scenes = scene_list()
    summ = np.zeros((len(scenes), 6), dtype=np.int32)   # used_H, ok, best_F, best_H, popcount(mask_F), popcount(mask_H)
    T21 = np.zeros((len(scenes), 4, 4))
    masks = []
    for i, row in enumerate(scenes):
        r = L.reconstruct(make_scene(row))
        summ[i] = (r["used_H"], int(r["ok"]), r["best_F"], r["best_H"], int(r["mask_F"].sum()), int(r["mask_H"].sum()))
        if r["ok"]:
            T21[i] = r["T21"]
        win = r["mask_H"] if r["used_H"] == 1 else r["mask_F"]
        masks.append(np.packbits(win.astype(np.uint8)))
    off = np.r_[0, np.cumsum([len(m) for m in masks])].astype(np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_twoview_lapack.npz"), scenes=scenes, summary=summ, T21=T21,
                        mask_bits=np.concatenate(masks), mask_off=off, numpy_version=np.__version__)
    print("scenes", len(scenes), "ok", int(summ[:, 1].sum()), "homography chosen", int((summ[:, 0] == 1).sum()))
