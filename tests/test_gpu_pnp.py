"""GPU (-m gpu): SolvePnPWithCV through the C ABI (urmvo_pnp_ransac[_batch]) against (1) the CPU oracle — identical
inlier masks, iteration counts and model counts, refined pose to 1e-9 — and (2) the committed outputs of the real
cv2.solvePnPRansac (same bar as the oracle's own pin in tests/test_golden_pnp.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from urmvo_b200 import synth
import urmvo_b200 as U

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, "golden_pnp_r02.npz"))
N_CASES = int(G["n_cases"])


def _problem(k):
    seed, n, frac, sig = G[f"case_{k}"]
    return synth.make_pnp(int(seed), int(n), float(frac), float(sig))


@pytest.mark.parametrize("seed,n,frac", [(4001, 8, 0.0), (4002, 33, 0.2), (4003, 300, 0.3), (4004, 1000, 0.1),
                                         (4005, 1000, 0.55), (4006, 4000, 0.3), (4007, 64, 0.7)])
def test_gpu_matches_oracle(ctx, oracle, seed, n, frac):
    p = synth.make_pnp(seed, n, frac)
    g = ctx.pnp_ransac(p["obj"], p["img"], p["intr"])
    o = oracle.pnp_ransac(p["obj"], p["img"], p["intr"])
    assert g["found"] == o["found"]
    assert np.array_equal(g["mask"], o["mask"]), f"{(g['mask'] != o['mask']).sum()} flags differ"
    assert (g["iters"], g["n_inliers"], g["models"]) == (o["iters"], o["n_inliers"], o["models"])
    if o["found"]:
        assert np.abs(g["R"] - o["R"]).max() < 1e-9 and np.abs(g["t"] - o["t"]).max() < 1e-9


def test_gpu_against_real_opencv_golden(ctx):
    probs = [_problem(k) for k in range(N_CASES)]
    res = ctx.pnp_ransac_batch([(p["obj"], p["img"]) for p in probs], probs[0]["intr"])
    same, worst_pose, worst_flags = 0, 0.0, 0
    for k, g in enumerate(res):
        assert g["found"] == 1
        diff = int((g["mask"] != G[f"mask_{k}"]).sum())
        worst_flags = max(worst_flags, diff)
        if diff == 0:
            same += 1
            worst_pose = max(worst_pose, np.abs(g["R"] - G[f"R_{k}"]).max(), np.abs(g["t"] - G[f"t_{k}"]).max())
    assert same >= 116 and worst_flags <= 2 and worst_pose < 1e-6, (same, worst_flags, worst_pose)


def test_batch_equals_single_calls_and_no_model_case(ctx, oracle):
    probs = [synth.make_pnp(4100 + b, 50 + 37 * b, 0.25) for b in range(9)]
    # a frame of pure outliers: no minimal sample reaches 5 inliers with probability ~1 -> found may be 0
    junk = synth.make_pnp(4200, 40, 1.0)
    probs.append(junk)
    batch = ctx.pnp_ransac_batch([(p["obj"], p["img"]) for p in probs], probs[0]["intr"])
    for p, gb in zip(probs, batch):
        g1 = ctx.pnp_ransac(p["obj"], p["img"], p["intr"])
        o = oracle.pnp_ransac(p["obj"], p["img"], p["intr"])
        assert gb["found"] == g1["found"] == o["found"]
        assert np.array_equal(gb["mask"], g1["mask"]) and np.array_equal(gb["mask"], o["mask"])
        assert np.array_equal(gb["R"], g1["R"]) and np.array_equal(gb["t"], g1["t"])
        assert gb["iters"] == o["iters"]


def test_rejects_bad_input(ctx):
    p = synth.make_pnp(4300, 5, 0.0)
    with pytest.raises(U.UrmvoError):
        ctx.pnp_ransac(p["obj"], p["img"], p["intr"])  # fewer than 6 points
    q = synth.make_pnp(4301, 20, 0.0)
    with pytest.raises(U.UrmvoError):
        ctx.pnp_ransac(q["obj"], q["img"], np.array([0.0, 400.0, 320.0, 256.0]))
