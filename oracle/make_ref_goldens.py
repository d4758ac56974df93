"""Feeds the committed synthetic problems to oracle/_ref/ref_driver (the UNMODIFIED reference built by
oracle/build_ref.sh) and writes tests/golden/golden_ref.npz — real g2o / Eigen outputs that pin the CPU
oracle (tests/test_golden_oracle.py::test_oracle_against_reference_goldens).  Needs the reference's
toolchain; it has never been run in the image this repo was built in."""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
from urmvo_b200 import synth  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
CAMERA = os.environ.get("URMVO_REF_CAMERA", "/root/reference/configs/camera_settings/aqua.yaml")


def run(mode, payload):
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        open(fi, "wb").write(payload)
        subprocess.check_call([DRIVER, mode, CAMERA, fi, fo])
        return open(fo, "rb").read()


def problems():
    """name -> (kind, problem); small enough for g2o to finish in seconds."""
    out = {}
    for s in (3, 7, 19):
        out[f"ba_small_{s}"] = ("ba", synth.small_ba(seed=s))
    out["ba_cfg1"] = ("ba", synth.cfg1())
    out["ba_noisy"] = ("ba", synth.small_ba(seed=31, rot_sigma_deg=6.0, trans_sigma=0.5, pt_sigma=0.8, outlier_frac=0.1))
    for s in (5, 6):
        out[f"pose_{s}"] = ("pose", synth.make_pose_batch(s, B=1, n_obs=250))
    for s in (1003, 11):
        out[f"tv_{s}"] = ("tv", synth.make_two_view(s, n_keys=400))
    return out


if __name__ == "__main__":
    if not os.path.exists(DRIVER):
        raise SystemExit("oracle/_ref/ref_driver is missing: run oracle/build_ref.sh on a machine with g2o, Eigen3, OpenCV and yaml-cpp")
    G = {}
    for name, (kind, p) in problems().items():
        if kind == "ba":
            Nc, Np, No = p["poses"].shape[0], p["pts"].shape[0], p["uv"].shape[0]
            ids = np.arange(Nc, dtype=np.int32); pids = np.arange(Np, dtype=np.int32)
            buf = struct.pack("3i", Nc, Np, No) + p["intr"].tobytes() + ids.tobytes() + p["poses"].tobytes() + p["fixed"].tobytes() \
                + pids.tobytes() + p["pts"].tobytes() + p["uv"].tobytes() + p["obs_cam"].tobytes() + p["obs_pt"].tobytes()
            out = run("ba", buf)
            G[name + "_poses"] = np.frombuffer(out[:Nc * 56], dtype=np.float64).reshape(Nc, 7)
            G[name + "_pts"] = np.frombuffer(out[Nc * 56:Nc * 56 + Np * 24], dtype=np.float64).reshape(Np, 3)
            G[name + "_inlier"] = np.frombuffer(out[Nc * 56 + Np * 24:], dtype=np.uint8)
        elif kind == "pose":
            n = p["uv"].shape[0]
            out = run("pose", struct.pack("i", n) + p["intr"].tobytes() + p["poses"][0].tobytes() + p["uv"].tobytes() + p["Xw"].tobytes())
            G[name + "_pose"] = np.frombuffer(out[:56], dtype=np.float64)
            G[name + "_inlier"] = np.frombuffer(out[56:56 + n], dtype=np.uint8)
            G[name + "_n"] = np.array(struct.unpack("i", out[56 + n:]))
        else:
            n = p["keys1"].shape[0]
            out = run("tv", struct.pack("3i", n, n, 200) + p["K"].tobytes() + p["keys1"].tobytes() + p["keys2"].tobytes() + p["matches12"].tobytes())
            G[name + "_ok"] = np.array(struct.unpack("i", out[:4]))
            G[name + "_T21"] = np.frombuffer(out[4:68], dtype=np.float32).reshape(4, 4)
            G[name + "_tri"] = np.frombuffer(out[68 + n * 12:], dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_ref.npz"), **G)
    print("wrote tests/golden/golden_ref.npz with", len(G), "arrays")
