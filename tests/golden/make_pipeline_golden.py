#!/usr/bin/env python
"""Generates tests/golden/pipeline_golden.npz: outputs of the CPU oracle (oracle/, the restated reference) for the hot path on
seeded synthetic inputs -- SURVEY.md 8(c) "goldens the repo must create itself": keypoints (every cv::KeyPoint field as bits),
descriptors, brute-force match list, pose-optimiser result and local-BA result, stored with the seeds and the git revision of
the oracle that made them.  Two jobs: (1) a regression pin of the oracle itself (tests/test_pipeline_golden.py, CPU), (2) a
second reference for the CUDA path that does not depend on the oracle being rebuilt on the GPU box (GPU tests of the same file).
The reference itself cannot make these vectors (no source under /root/reference).
Re-run: python tests/golden/make_pipeline_golden.py"""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openvslam_b200 import synth  # noqa: E402
from oracle import oracle as O    # noqa: E402

KP_FIELDS = ("x", "y", "size", "angle", "response")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.array(h.hexdigest())


def kp_bits(kps):
    """(n, 6) uint32: x, y, size, angle, response as float bits + octave"""
    return np.stack([kps[f].view(np.uint32) for f in KP_FIELDS] + [kps["octave"].astype(np.uint32)], 1)


def cases():
    """(name, seed, width, height, max_num_keypts, store_full)"""
    return [("cfg1", 3, 640, 480, 1000, True), ("cfg4", 2880, 1920, 960, 4000, False)]


def pose_case():
    return dict(n=600, model="equirectangular", seed=31, stereo=False)


def ba_case():
    return dict(num_free=8, num_fixed=3, num_landmarks=700, model="equirectangular", seed=32)


def main():
    O.build()
    out = {}
    try:
        rev = subprocess.check_output(["git", "-C", ROOT, "log", "-1", "--format=%H", "--", "oracle"], text=True).strip()
    except Exception:
        rev = "unknown"
    out["oracle_git_revision"] = np.array(rev)
    for name, seed, w, h, n, full in cases():
        img = synth.frame(w, h, seed=seed)
        kps, desc, _ = O.extract(img, O.params(n))
        out[name + "_count"] = np.array(len(kps))
        out[name + "_per_level"] = np.bincount(kps["octave"], minlength=8).astype(np.int32)
        out[name + "_digest"] = digest(kp_bits(kps), desc)
        if full:
            out[name + "_kp_bits"] = kp_bits(kps)
            out[name + "_desc"] = desc
            img2 = synth.shifted(img, 2, 1)
            kps2, desc2, _ = O.extract(img2, O.params(n))
            out[name + "_match_pairs"] = O.robust_brute_force_match(desc, desc2, None, 0.75).astype(np.int32)
            out[name + "_shift_digest"] = digest(kp_bits(kps2), desc2)
    pc = pose_case()
    p = synth.pose_problem(pc["n"], model=pc["model"], seed=pc["seed"], stereo=pc["stereo"])
    ninl, pose, flags, st = O.pose_optimize(O.camera(**p["cam"]), True, p["pts_w"], p["obs_xy"], None, p["inv_sigma_sq"], p["poses"][0])
    out["pose_num_inliers"] = np.array(ninl); out["pose_pose"] = np.asarray(pose, np.float64); out["pose_flags"] = np.asarray(flags, np.uint8)
    out["pose_lambda_init0"] = np.array(st["lambda_init"][0])
    q = synth.ba_problem(**ba_case())
    poses, points, outl, st = O.local_ba(O.camera(**q["cam"]), True, q["poses"], q["fixed"], q["points"], q["obs_kf"], q["obs_lm"], q["obs_xy"], None,
                                         q["inv_sigma_sq"])
    out["ba_poses"] = np.asarray(poses, np.float64); out["ba_points"] = np.asarray(points, np.float64); out["ba_outliers"] = np.asarray(outl, np.uint8)
    out["ba_final_chi2"] = np.array(st["final_chi2"]); out["ba_num_iterations"] = np.array(st["num_iterations"])
    out["ba_chi2_of_state"] = np.array(synth.reprojection_chi2(q["cam"], poses, points, q["obs_kf"], q["obs_lm"], q["obs_xy"], None, q["inv_sigma_sq"],
                                                               ~np.asarray(outl, bool)))
    path = os.path.join(ROOT, "tests", "golden", "pipeline_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; oracle revision", rev)


if __name__ == "__main__":
    main()
