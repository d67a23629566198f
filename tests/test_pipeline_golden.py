"""Committed pipeline vectors (tests/golden/pipeline_golden.npz, made by tests/golden/make_pipeline_golden.py from the CPU oracle on
seeded synthetic inputs -- SURVEY.md 8(c)).  CPU tests: the oracle still reproduces them (a regression pin of the checker
itself).  GPU tests: the CUDA path reproduces them -- bit-exact keypoints / descriptors / match list, pose and BA results within
the north-star tolerances -- without going through the oracle at all."""
import importlib.util
import os

import numpy as np
import pytest

from openvslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_pipeline_golden", os.path.join(ROOT, "tests", "golden", "make_pipeline_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(ROOT, "tests", "golden", "pipeline_golden.npz"))


def _check_extract(vec, name, full, kps, desc):
    assert len(kps) == int(vec[name + "_count"])
    assert np.array_equal(np.bincount(kps["octave"], minlength=8), vec[name + "_per_level"])
    if full:
        assert np.array_equal(G.kp_bits(kps), vec[name + "_kp_bits"])
        assert np.array_equal(desc, vec[name + "_desc"])
    assert str(G.digest(G.kp_bits(kps), desc)) == str(vec[name + "_digest"])


# ------------------------------------------------------------------ CPU: the oracle against its own committed outputs
@pytest.mark.parametrize("case", G.cases(), ids=lambda c: c[0])
def test_oracle_extract_reproduces_golden(oracle, vec, case):
    name, seed, w, h, n, full = case
    kps, desc, _ = oracle.extract(synth.frame(w, h, seed=seed), oracle.params(n))
    _check_extract(vec, name, full, kps, desc)


def test_oracle_match_reproduces_golden(oracle, vec):
    name, seed, w, h, n, _ = G.cases()[0]
    img = synth.frame(w, h, seed=seed)
    _, d1, _ = oracle.extract(img, oracle.params(n))
    k2, d2, _ = oracle.extract(synth.shifted(img, 2, 1), oracle.params(n))
    assert str(G.digest(G.kp_bits(k2), d2)) == str(vec[name + "_shift_digest"])
    assert np.array_equal(oracle.robust_brute_force_match(d1, d2, None, 0.75), vec[name + "_match_pairs"])


def test_oracle_optimisers_reproduce_golden(oracle, vec):
    pc = G.pose_case()
    p = synth.pose_problem(pc["n"], model=pc["model"], seed=pc["seed"], stereo=pc["stereo"])
    ninl, pose, flags, st = oracle.pose_optimize(oracle.camera(**p["cam"]), True, p["pts_w"], p["obs_xy"], None, p["inv_sigma_sq"], p["poses"][0])
    assert ninl == int(vec["pose_num_inliers"]) and np.array_equal(np.asarray(flags, np.uint8), vec["pose_flags"])
    assert np.allclose(pose, vec["pose_pose"], rtol=0, atol=1e-11)
    q = synth.ba_problem(**G.ba_case())
    poses, points, outl, st = oracle.local_ba(oracle.camera(**q["cam"]), True, q["poses"], q["fixed"], q["points"], q["obs_kf"], q["obs_lm"],
                                              q["obs_xy"], None, q["inv_sigma_sq"])
    assert st["num_iterations"] == int(vec["ba_num_iterations"]) and np.array_equal(np.asarray(outl, np.uint8), vec["ba_outliers"])
    assert np.allclose(poses, vec["ba_poses"], rtol=0, atol=1e-10) and np.allclose(points, vec["ba_points"], rtol=0, atol=1e-9)
    assert abs(st["final_chi2"] - float(vec["ba_final_chi2"])) <= 1e-9 * float(vec["ba_final_chi2"])


# ------------------------------------------------------------------ GPU: the CUDA path against the committed vectors (no oracle)
@pytest.mark.gpu
@pytest.mark.parametrize("case", G.cases(), ids=lambda c: c[0])
def test_cuda_extract_reproduces_golden(vec, case):
    from openvslam_b200 import feature
    name, seed, w, h, n, full = case
    ext = feature.orb_extractor(feature.orb_params(max_num_keypts=n))
    kps, desc = ext.extract(synth.frame(w, h, seed=seed))
    _check_extract(vec, name, full, kps, desc)
    ext.close()


@pytest.mark.gpu
def test_cuda_match_reproduces_golden(vec):
    from openvslam_b200 import feature, match
    name, seed, w, h, n, _ = G.cases()[0]
    img = synth.frame(w, h, seed=seed)
    ext = feature.orb_extractor(feature.orb_params(max_num_keypts=n))
    _, d1 = ext.extract(img)
    k2, d2 = ext.extract(synth.shifted(img, 2, 1))
    assert str(G.digest(G.kp_bits(k2), d2)) == str(vec[name + "_shift_digest"])
    mt = match.robust(lowe_ratio=0.75)
    assert np.array_equal(mt.brute_force_match(d1, d2), vec[name + "_match_pairs"])
    ext.close(); mt.close()


@pytest.mark.gpu
def test_cuda_optimisers_reproduce_golden(vec):
    from openvslam_b200 import optimize
    pc = G.pose_case()
    p = synth.pose_problem(pc["n"], model=pc["model"], seed=pc["seed"], stereo=pc["stereo"])
    po = optimize.pose_optimizer()
    ninl, pose, flags, st = po.optimize(optimize.camera(**p["cam"]), True, p["pts_w"], p["obs_xy"], None, p["inv_sigma_sq"], p["poses"][0])
    assert ninl == int(vec["pose_num_inliers"]) and np.array_equal(np.asarray(flags, np.uint8), vec["pose_flags"])
    assert np.allclose(pose, vec["pose_pose"], rtol=0, atol=1e-8)
    assert np.isclose(st["lambda_init"][0], float(vec["pose_lambda_init0"]), rtol=1e-9)
    po.close()
    q = synth.ba_problem(**G.ba_case())
    ba = optimize.local_bundle_adjuster()
    poses, points, outl, st = ba.optimize(optimize.camera(**q["cam"]), True, q["poses"], q["fixed"], q["points"], q["obs_kf"], q["obs_lm"], q["obs_xy"], None,
                                          q["inv_sigma_sq"])
    ba.close()
    gold = float(vec["ba_chi2_of_state"])
    inl = ~np.asarray(outl, bool)
    chi = synth.reprojection_chi2(q["cam"], poses, points, q["obs_kf"], q["obs_lm"], q["obs_xy"], None, q["inv_sigma_sq"], inl)
    differ = np.asarray(outl, np.uint8) != vec["ba_outliers"]
    assert differ.mean() <= 1e-3                                   # only edges sitting on the chi2 bound may flip
    if not differ.any():
        assert abs(chi - gold) <= 1e-4 * gold                      # north-star tolerance on the final reprojection error
    assert np.allclose(poses, vec["ba_poses"], rtol=0, atol=1e-5) and np.allclose(points, vec["ba_points"], rtol=0, atol=1e-4)
    assert st["num_iterations"] == int(vec["ba_num_iterations"])
