"""CPU checks of the BA oracle (it cannot be pinned against g2o, which is absent here): analytic
Jacobians vs finite differences, convergence on ground-truth problems, LM bookkeeping; and of the
product's FP64 device arithmetic (csrc/ba_math.cuh compiled for the host) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from openvslam_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "hostcheck")])
    return C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))


@pytest.mark.parametrize("model,stereo", [("perspective", False), ("perspective", True), ("equirectangular", False)])
def test_jacobians_match_finite_differences(oracle, model, stereo):
    p = synth.ba_problem(3, 1, 30, model=model, seed=1, stereo=stereo)
    cam = oracle.camera(**p["cam"])
    for i in (0, 7, 19):
        k, l = p["obs_kf"][i], p["obs_lm"][i]
        obs = np.array([p["obs_xy"][i][0], p["obs_xy"][i][1], 300.0])
        e, Jp, Jl, _ = oracle.edge_eval(cam, p["poses"][k], p["points"][l], obs, stereo)
        eps = 1e-6
        for a in range(6):
            u = np.zeros(6); u[a] = eps
            e2 = oracle.edge_eval(cam, oracle.pose_oplus(p["poses"][k], u), p["points"][l], obs, stereo)[0]
            e1 = oracle.edge_eval(cam, oracle.pose_oplus(p["poses"][k], -u), p["points"][l], obs, stereo)[0]
            assert np.allclose((e2 - e1) / (2 * eps), Jp[:, a], rtol=1e-5, atol=1e-5 * np.abs(Jp).max())
        for a in range(3):
            d = np.zeros(3); d[a] = eps
            e2 = oracle.edge_eval(cam, p["poses"][k], p["points"][l] + d, obs, stereo)[0]
            e1 = oracle.edge_eval(cam, p["poses"][k], p["points"][l] - d, obs, stereo)[0]
            assert np.allclose((e2 - e1) / (2 * eps), Jl[:, a], rtol=1e-5, atol=1e-5 * np.abs(Jl).max())


@pytest.mark.parametrize("model,stereo", [("equirectangular", False), ("perspective", True), ("perspective", False)])
def test_local_ba_converges_to_ground_truth(oracle, model, stereo):
    p = synth.ba_problem(8, 3, 600, model=model, seed=2, stereo=stereo)
    cam = oracle.camera(**p["cam"])
    xr = p["obs_xr"] if stereo else None
    c0 = synth.reprojection_chi2(p["cam"], p["poses"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"])
    poses, points, outl, st = oracle.local_ba(cam, not stereo, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"],
                                              p["obs_xy"], xr, p["inv_sigma_sq"])
    c1 = synth.reprojection_chi2(p["cam"], poses, points, p["obs_kf"], p["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"], ~outl)
    cgt = synth.reprojection_chi2(p["cam"], p["poses_gt"], p["points_gt"], p["obs_kf"], p["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"], ~p["is_outlier"])
    assert st["num_rounds"] == 2 and st["round_iterations"][0] <= 5 and st["round_iterations"][1] <= 10
    assert c1 < 0.05 * c0 and c1 < 1.2 * cgt
    assert (outl == p["is_outlier"]).mean() > 0.97
    assert np.abs(poses[:8] - p["poses_gt"][:8]).max() < 0.5 * np.abs(p["poses"][:8] - p["poses_gt"][:8]).max()
    assert np.array_equal(poses[8:], p["poses"][8:])  # fixed keyframes untouched


def test_pose_optimizer_recovers_pose_and_flags_outliers(oracle):
    p = synth.pose_problem(800, model="perspective", seed=3, stereo=True)
    cam = oracle.camera(**p["cam"])
    n, pose, flags, st = oracle.pose_optimize(cam, False, p["pts_w"], p["obs_xy"], p["obs_xr"], p["inv_sigma_sq"], p["poses"][0])
    assert st["num_rounds"] == 4 and n == (~flags).sum()
    assert np.abs(pose - p["poses_gt"][0]).max() < 0.02
    assert (flags == p["is_outlier"]).mean() > 0.95
    # fewer than 5 observations: untouched, returns 0
    n, pose2, flags, _ = oracle.pose_optimize(cam, False, p["pts_w"][:4], p["obs_xy"][:4], p["obs_xr"][:4], p["inv_sigma_sq"][:4], p["poses"][0])
    assert n == 0 and np.array_equal(pose2, p["poses"][0].reshape(12)) and not flags.any()


def test_force_stop_flag_returns_immediately(oracle):
    p = synth.ba_problem(4, 1, 100, seed=4)
    cam = oracle.camera(**p["cam"])
    poses, points, outl, st = oracle.local_ba(cam, True, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None,
                                              p["inv_sigma_sq"], force_stop=1)
    assert np.array_equal(poses, p["poses"]) and np.array_equal(points, p["points"]) and st["num_iterations"] == 0


def test_device_ba_arithmetic_equals_oracle(hc, oracle):
    rng = np.random.default_rng(0)
    for model, stereo in (("perspective", False), ("perspective", True), ("equirectangular", False)):
        p = synth.ba_problem(3, 1, 40, model=model, seed=7, stereo=stereo)
        cam = oracle.camera(**p["cam"])
        c = p["cam"]
        camp = np.array([c["fx"], c["fy"], c["cx"], c["cy"], c["focal_x_baseline"], c["cols"], c["rows"]])
        for i in range(0, len(p["obs_kf"]), 5):
            k, l = p["obs_kf"][i], p["obs_lm"][i]
            obs = np.array([p["obs_xy"][i][0], p["obs_xy"][i][1], 310.0])
            e, Jp, Jl, _ = oracle.edge_eval(cam, p["poses"][k], p["points"][l], obs, stereo)
            e2 = np.zeros(3); Jp2 = np.zeros(18); Jl2 = np.zeros(9)
            dim = hc.hc_edge_eval(1 if model == "equirectangular" else 0, camp.ctypes.data_as(C.c_void_p),
                                  np.ascontiguousarray(p["poses"][k]).ctypes.data_as(C.c_void_p),
                                  np.ascontiguousarray(p["points"][l]).ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p), int(stereo),
                                  e2.ctypes.data_as(C.c_void_p), Jp2.ctypes.data_as(C.c_void_p), Jl2.ctypes.data_as(C.c_void_p))
            assert dim == len(e)
            assert np.allclose(e2[:dim], e, rtol=1e-13, atol=1e-12)
            assert np.allclose(Jp2[:6 * dim].reshape(dim, 6), Jp, rtol=1e-12, atol=1e-12)
            assert np.allclose(Jl2[:3 * dim].reshape(dim, 3), Jl, rtol=1e-12, atol=1e-12)
        for _ in range(20):
            u = rng.standard_normal(6) * rng.choice([1e-7, 1e-2, 0.5])
            out = np.zeros(12)
            hc.hc_pose_oplus(np.ascontiguousarray(p["poses"][0]).ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
            assert np.allclose(out, oracle.pose_oplus(p["poses"][0], u), rtol=1e-14, atol=1e-15)


def test_global_ba_oracle_is_one_huber_round(oracle):
    """global_bundle_adjuster (oracle only so far): one Levenberg round over the whole graph with only the
    origin keyframe fixed; with the same iteration count it is exactly the first round of the local BA on the same graph."""
    p = synth.ba_problem(8, 0, 600, model="perspective", seed=21)
    fixed = np.zeros(8, np.uint8); fixed[0] = 1
    cam = oracle.camera(**p["cam"])
    args = (p["poses"], fixed, p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
    gposes, gpoints, gst = oracle.global_ba(cam, True, *args, num_iter=5)
    lposes, lpoints, _, lst = oracle.local_ba(cam, True, *args, num_first_iter=5, num_second_iter=0)
    assert gst["num_rounds"] == 1 and gst["round_iterations"][0] == lst["round_iterations"][0]
    assert np.array_equal(gposes, lposes) and np.array_equal(gpoints, lpoints)
    assert np.array_equal(gposes[0], p["poses"][0])                       # the origin keyframe does not move
    g10 = oracle.global_ba(cam, True, *args, num_iter=10)[2]
    assert g10["final_chi2"] <= gst["final_chi2"] * (1 + 1e-9)
    untouched = oracle.global_ba(cam, True, *args, num_iter=10, force_stop=1)
    assert np.array_equal(untouched[0], p["poses"]) and np.array_equal(untouched[1], p["points"])


# ---------------------------------------------------------------------------------------------------------------------
# Independent pin of the optimisers' FIXED POINT.  g2o is absent, so the Levenberg loop of the oracle
# (damping schedule, Huber weights, Schur complement, update on the manifold) is checked against optimisers that know nothing
# about this code: (a) without a robust kernel the converged cost must equal scipy.optimize.least_squares' (trust-region
# reflective, numerical Jacobian) from the same start; (b) with the Huber kernel the oracle's result must be a stationary
# point of g2o's robust cost -- a quasi-Newton polish (L-BFGS-B, numerical gradient) started there may not find anything better.
def _edge_residuals(cam, poses, points, p, xr):
    """per edge: weighted residual components (3 columns, the third zero for monocular edges)"""
    M = len(p["obs_kf"])
    r = np.zeros((M, 3))
    for k in np.unique(p["obs_kf"]):
        idx = np.flatnonzero(p["obs_kf"] == k)
        uv, _ = synth.project(cam, poses[k], points[p["obs_lm"][idx]])
        sw = np.sqrt(p["inv_sigma_sq"][idx].astype(np.float64))
        r[idx, 0] = sw * (p["obs_xy"][idx, 0] - uv[:, 0]); r[idx, 1] = sw * (p["obs_xy"][idx, 1] - uv[:, 1])
        if xr is not None and uv.shape[1] > 2:
            st = xr[idx] >= 0
            r[idx[st], 2] = sw[st] * (xr[idx][st] - uv[st, 2])
    return r


def _robust_cost(cam, poses, points, p, xr, delta):
    chi = (_edge_residuals(cam, poses, points, p, xr) ** 2).sum(1)
    if delta is None:
        return chi.sum()
    d2 = delta * delta
    return np.where(chi <= d2, chi, 2 * np.sqrt(chi) * delta - d2).sum()


def _unpacker(oracle, p):
    L = len(p["points"])
    free = np.flatnonzero(p["fixed"] == 0)

    def unpack(x, base_poses, base_points):
        ps = base_poses.copy()
        for j, k in enumerate(free):
            ps[k] = oracle.pose_oplus(base_poses[k], x[6 * j:6 * j + 6])
        return ps, base_points + x[6 * len(free):].reshape(L, 3)
    return unpack, 6 * len(free) + 3 * L


@pytest.mark.parametrize("model,stereo,seed", [("perspective", True, 71), ("perspective", False, 72)])
def test_plain_ba_converges_to_the_scipy_least_squares_optimum(oracle, model, stereo, seed):
    scipy_opt = pytest.importorskip("scipy.optimize")
    p = synth.ba_problem(3, 2, 40, model=model, seed=seed, stereo=stereo, outlier_frac=0.0)
    cam = p["cam"]; xr = p["obs_xr"] if stereo else None
    args = (p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"])
    poses, points, st = oracle.global_ba(oracle.camera(**cam), not stereo, *args, num_iter=60, use_huber_kernel=False)
    c_oracle = _robust_cost(cam, poses, points, p, xr, None)
    unpack, nx = _unpacker(oracle, p)
    sol = scipy_opt.least_squares(lambda x: _edge_residuals(cam, *unpack(x, p["poses"], p["points"]), p, xr).ravel(), np.zeros(nx), method="trf",
                                  xtol=1e-13, ftol=1e-13, gtol=1e-13, max_nfev=200)
    c_scipy = 2 * sol.cost
    assert abs(c_oracle - c_scipy) <= 1e-6 * c_scipy, (c_oracle, c_scipy, st)
    assert c_oracle < 0.1 * _robust_cost(cam, p["poses"], p["points"], p, xr, None)


# (the equirectangular edge is covered by the stationarity test only: its longitude wraps at the seam, where two optimisers started
#  from the same noisy estimate may settle in different basins)
@pytest.mark.parametrize("model,stereo,seed", [("perspective", True, 74), ("equirectangular", False, 75), ("perspective", False, 76)])
def test_huber_ba_result_is_a_stationary_point_of_the_robust_cost(oracle, model, stereo, seed):
    scipy_opt = pytest.importorskip("scipy.optimize")
    p = synth.ba_problem(3, 2, 40, model=model, seed=seed, stereo=stereo, outlier_frac=0.1)
    cam = p["cam"]; xr = p["obs_xr"] if stereo else None
    delta = float(np.sqrt(np.float32(7.81473))) if stereo else float(np.sqrt(np.float32(5.99146)))
    args = (p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"])
    poses, points, st = oracle.global_ba(oracle.camera(**cam), not stereo, *args, num_iter=100, use_huber_kernel=True)
    c_oracle = _robust_cost(cam, poses, points, p, xr, delta)
    c_start = _robust_cost(cam, p["poses"], p["points"], p, xr, delta)
    assert c_oracle < 0.6 * c_start
    unpack, nx = _unpacker(oracle, p)
    f = lambda x: _robust_cost(cam, *unpack(x, poses, points), p, xr, delta)          # noqa: E731  (parametrised around the oracle's result)
    g = scipy_opt.approx_fprime(np.zeros(nx), f, 1e-7)
    g0 = scipy_opt.approx_fprime(np.zeros(nx), lambda x: _robust_cost(cam, *unpack(x, p["poses"], p["points"]), p, xr, delta), 1e-7)
    assert np.abs(g).max() <= 1e-3 * np.abs(g0).max(), (np.abs(g).max(), np.abs(g0).max())
    pol = scipy_opt.minimize(f, np.zeros(nx), method="L-BFGS-B", options=dict(maxiter=200, ftol=1e-15, gtol=1e-10))
    assert c_oracle - pol.fun <= 1e-4 * c_oracle, (c_oracle, pol.fun)


def test_pose_optimizer_round_is_the_huber_optimum(oracle):
    """One round (num_trials = 1) of pose_optimizer = the Huber optimum over all edges, 6 parameters: a generic quasi-Newton
    minimiser from the same start lands on the same cost."""
    scipy_opt = pytest.importorskip("scipy.optimize")
    p = synth.pose_problem(150, model="perspective", seed=81, stereo=True, outlier_frac=0.15)
    cam = p["cam"]; xr = p["obs_xr"]
    delta = float(np.sqrt(np.float32(7.81473)))
    n, pose, flags, st = oracle.pose_optimize(oracle.camera(**cam), False, p["pts_w"], p["obs_xy"], xr, p["inv_sigma_sq"], p["poses"][0], num_trials=1,
                                             num_each_iter=100)
    q = dict(p); q["obs_lm"] = np.arange(len(p["pts_w"]), dtype=np.int32)
    f = lambda x: _robust_cost(cam, oracle.pose_oplus(p["poses"][0], x)[None], p["pts_w"], q, xr, delta)     # noqa: E731
    sol = scipy_opt.minimize(f, np.zeros(6), method="BFGS", options=dict(gtol=1e-9))
    c_oracle = _robust_cost(cam, np.asarray(pose)[None], p["pts_w"], q, xr, delta)
    assert abs(c_oracle - sol.fun) <= 1e-4 * sol.fun, (c_oracle, sol.fun)
