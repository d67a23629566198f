"""GPU parity of pose_optimizer / local_bundle_adjuster (FP64 CUDA path through the C ABI) against
the CPU oracle.  Bar (BASELINE.json north_star): final reprojection error within 1e-4 relative;
outlier flags and inlier counts identical."""
import numpy as np
import pytest

from openvslam_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-4  # north-star tolerance on the final reprojection error


def _chi(p, poses, points, xr, mask):
    return synth.reprojection_chi2(p["cam"], poses, points, p["obs_kf"], p["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"], mask)


@pytest.mark.parametrize("model,stereo,n,seed", [("perspective", True, 2000, 1), ("perspective", False, 1000, 2),
                                                 ("equirectangular", False, 4000, 3), ("perspective", True, 37, 4)])
def test_pose_optimizer_matches_oracle(oracle, model, stereo, n, seed):
    from openvslam_b200 import optimize
    p = synth.pose_problem(n, model=model, seed=seed, stereo=stereo)
    xr = p["obs_xr"] if stereo else None
    po = optimize.pose_optimizer()
    ninl, pose, flags, st = po.optimize(optimize.camera(**p["cam"]), not stereo, p["pts_w"], p["obs_xy"], xr, p["inv_sigma_sq"], p["poses"][0])
    on, opose, oflags, ost = oracle.pose_optimize(oracle.camera(**p["cam"]), not stereo, p["pts_w"], p["obs_xy"], xr, p["inv_sigma_sq"], p["poses"][0])
    assert ninl == on and np.array_equal(flags, oflags)
    # Iteration / trial counts are NOT compared: near convergence the LM accept test (rho > 0) is
    # decided by differences at the rounding level of the chi2 sum, whose summation order differs
    # between the CPU loop and the GPU reduction tree.  The damping start values must agree.
    assert st["num_rounds"] == ost["num_rounds"]
    assert np.allclose(st["lambda_init"][:1], ost["lambda_init"][:1], rtol=1e-9)
    assert np.allclose(pose, opose, rtol=0, atol=1e-8)
    pts = p["pts_w"]
    q = dict(p); q["obs_lm"] = np.arange(len(pts), dtype=np.int32)
    c = synth.reprojection_chi2(p["cam"], pose[None], pts, p["obs_kf"], q["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"], ~flags)
    oc = synth.reprojection_chi2(p["cam"], opose[None], pts, p["obs_kf"], q["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"], ~oflags)
    assert abs(c - oc) <= RTOL * oc
    po.close()


def test_pose_optimizer_too_few_observations():
    from openvslam_b200 import optimize
    p = synth.pose_problem(50, seed=9)
    po = optimize.pose_optimizer()
    n, pose, flags, _ = po.optimize(optimize.camera(**p["cam"]), False, p["pts_w"][:4], p["obs_xy"][:4], p["obs_xr"][:4], p["inv_sigma_sq"][:4], p["poses"][0])
    assert n == 0 and np.array_equal(pose, p["poses"][0].reshape(12)) and not flags.any()
    po.close()


@pytest.mark.parametrize("model,stereo,kf,kx,nl,seed", [("equirectangular", False, 8, 3, 800, 1), ("perspective", True, 10, 4, 1500, 2),
                                                       ("perspective", False, 6, 2, 500, 3), ("equirectangular", False, 50, 10, 20000, 4)])
def test_local_ba_matches_oracle(oracle, model, stereo, kf, kx, nl, seed):
    from openvslam_b200 import optimize
    p = synth.ba_problem(kf, kx, nl, model=model, seed=seed, stereo=stereo)
    xr = p["obs_xr"] if stereo else None
    ba = optimize.local_bundle_adjuster()
    poses, points, outl, st = ba.optimize(optimize.camera(**p["cam"]), not stereo, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"],
                                          p["obs_xy"], xr, p["inv_sigma_sq"])
    oposes, opoints, ooutl, ost = oracle.local_ba(oracle.camera(**p["cam"]), not stereo, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"],
                                                  p["obs_xy"], xr, p["inv_sigma_sq"])
    assert st["num_rounds"] == ost["num_rounds"] and len(st["round_iterations"]) == 2
    assert np.allclose(st["lambda_init"][:1], ost["lambda_init"][:1], rtol=1e-9)
    assert (outl != ooutl).mean() < 1e-3   # flags may flip only for an edge sitting on the chi2 bound
    c = _chi(p, poses, points, xr, ~outl)
    oc = _chi(p, oposes, opoints, xr, ~ooutl)
    assert abs(c - oc) <= RTOL * oc, (c, oc)
    assert np.allclose(poses, oposes, rtol=0, atol=1e-5) and np.allclose(points, opoints, rtol=0, atol=1e-4)
    assert np.array_equal(poses[kf:], p["poses"][kf:])  # fixed keyframes untouched
    # size-independent property: the optimum is a fixed point -- a second BA from the result moves nothing much
    c0 = _chi(p, p["poses"], p["points"], xr, None)
    assert c < 0.05 * c0
    ba.close()


def test_local_ba_force_stop_and_validation():
    from openvslam_b200 import optimize, _lib
    p = synth.ba_problem(4, 1, 100, seed=4)
    ba = optimize.local_bundle_adjuster()
    poses, points, outl, st = ba.optimize(optimize.camera(**p["cam"]), True, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"],
                                          p["obs_xy"], None, p["inv_sigma_sq"], force_stop_flag=1)
    assert np.array_equal(poses, p["poses"]) and st["num_iterations"] == 0 and not outl.any()
    perm = np.random.default_rng(0).permutation(len(p["obs_kf"]))
    with pytest.raises(_lib.OvsError):  # observations must be grouped by landmark
        ba.optimize(optimize.camera(**p["cam"]), True, p["poses"], p["fixed"], p["points"], p["obs_kf"][perm], p["obs_lm"][perm],
                    p["obs_xy"][perm], None, p["inv_sigma_sq"][perm])
    ba.close()


def test_local_ba_graph_replay_is_bit_identical():
    """CUDA-graph replay of the LM launch sequences (ovs_optimizer_set_graphs) changes scheduling, not arithmetic;
    also covers re-preparing a different problem on a handle whose graphs were instantiated for another one."""
    from openvslam_b200 import optimize
    ba = optimize.local_bundle_adjuster()
    for kf, kx, nl, seed in ((8, 3, 800, 11), (12, 2, 1500, 12)):
        p = synth.ba_problem(kf, kx, nl, model="equirectangular", seed=seed)
        args = (optimize.camera(**p["cam"]), True, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
        ba.set_graphs(False)
        poses0, points0, outl0, st0 = ba.optimize(*args)
        ba.set_graphs(True)
        poses1, points1, outl1, st1 = ba.optimize(*args)
        assert np.array_equal(poses0, poses1) and np.array_equal(points0, points1) and np.array_equal(outl0, outl1)
        assert st0["num_trials"] == st1["num_trials"] and st0["final_chi2"] == st1["final_chi2"]
    ba.close()


def test_local_ba_largest_supported_system(oracle):
    """110 free keyframes: reduced system n = 660, close to the shared-memory limit of the cluster Cholesky (n <= 684) --
    21 block steps, a narrow last block, the single-buffer back-substitution path."""
    from openvslam_b200 import optimize
    p = synth.ba_problem(110, 4, 2500, model="perspective", seed=31, stereo=False)
    ba = optimize.local_bundle_adjuster()
    args = (p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
    poses, points, outl, st = ba.optimize(optimize.camera(**p["cam"]), True, *args)
    oposes, opoints, ooutl, ost = oracle.local_ba(oracle.camera(**p["cam"]), True, *args)
    assert st["reduced_dim"] == 660 and st["num_rounds"] == ost["num_rounds"]
    assert (outl != ooutl).mean() < 1e-3
    c = _chi(p, poses, points, None, ~outl)
    oc = _chi(p, oposes, opoints, None, ~ooutl)
    assert abs(c - oc) <= RTOL * oc, (c, oc)
    assert np.allclose(poses, oposes, rtol=0, atol=1e-5) and np.allclose(points, opoints, rtol=0, atol=1e-4)
    ba.close()


@pytest.mark.parametrize("kf,nl,seed", [(130, 3000, 33), (200, 2500, 34)])
def test_local_ba_beyond_the_cluster_solver(oracle, kf, nl, seed):
    """More than 114 free keyframes: the reduced system no longer fits the shared-memory panel of the cluster Cholesky and
    the multi-launch path with the panel in global memory takes over (n = 780: narrow last block of 12; n = 1200)."""
    from openvslam_b200 import optimize
    p = synth.ba_problem(kf, 3, nl, model="perspective", seed=seed, stereo=False)
    ba = optimize.local_bundle_adjuster()
    args = (p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
    poses, points, outl, st = ba.optimize(optimize.camera(**p["cam"]), True, *args)
    oposes, opoints, ooutl, ost = oracle.local_ba(oracle.camera(**p["cam"]), True, *args)
    assert st["reduced_dim"] == 6 * kf and st["num_rounds"] == ost["num_rounds"]
    assert (outl != ooutl).mean() < 1e-3
    c = _chi(p, poses, points, None, ~outl)
    oc = _chi(p, oposes, opoints, None, ~ooutl)
    assert abs(c - oc) <= RTOL * oc, (c, oc)
    assert np.allclose(poses, oposes, rtol=0, atol=1e-5) and np.allclose(points, opoints, rtol=0, atol=1e-4)
    ba.close()


def test_local_ba_too_many_keyframes_is_reported():
    """More than 1000 free keyframes: the call fails loudly, there is no fallback."""
    from openvslam_b200 import optimize, _lib
    p = synth.ba_problem(1001, 1, 300, model="perspective", seed=32, stereo=False)
    ba = optimize.local_bundle_adjuster()
    with pytest.raises(_lib.OvsError) as e:
        ba.optimize(optimize.camera(**p["cam"]), True, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
    assert "keyframes" in str(e.value)
    ba.close()


def test_local_ba_speculation_width_is_invisible():
    """1, 2 or 4 Levenberg trials per launch sequence, with or without a static second batch: same accepted states, same counts, same bits."""
    from openvslam_b200 import optimize
    p = synth.ba_problem(10, 3, 1200, model="equirectangular", seed=14)
    args = (optimize.camera(**p["cam"]), True, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
    ba = optimize.local_bundle_adjuster()
    ref = None
    for width, second in ((4, 0), (1, 0), (2, 0), (3, 0), (2, 2), (1, 3), (1, 1)):
        ba.set_speculation(width)
        ba.set_second_batch(second)       # static follow-up batch: skipped on the device when the first one decided
        poses, points, outl, st = ba.optimize(*args)
        cur = (poses, points, outl, st["num_trials"], st["num_iterations"], st["final_chi2"])
        if ref is None:
            ref = cur
        else:
            assert np.array_equal(ref[0], cur[0]) and np.array_equal(ref[1], cur[1]) and np.array_equal(ref[2], cur[2]) and ref[3:] == cur[3:]
    ba.close()


def test_local_ba_host_sync_mode_is_invisible():
    """The Levenberg loop runs on the device; reading its decisions on the host after every batch (to skip launches)
    or never synchronising at all gives the same bits and the same counts."""
    from openvslam_b200 import optimize
    p = synth.ba_problem(9, 2, 900, model="perspective", seed=21, stereo=True)
    args = (optimize.camera(**p["cam"]), False, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], p["obs_xr"], p["inv_sigma_sq"])
    ba = optimize.local_bundle_adjuster()
    res = []
    for mode in (0, 1, -1):
        ba.set_host_sync(mode)
        poses, points, outl, st = ba.optimize(*args)
        res.append((poses, points, outl, st["num_trials"], st["num_iterations"], st["final_chi2"], st["solver_trials"]))
    for r in res[1:]:
        assert np.array_equal(res[0][0], r[0]) and np.array_equal(res[0][1], r[1]) and np.array_equal(res[0][2], r[2]) and res[0][3:] == r[3:]
    assert res[0][3] >= res[0][4] >= 1 and res[0][6] >= res[0][3]   # speculative trials >= the sequential loop's trials
    ba.close()


def test_one_handle_holds_one_problem():
    """the pose optimiser reuses the handle's buffers, so a prepared local BA on the same handle is
    invalidated (run / fetch fail loudly instead of reading overwritten state); preparing again restores it."""
    import ctypes as C
    from openvslam_b200 import optimize, _lib
    p = synth.ba_problem(6, 2, 400, model="equirectangular", seed=22)
    cam = optimize.camera(**p["cam"])
    pba = optimize.prepared_local_ba(cam, True, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
    st0 = pba.run()
    ref = pba.fetch()
    q = synth.pose_problem(300, model="perspective", seed=23, stereo=True)
    po = optimize.pose_optimizer()
    want = po.optimize(optimize.camera(**q["cam"]), False, q["pts_w"], q["obs_xy"], q["obs_xr"], q["inv_sigma_sq"], q["poses"][0])
    po._h, keep = pba._h, po._h                       # same handle for both calls
    got = po.optimize(optimize.camera(**q["cam"]), False, q["pts_w"], q["obs_xy"], q["obs_xr"], q["inv_sigma_sq"], q["poses"][0])
    po._h = keep
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    with pytest.raises(_lib.OvsError):
        pba.run()
    with pytest.raises(_lib.OvsError):
        pba.fetch()
    pba2 = optimize.prepared_local_ba(cam, True, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
    st1 = pba2.run()
    again = pba2.fetch()
    assert st0["num_trials"] == st1["num_trials"] and all(np.array_equal(a, b) for a, b in zip(ref, again))
    po.close(); pba.close(); pba2.close()


def test_local_ba_stop_flag_raised_during_the_run():
    """force_stop_flag raised by another thread while the device-side loop runs: the call returns early with a
    consistent state (fewer iterations than the budget, finite poses); g2o's terminate() semantics."""
    import ctypes as C
    import threading
    import time
    from openvslam_b200 import optimize, _lib
    p = synth.ba_problem(50, 10, 20000, model="equirectangular", seed=4)
    cam = optimize.camera(**p["cam"])
    pba = optimize.prepared_local_ba(cam, True, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
    full = pba.run(5, 400)
    flag = C.c_uint8(0)
    st = optimize.BaStats()

    def raise_flag():
        time.sleep(0.004)
        flag.value = 1
    th = threading.Thread(target=raise_flag)
    th.start()
    _lib.check(_lib.lib().ovs_local_ba_run(pba._h, 5, 400, C.byref(flag), C.byref(st)))
    th.join()
    poses, points, outl = pba.fetch()
    assert np.isfinite(poses).all() and np.isfinite(points).all()
    assert st.num_iterations <= full["num_iterations"]
    pba.close()


def test_local_ba_from_device_resident_graph(oracle):
    """ovs_local_ba_prepare_device / _fetch_device: the graph lives in HBM, the bookkeeping (free-keyframe ids, edge ranges,
    validation) runs on the device; same bits as the host entry, and invalid graphs are still reported."""
    import torch
    from openvslam_b200 import optimize, _lib
    p = synth.ba_problem(12, 3, 1500, model="perspective", seed=41, stereo=True)
    cam = optimize.camera(**p["cam"])
    ba = optimize.local_bundle_adjuster()
    poses_h, points_h, outl_h, st_h = ba.optimize(cam, False, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], p["obs_xr"], p["inv_sigma_sq"])
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(p[k], dt)).to(dev) for k, dt in (("poses", np.float64), ("fixed", np.uint8), ("points", np.float64),
         ("obs_kf", np.int32), ("obs_lm", np.int32), ("obs_xy", np.float32), ("obs_xr", np.float32), ("inv_sigma_sq", np.float32))}
    K, L, M = len(p["poses"]), len(p["points"]), len(p["obs_kf"])
    pba = optimize.prepared_local_ba.from_device(cam, False, K, L, M, t["poses"].data_ptr(), t["fixed"].data_ptr(), t["points"].data_ptr(), t["obs_kf"].data_ptr(),
                                                 t["obs_lm"].data_ptr(), t["obs_xy"].data_ptr(), t["obs_xr"].data_ptr(), t["inv_sigma_sq"].data_ptr())
    st_d = pba.run()
    poses_d, points_d, outl_d = pba.fetch()
    assert np.array_equal(poses_d, poses_h) and np.array_equal(points_d, points_h) and np.array_equal(outl_d, outl_h)
    assert st_d["num_trials"] == st_h["num_trials"] and st_d["final_chi2"] == st_h["final_chi2"]
    o_poses = torch.zeros((K, 12), dtype=torch.float64, device=dev); o_points = torch.zeros((L, 3), dtype=torch.float64, device=dev)
    o_out = torch.zeros(M, dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().ovs_local_ba_fetch_device(pba._h, __import__("ctypes").c_void_p(o_poses.data_ptr()), __import__("ctypes").c_void_p(o_points.data_ptr()),
                                                    __import__("ctypes").c_void_p(o_out.data_ptr())))
    assert np.array_equal(o_poses.cpu().numpy(), poses_h) and np.array_equal(o_points.cpu().numpy(), points_h) and np.array_equal(o_out.cpu().numpy().astype(bool), outl_h)
    # validation on the device: an out-of-range keyframe index and an ungrouped observation list
    bad = t["obs_kf"].clone(); bad[17] = K + 3
    with pytest.raises(_lib.OvsError) as e:
        optimize.prepared_local_ba.from_device(cam, False, K, L, M, t["poses"].data_ptr(), t["fixed"].data_ptr(), t["points"].data_ptr(), bad.data_ptr(),
                                               t["obs_lm"].data_ptr(), t["obs_xy"].data_ptr(), None, t["inv_sigma_sq"].data_ptr(), handle=pba)
    assert "observation 17" in str(e.value)
    perm = torch.from_numpy(np.random.default_rng(0).permutation(M)).to(dev)
    kf_p = t["obs_kf"][perm].contiguous(); lm_p = t["obs_lm"][perm].contiguous()      # kept alive across the call
    with pytest.raises(_lib.OvsError) as e:
        optimize.prepared_local_ba.from_device(cam, False, K, L, M, t["poses"].data_ptr(), t["fixed"].data_ptr(), t["points"].data_ptr(), kf_p.data_ptr(),
                                               lm_p.data_ptr(), t["obs_xy"].data_ptr(), None, t["inv_sigma_sq"].data_ptr(), handle=pba)
    assert "grouped by landmark" in str(e.value)
    ba.close(); pba.close()


@pytest.mark.parametrize("model,stereo,kf,nl,huber,seed", [("perspective", False, 40, 3000, True, 51), ("equirectangular", False, 150, 5000, True, 52),
                                                          ("perspective", True, 260, 6000, False, 53)])
def test_global_ba_matches_oracle(oracle, model, stereo, kf, nl, huber, seed):
    """optimize::global_bundle_adjuster: every keyframe but the origin free, ONE Levenberg round (10 iterations), Huber on every
    edge when requested, no outlier cut.  150 / 260 free keyframes (n = 894 / 1554) take the multi-launch reduced solver."""
    from openvslam_b200 import optimize
    p = synth.ba_problem(kf - 1, 1, nl, model=model, seed=seed, stereo=stereo)
    xr = p["obs_xr"] if stereo else None
    gba = optimize.global_bundle_adjuster(10, huber)
    args = (p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"])
    poses, points, st = gba.optimize(optimize.camera(**p["cam"]), not stereo, *args)
    oposes, opoints, ost = oracle.global_ba(oracle.camera(**p["cam"]), not stereo, *args, num_iter=10, use_huber_kernel=huber)
    assert st["num_rounds"] == 1 == ost["num_rounds"] and st["reduced_dim"] == 6 * (kf - 1)
    assert np.allclose(st["lambda_init"][:1], ost["lambda_init"][:1], rtol=1e-9)
    c = _chi(p, poses, points, xr, None)
    oc = _chi(p, oposes, opoints, xr, None)
    assert abs(c - oc) <= RTOL * oc, (c, oc)
    assert np.allclose(poses, oposes, rtol=0, atol=1e-5) and np.allclose(points, opoints, rtol=0, atol=1e-4)
    assert np.array_equal(poses[kf - 1:], p["poses"][kf - 1:])       # the origin keyframe stays put
    assert c < 0.6 * _chi(p, p["poses"], p["points"], xr, None)      # one robust round on 5 % outliers: the plain chi2 still falls
    gba.close()


def test_local_ba_cluster_width_is_invisible():
    """The reduced-system solver on clusters of 1, 2, 4 or 8 CTAs: same bits (the trailing update is dealt to the warps of the
    cluster by tile index, every element has one writer and a fixed summation order)."""
    from openvslam_b200 import optimize
    p = synth.ba_problem(14, 3, 1800, model="equirectangular", seed=15)
    args = (optimize.camera(**p["cam"]), True, p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], None, p["inv_sigma_sq"])
    ba = optimize.local_bundle_adjuster()
    ref = None
    for width in (8, 1, 2, 4):
        ba.set_cluster_width(width)
        poses, points, outl, st = ba.optimize(*args)
        cur = (poses, points, outl, st["num_trials"], st["final_chi2"])
        if ref is None:
            ref = cur
        else:
            assert np.array_equal(ref[0], cur[0]) and np.array_equal(ref[1], cur[1]) and np.array_equal(ref[2], cur[2]) and ref[3:] == cur[3:]
    ba.close()


@pytest.mark.parametrize("n,end_bit", [(1, 1), (31, 5), (2047, 11), (2048, 11), (2049, 11), (300000, 11), (300000, 7), (1000003, 19), (70000, 23)])
def test_graph_preparation_sort_is_a_stable_sort(n, end_bit):
    """The hand-written radix sort of the BA graph preparation (co-observations by keyframe pair; replaces the one library call
    the path had) against numpy's stable argsort: same keys, same values, equal keys in input order."""
    import ctypes as C
    from openvslam_b200 import _lib
    rng = np.random.default_rng(n + end_bit)
    keys = rng.integers(0, 1 << end_bit, n, dtype=np.uint64).astype(np.uint32)
    if n > 4096:
        keys[: n // 3] = keys[0]                      # a long run of one pair, as the diagonal pairs of a real graph
    vals = (rng.integers(0, 1 << 62, n, dtype=np.uint64) << np.uint64(1)) | np.uint64(1)
    ko = np.zeros(n, np.uint32); vo = np.zeros(n, np.uint64)
    L = _lib.lib()
    _lib.check(L.ovs_debug_sort_pairs(0, keys.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p), n, end_bit,
                                      ko.ctypes.data_as(C.c_void_p), vo.ctypes.data_as(C.c_void_p)))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ko, keys[order]) and np.array_equal(vo, vals[order])
