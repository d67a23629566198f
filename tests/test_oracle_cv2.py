"""Cross-checks the oracle against the live cv2 module on fresh seeded inputs (beyond the
committed goldens), and documents the sin/cos convention."""
import ctypes
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from openvslam_b200 import synth  # noqa: E402


def test_pyramid_chain_matches_cv2(oracle):
    img = synth.frame(752, 480, seed=11)
    P = oracle.params(1000)
    levels = oracle.build_pyramid(img, P)
    prev = img
    for l in range(1, 8):
        h, w = levels[l].shape
        ref = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(levels[l], ref), l
        prev = ref
    assert [lv.shape for lv in levels][:3] == [(480, 752), (400, 627), (333, 522)]


def test_cellwise_fast_matches_cv2(oracle):
    img = synth.frame(320, 240, seed=12)
    for thr in (20, 7):
        f = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        for (x0, y0, w, h) in [(19, 19, 70, 70), (83, 147, 70, 74), (250, 19, 51, 70)]:
            roi = np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w])
            ref = [(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in f.detect(roi)]
            out = oracle.fast_detect(roi, thr)
            assert [(int(p["x"]), int(p["y"]), int(p["score"])) for p in out] == ref


def test_ic_angle_uses_fastatan2(oracle):
    img = synth.frame(200, 150, seed=13)
    for (x, y) in [(30, 40), (100, 75), (170, 120)]:
        a, m01, m10 = oracle.ic_angle(img, x, y)
        assert np.float32(a) == np.float32(cv2.fastAtan2(float(m01), float(m10)))
        # moments against a direct evaluation over the circular patch
        um = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
        M01 = M10 = 0
        for v in range(-15, 16):
            for u in range(-um[abs(v)], um[abs(v)] + 1):
                M01 += v * int(img[y + v, x + u]); M10 += u * int(img[y + v, x + u])
        assert (M01, M10) == (m01, m10)


def test_sincos_convention_vs_libm(oracle):
    """The oracle defines sin/cos of the keypoint angle as the correctly rounded float of the
    double result.  The reference calls std::cos(float) (glibc cosf, < 1 ULP, not always
    correctly rounded).  Measure the disagreement; it must be rare and never exceed 1 ULP."""
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = ctypes.c_float; libm.cosf.argtypes = [ctypes.c_float]
    libm.sinf.restype = ctypes.c_float; libm.sinf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(0)
    angles = rng.uniform(0, 2 * np.pi, 20000).astype(np.float32)
    bad = 0
    for a in angles:
        s, c = oracle.sincosf(float(a))
        ls, lc = libm.sinf(float(a)), libm.cosf(float(a))
        if s != ls or c != lc:
            bad += 1
            assert abs(np.float32(s).view(np.int32) - np.float32(ls).view(np.int32)) <= 1
            assert abs(np.float32(c).view(np.int32) - np.float32(lc).view(np.int32)) <= 1
    assert bad / len(angles) < 0.05


def test_extract_invariants(oracle):
    img = synth.frame(640, 480, seed=3)
    P = oracle.params(1000)
    kps, desc, dbg = oracle.extract(img, P)
    per_level = oracle.keypts_per_level(1000, 1.2, 8)
    assert dbg["level_w"][:3] == [640, 533, 444]
    for l in range(8):
        n = int((kps["octave"] == l).sum())
        assert n == dbg["num_selected"][l]
        assert per_level[l] <= n <= per_level[l] + 3 or dbg["num_candidates"][l] < per_level[l]
    assert len(kps) == len(desc) and len(kps) >= 1000
    # keypoints stay 19 px inside their level, responses are FAST scores >= min threshold
    sf = oracle.scale_factors(1.2, 8)
    for k in kps[::37]:
        l = k["octave"]
        assert 19 <= k["lx"] < dbg["level_w"][l] - 19 and 19 <= k["ly"] < dbg["level_h"][l] - 19
        assert k["response"] >= 7 and 0 <= k["angle"] <= 360
        assert k["size"] == np.float32(int(31 * sf[l]))


def test_extract_mask_and_rects(oracle):
    img = synth.frame(400, 300, seed=5)
    P = oracle.params(500)
    mask = oracle.rect_mask(400, 300, [[0.0, 0.5, 0.0, 1.0]])  # left half masked out
    kps, _, _ = oracle.extract(img, P, mask=mask)
    assert len(kps) > 0 and (kps["x"] >= 200 - 1e-3).all()
    empty, _, _ = oracle.extract(img, P, mask=np.zeros_like(img))
    assert len(empty) == 0


@pytest.mark.parametrize("w,h", [(640, 480), (37, 53), (1, 1), (129, 7)])
def test_color_to_gray_live(oracle, w, h):
    """util::convert_to_grayscale against the cv2 in this image, all four channel orders."""
    rng = np.random.default_rng(w * 1000 + h)
    col = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    bgr = np.ascontiguousarray(col[..., :3])
    assert np.array_equal(oracle.color_to_gray(bgr, False), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    assert np.array_equal(oracle.color_to_gray(bgr, True), cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY))
    assert np.array_equal(oracle.color_to_gray(col, False), cv2.cvtColor(col, cv2.COLOR_BGRA2GRAY))
    assert np.array_equal(oracle.color_to_gray(col, True), cv2.cvtColor(col, cv2.COLOR_RGBA2GRAY))


def test_undistort_points_live(oracle):
    """camera::perspective::undistort_keypoints against the cv2 in this image: several lens models, 20 iterations
    (OpenVSLAM's criteria) and OpenCV's default 5."""
    rng = np.random.default_rng(9)
    for fx, fy, cx, cy, dist in ((458.654, 457.296, 367.215, 248.375, (-0.2834, 0.0740, 1.9e-4, 1.8e-5, 0.0)),
                                 (718.9, 718.9, 607.2, 185.2, (0.0, 0.0, 0.0, 0.0, 0.0)),
                                 (520.9, 521.0, 325.1, 249.7, (0.2312, -0.7849, -0.0033, -0.0001, 0.9172))):
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
        pts = np.stack([rng.uniform(0, 2 * cx, 800), rng.uniform(0, 2 * cy, 800)], 1).astype(np.float32)
        ref20 = cv2.undistortPointsIter(pts.reshape(-1, 1, 2), K, np.array(dist), None, K, (cv2.TERM_CRITERIA_MAX_ITER, 20, 1e-6)).reshape(-1, 2)
        ref5 = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, np.array(dist), None, K).reshape(-1, 2)
        assert np.array_equal(oracle.undistort_points(pts, fx, fy, cx, cy, dist, 20), ref20)
        assert np.array_equal(oracle.undistort_points(pts, fx, fy, cx, cy, dist, 5), ref5)


def test_bearings_and_equirectangular_round_trip(oracle):
    """convert_keypoints_to_bearings: unit vectors; perspective against numpy; equirectangular against its own projection
    and against the BA oracle's equirectangular edge (residual 0 at the pixel the bearing came from)."""
    rng = np.random.default_rng(10)
    xy = np.stack([rng.uniform(0, 1920, 500), rng.uniform(1, 959, 500)], 1).astype(np.float32)
    b = oracle.bearings_equirectangular(xy, 1920, 960)
    assert np.allclose(np.linalg.norm(b, axis=1), 1.0, atol=1e-15)
    back = oracle.project_equirectangular(b, 1920, 960)
    assert np.allclose(back, xy.astype(np.float64), atol=1e-9)
    cam = oracle.camera("equirectangular", cols=1920, rows=960)
    pose = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.float64)
    for i in range(0, 500, 50):
        e = oracle.edge_eval(cam, pose, 7.5 * b[i], np.array([xy[i, 0], xy[i, 1], -1.0]), False)[0]
        assert abs(e[0]) < 1e-9 and abs(e[1]) < 1e-9
    bp = oracle.bearings_perspective(xy[:100], 500.0, 510.0, 320.0, 240.0)
    ref = np.stack([(xy[:100, 0].astype(np.float64) - 320.0) / 500.0, (xy[:100, 1].astype(np.float64) - 240.0) / 510.0, np.ones(100)], 1)
    ref /= np.linalg.norm(ref, axis=1, keepdims=True)
    assert np.allclose(bp, ref, rtol=0, atol=1e-15)


def test_fisheye_undistort_points_vs_cv2(oracle):
    """camera::fisheye::undistort_keypoints = cv::fisheye::undistortPoints(pts, K, D, R = I, P = K): bit-for-bit, including the
    (-1e6, -1e6) marker of points whose Newton iteration does not converge."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    fx, fy, cx, cy = 400.0, 410.0, 640.3, 359.7
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    for D in ([-0.02, 0.003, -0.001, 0.0002], [0.1, -0.05, 0.02, -0.004], [0, 0, 0, 0], [0.5, 0.3, 0.1, 0.05]):
        pts = np.stack([rng.uniform(-200, 1500, 4000), rng.uniform(-200, 920, 4000)], 1).astype(np.float32)
        pts[0] = [cx, cy]
        ref = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2), K, np.array(D, np.float64), None, K).reshape(-1, 2)
        got = oracle.fisheye_undistort_points(pts, fx, fy, cx, cy, D)
        assert np.array_equal(got, ref)


def test_hamming_nearest_neighbours_match_cv2_bfmatcher(oracle):
    """The oracle's 256-bit Hamming distance and its nearest / second-nearest scan (match_oracle.c, the arithmetic under every
    matcher) against cv2.BFMatcher(NORM_HAMMING).knnMatch: same best and second-best distances for every query; the same best
    index wherever the best distance is unique (OpenCV's tie order is its own)."""
    import cv2
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (900, 32), dtype=np.uint8)
    t[:200] = q[:200] ^ (rng.integers(0, 256, (200, 32), dtype=np.uint8) & rng.integers(0, 256, (200, 32), dtype=np.uint8)
                        & rng.integers(0, 256, (200, 32), dtype=np.uint8))          # near duplicates: small distances
    bi, bd, sd = oracle.bruteforce(q, t)
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    cb = np.array([m[0].distance for m in knn], np.int32); cs = np.array([m[1].distance for m in knn], np.int32)
    ci = np.array([m[0].trainIdx for m in knn], np.int32)
    assert np.array_equal(bd, cb) and np.array_equal(sd, cs)
    unique = cb < cs
    assert np.array_equal(bi[unique], ci[unique])
    # single pairs through the distance function itself
    for a, b in ((0, 0), (3, 7), (699, 899)):
        assert oracle.hamming(q[a], t[b]) == int(cv2.norm(q[a], t[b], cv2.NORM_HAMMING))


def test_pose_optimizer_fixed_point_matches_cv2_solvepnp(oracle):
    """pose_optimizer's oracle (ba_oracle.c) against OpenCV's own pose refinement, which shares no code with it: on inlier-only
    data with unit information the Huber kernel is inactive and no edge is cut, so the optimiser's result must be the plain
    least-squares reprojection optimum -- cv2.solvePnP (iterative) + cv2.solvePnPRefineLM from the same start."""
    import cv2
    from openvslam_b200 import synth
    p = synth.pose_problem(400, model="perspective", seed=77, stereo=False, pixel_sigma=0.5, outlier_frac=0.0)
    w = np.ones_like(p["inv_sigma_sq"])
    cam = p["cam"]
    ninl, pose, flags, st = oracle.pose_optimize(oracle.camera(**cam), True, p["pts_w"], p["obs_xy"], None, w, p["poses"][0])
    assert ninl == 400 and not np.asarray(flags).any()
    K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1.0]])
    rvec0, _ = cv2.Rodrigues(p["poses"][0][:9].reshape(3, 3))
    pts, obs = p["pts_w"].astype(np.float64), p["obs_xy"].astype(np.float64)
    ok, rvec, tvec = cv2.solvePnP(pts, obs, K, None, rvec0.copy(), p["poses"][0][9:].reshape(3, 1).copy(), True, cv2.SOLVEPNP_ITERATIVE)
    assert ok
    rvec, tvec = cv2.solvePnPRefineLM(pts, obs, K, None, rvec, tvec, criteria=(cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 200, 1e-14))
    R, _ = cv2.Rodrigues(rvec)
    ref = np.concatenate([R.reshape(-1), tvec.reshape(-1)])

    def cost(ps):
        return synth.reprojection_chi2(cam, ps[None], p["pts_w"], p["obs_kf"], np.arange(len(pts), dtype=np.int32), p["obs_xy"], None, w, None)
    assert np.allclose(np.asarray(pose), ref, rtol=0, atol=1e-9)
    assert abs(cost(np.asarray(pose)) - cost(ref)) <= 1e-10 * cost(ref)
    assert abs(st["final_chi2"] - cost(ref)) <= 1e-9 * cost(ref)
