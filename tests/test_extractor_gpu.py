"""GPU parity of openvslam_b200.feature.orb_extractor (through the C ABI) against the CPU oracle:
bit-exact pyramid, FAST scores, candidate lists, keypoints (every cv::KeyPoint field) and 256-bit
descriptors, at every BASELINE.json configuration size plus edge cases."""
import numpy as np
import pytest

from openvslam_b200 import synth

pytestmark = pytest.mark.gpu

CONFIGS = [  # (width, height, max_num_keypts)  -- BASELINE.json configs[0..4]
    (640, 480, 1000),
    (752, 480, 1000),
    (1241, 376, 2000),
    (1920, 960, 4000),
    (1920, 1080, 2000),
]


def _extractor(n, **kw):
    from openvslam_b200 import feature
    return feature.orb_extractor(feature.orb_params(max_num_keypts=n, **kw))


def _assert_same(oracle, img, ext, n, mask=None, **okw):
    P = oracle.params(n, **okw)
    kps, desc = ext.extract(img, mask)
    okps, odesc, dbg = oracle.extract(img, P, mask=mask)
    assert len(kps) == len(okps), (len(kps), len(okps), dbg)
    for f in ("x", "y", "size", "angle", "response"):
        assert np.array_equal(kps[f].view(np.uint32), okps[f].view(np.uint32)), f
    assert np.array_equal(kps["octave"], okps["octave"])
    assert (kps["class_id"] == -1).all()
    assert np.array_equal(desc, odesc)
    return kps, desc, dbg


@pytest.mark.parametrize("w,h,n", CONFIGS)
def test_extract_bit_exact(oracle, w, h, n):
    img = synth.frame(w, h, seed=w + h)
    ext = _extractor(n)
    kps, desc, dbg = _assert_same(oracle, img, ext, n)
    assert len(kps) >= n * 0.9
    # stage taps: pyramid, score maps and candidate lists are bit-exact too
    P = oracle.params(n)
    levels = oracle.build_pyramid(img, P)
    sf = oracle.scale_factors(1.2, 8)
    for l in range(8):
        assert np.array_equal(ext.image_pyramid(l), levels[l]), l
        ref = oracle.fast_score_map(levels[l]); ref[ref < 7] = 0
        assert np.array_equal(ext.debug_score_map(l), ref), l
        c = oracle.level_candidates(P, levels[l], float(sf[l]))
        got = ext.debug_candidates(l)
        assert np.array_equal(got, np.stack([c["x"], c["y"], c["score"]], 1).reshape(-1, 3)), l
    ext.close()


def test_extract_repeatable_and_reconfigures(oracle):
    ext = _extractor(1000)
    a = synth.frame(752, 480, seed=1)
    b = synth.frame(640, 480, seed=2)
    k1, d1 = ext.extract(a)
    _assert_same(oracle, b, ext, 1000)          # geometry change on the same handle
    k2, d2 = ext.extract(a)
    assert np.array_equal(k1, k2) and np.array_equal(d1, d2)
    ext.close()


def test_extract_masks(oracle):
    img = synth.frame(752, 480, seed=21)
    ext = _extractor(1000)
    mask = np.full(img.shape, 255, np.uint8)
    mask[:, :300] = 0
    mask[200:260, 500:640] = 0
    kps, _, _ = _assert_same(oracle, img, ext, 1000, mask=mask)
    assert (kps["x"] >= 299).all()
    k0, d0 = ext.extract(img, np.zeros_like(img))
    assert len(k0) == 0 and d0.shape == (0, 32)
    ext.close()
    # rectangle masks of orb_params::mask_rects_ ({x_min, x_max, y_min, y_max} in [0,1])
    rects = [[0.0, 0.25, 0.0, 1.0], [0.6, 0.9, 0.5, 1.0]]
    ext = _extractor(1000, mask_rects=rects)
    kps, desc = ext.extract(img)
    okps, odesc, _ = oracle.extract(img, oracle.params(1000), mask=oracle.rect_mask(752, 480, rects))
    assert np.array_equal(kps["x"], okps["x"]) and np.array_equal(kps["y"], okps["y"]) and np.array_equal(desc, odesc)
    ext.close()


def test_extract_threshold_fallback_and_flat_images(oracle):
    # low-contrast image: most cells only yield corners at the min threshold
    base = synth.frame(640, 480, seed=33).astype(np.float32)
    low = np.clip(110 + (base - 110) * 0.12, 0, 255).astype(np.uint8)
    ext = _extractor(1000)
    kps, _, _ = _assert_same(oracle, low, ext, 1000)
    assert len(kps) > 0 and (kps["response"] < 20).any()
    flat = np.full((480, 640), 128, np.uint8)
    k, d = ext.extract(flat)
    assert len(k) == 0
    ext.close()


@pytest.mark.parametrize("w,h", [(97, 83), (130, 200), (333, 131), (1000, 64)])
def test_extract_odd_and_small_sizes(oracle, w, h):
    img = synth.frame(w, h, seed=w * 3 + h)
    for levels in (8, 3, 1):
        sizes = oracle.level_sizes(w, h, 1.2, levels)
        if min(min(s) for s in sizes) < 8:
            continue
        ext = _extractor(300, num_levels=levels)
        _assert_same(oracle, img, ext, 300, num_levels=levels)
        ext.close()


def test_extract_very_wide_image_small_budget(oracle):
    """the first pass of the tree distribution splits every initial node (round(aspect ratio) of them), so a very wide
    image returns far more keypoints than a small max_num_keypts: the handle grows its buffers, the reference's result comes back."""
    img = synth.frame(2000, 60, seed=77)
    ext = _extractor(50, num_levels=1)
    kps, _ = ext.extract(img)
    assert len(kps) > 50 + 67                      # beyond what the handle was sized for at creation
    _assert_same(oracle, img, ext, 50, num_levels=1)
    ext.close()


def test_extract_other_parameters(oracle):
    img = synth.frame(800, 600, seed=44)
    ext = _extractor(1500, scale_factor=1.5, num_levels=5, ini_fast_thr=30, min_fast_thr=10)
    _assert_same(oracle, img, ext, 1500, scale_factor=1.5, num_levels=5, ini_fast_thr=30, min_fast_thr=10)
    ext.close()


def test_extract_strided_input(oracle):
    big = synth.frame(900, 500, seed=55)
    view = big[10:490, 20:772]  # 752 x 480 view with pitch 900
    ext = _extractor(1000)
    kps, desc = ext.extract(view)
    okps, odesc, _ = oracle.extract(np.ascontiguousarray(view), oracle.params(1000))
    assert np.array_equal(kps["x"], okps["x"]) and np.array_equal(desc, odesc)
    ext.close()


def test_invalid_arguments():
    from openvslam_b200 import feature, _lib
    with pytest.raises(_lib.OvsError):
        feature.orb_extractor(feature.orb_params(num_levels=40))
    with pytest.raises(_lib.OvsError):
        feature.orb_extractor(feature.orb_params(ini_fast_thr=5, min_fast_thr=7))
    ext = feature.orb_extractor(feature.orb_params(max_num_keypts=100))
    with pytest.raises(_lib.OvsError):
        ext.extract(np.zeros((30, 30), np.uint8))
    ext.close()


@pytest.mark.parametrize("channels,order", [(3, "BGR"), (3, "RGB"), (4, "BGR"), (4, "RGB")])
def test_extract_color_input(oracle, channels, order):
    """util::convert_to_grayscale + extract in one call: the on-device gray conversion is cv::cvtColor's, bit-exact."""
    from openvslam_b200 import feature, synth
    gray = synth.frame(645, 483, seed=51)      # odd width: the 4-pixel packing has a tail
    rng = np.random.default_rng(51)
    # a colour image whose channels are different perturbations of the same structure
    col = np.stack([np.clip(gray.astype(np.int32) + rng.integers(-40, 41, gray.shape), 0, 255) for _ in range(channels)], -1).astype(np.uint8)
    g = oracle.color_to_gray(col, order == "RGB")
    ext = feature.orb_extractor(feature.orb_params(max_num_keypts=1000))
    k_c, d_c = ext.extract(col, color_order=order)
    assert np.array_equal(ext.image_pyramid(0), g)
    k_g, d_g = ext.extract(g)
    assert len(k_c) > 500 and np.array_equal(k_c, k_g) and np.array_equal(d_c, d_g)
    ext.close()


def test_undistort_keypoints_and_bearings(oracle):
    """camera->undistort_keypoints + convert_keypoints_to_bearings after extract: the undistorted float
    keypoints are bit-exact with the cv2-pinned oracle, the f64 bearings agree to the last bits."""
    from openvslam_b200 import feature, optimize, synth
    img = synth.frame(752, 480, seed=61)
    ext = feature.orb_extractor(feature.orb_params(max_num_keypts=1200))
    kps, _ = ext.extract(img)
    xy = np.stack([kps["x"], kps["y"]], 1)
    fx, fy, cx, cy = 458.654, 457.296, 367.215, 248.375
    dist = (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.011)
    cam = optimize.camera("perspective", fx=fx, fy=fy, cx=cx, cy=cy, cols=752, rows=480)
    for iters in (20, 5):
        und, bear = ext.undistort_keypoints(kps, cam, dist, iters)
        ref = oracle.undistort_points(xy, fx, fy, cx, cy, dist, iters)
        assert np.array_equal(np.stack([und["x"], und["y"]], 1), ref)
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(und[f], kps[f])
        assert np.allclose(bear, oracle.bearings_perspective(ref, fx, fy, cx, cy), rtol=0, atol=4e-16)
    und, bear = ext.undistort_keypoints(kps, cam, None)                      # no distortion: keypoints unchanged
    assert np.array_equal(und, kps)
    ecam = optimize.camera("equirectangular", cols=752, rows=480)
    und, bear = ext.undistort_keypoints(kps, ecam)
    assert np.array_equal(und, kps)
    assert np.allclose(bear, oracle.bearings_equirectangular(xy, 752, 480), rtol=0, atol=1e-15)
    ext.close()


def test_undistort_keypoints_fisheye_and_radial_division(oracle):
    """The other two camera models of SURVEY 8f rank 3: camera::fisheye (cv::fisheye::undistortPoints, oracle pinned bit-for-bit
    against cv2 4.13.0) and camera::radial_division.  The Newton iteration ends in tan(): device and host libm may differ in the
    last bit of the double, so the float32 keypoints are compared to one float ulp (and counted: almost all are identical)."""
    from openvslam_b200 import feature, optimize, synth
    img = synth.frame(1280, 720, seed=62)
    ext = feature.orb_extractor(feature.orb_params(max_num_keypts=2000))
    kps, _ = ext.extract(img)
    xy = np.stack([kps["x"], kps["y"]], 1)
    fx, fy, cx, cy = 400.0, 410.0, 640.3, 359.7
    for dist in ((-0.02, 0.003, -0.001, 0.0002), (0.1, -0.05, 0.02, -0.004), (0.0, 0.0, 0.0, 0.0)):
        cam = optimize.camera("fisheye", fx=fx, fy=fy, cx=cx, cy=cy, cols=1280, rows=720)
        und, bear = ext.undistort_keypoints(kps, cam, dist, 10)
        ref = oracle.fisheye_undistort_points(xy, fx, fy, cx, cy, dist)
        got = np.stack([und["x"], und["y"]], 1)
        assert np.array_equal(got < -9e5, ref < -9e5)                                  # same non-converged points
        ok = ref > -9e5
        assert np.all(np.abs(got[ok] - ref[ok]) <= np.spacing(np.abs(ref[ok]).astype(np.float32)))
        assert (got == ref).mean() > 0.999
        assert np.allclose(bear, oracle.bearings_perspective(got, fx, fy, cx, cy), rtol=0, atol=4e-16)
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(und[f], kps[f])
    cam = optimize.camera("radial_division", fx=fx, fy=fy, cx=cx, cy=cy, cols=1280, rows=720)
    for distortion in (-0.12, 0.05, 0.0):
        und, bear = ext.undistort_keypoints(kps, cam, (distortion,))
        ref = oracle.radial_division_undistort_points(xy, fx, fy, cx, cy, distortion)
        assert np.array_equal(np.stack([und["x"], und["y"]], 1), ref)
        assert np.allclose(bear, oracle.bearings_perspective(ref, fx, fy, cx, cy), rtol=0, atol=4e-16)
    ext.close()
