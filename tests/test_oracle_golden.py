"""Pins the CPU oracle against the committed cv2 golden vectors (tests/golden/cv2_primitives.npz,
made by tests/golden/make_golden.py with cv2 4.13.0): every OpenCV primitive the reference's ORB
front-end calls must be reproduced bit for bit."""
import numpy as np


def test_resize_linear_bit_exact(oracle, golden):
    img = golden["image"]
    for i in range(3):
        ref = golden["resize_%d" % i]
        out = oracle.resize_linear(img, ref.shape[1], ref.shape[0])
        assert np.array_equal(out, ref)


def test_gaussian_blur_bit_exact(oracle, golden):
    assert np.array_equal(oracle.gaussian7(golden["image"]), golden["blur"])


def test_fast_bit_exact(oracle, golden):
    img = golden["image"]
    for thr in (20, 7):
        ref = golden["fast_%d" % thr]
        out = oracle.fast_detect(img, thr)
        got = np.stack([out["x"], out["y"], out["score"]], 1)
        assert np.array_equal(got, ref)
        roi = np.ascontiguousarray(img[19:19 + 70, 23:23 + 70])
        out = oracle.fast_detect(roi, thr)
        got = np.stack([out["x"], out["y"], out["score"]], 1)
        assert np.array_equal(got, golden["fast_roi_%d" % thr])


def test_fast_score_map_consistent_with_detector(oracle, golden):
    # score map S >= t  <=>  corner at threshold t; detector response == S
    img = golden["image"]
    S = oracle.fast_score_map(img)
    for thr in (20, 7):
        ref = golden["fast_%d" % thr]
        assert np.array_equal(S[ref[:, 1], ref[:, 0]], ref[:, 2].astype(np.uint8))
        assert (ref[:, 2] >= thr).all()


def test_fast_atan2_bit_exact(oracle, golden):
    yx = golden["atan2_in"]
    got = np.array([oracle.fast_atan2(float(y), float(x)) for y, x in yx], np.float32)
    assert np.array_equal(got.view(np.uint32), golden["atan2_out"].view(np.uint32))


def test_orb_descriptor_bit_exact(oracle, golden):
    blur = golden["orb_blur"]
    for (x, y, a), d in zip(golden["orb_kps"], golden["orb_desc"]):
        assert np.array_equal(oracle.orb_descriptor(blur, int(x), int(y), float(a)), d)


def test_umax_table(oracle):
    import ctypes as C
    um = (C.c_int * 16)()
    oracle.lib().oo_umax(um)
    assert list(um) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_color_to_gray_bit_exact(oracle, golden):
    """util::convert_to_grayscale (cv::cvtColor *2GRAY, 15-bit fixed point) -- SURVEY 8f rank 3."""
    col = golden["color_in"]
    assert np.array_equal(oracle.color_to_gray(col[..., :3], False), golden["gray_bgr"])
    assert np.array_equal(oracle.color_to_gray(col[..., :3], True), golden["gray_rgb"])
    assert np.array_equal(oracle.color_to_gray(col, False), golden["gray_bgra"])
    assert np.array_equal(oracle.color_to_gray(col, True), golden["gray_rgba"])


def test_undistort_points_bit_exact(oracle, golden):
    """camera::perspective::undistort_keypoints (cv::undistortPoints, fixed iteration counts) -- SURVEY 8f rank 3."""
    K, d, pts = golden["undist_K"], golden["undist_dist"], golden["undist_in"]
    for iters, key in ((20, "undist_out_20"), (5, "undist_out_5")):
        got = oracle.undistort_points(pts, K[0, 0], K[1, 1], K[0, 2], K[1, 2], d, iters)
        assert np.array_equal(got, golden[key])
