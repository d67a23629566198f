#!/usr/bin/env python
"""Generates tests/golden/cv2_primitives.npz: outputs of the OpenCV primitives the reference's
ORB front-end calls (cv::resize INTER_LINEAR, cv::GaussianBlur 7x7 s=2, cv::FAST 9_16 + NMS,
cv::fastAtan2, cv::ORB descriptors for given keypoints) on a small seeded image, computed with
the cv2 wheel in this image (cv2 %s).  The oracle is pinned against these vectors by
tests/test_oracle_golden.py.  Re-run: python tests/golden/make_golden.py
"""
import os, sys
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openvslam_b200 import synth  # noqa: E402

def main():
    img = synth.frame(200, 150, seed=1234)
    out = {"image": img, "cv2_version": np.array(cv2.__version__)}
    for i, (dw, dh) in enumerate([(167, 125), (139, 104), (73, 41)]):
        out["resize_%d" % i] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out["blur"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    for thr in (20, 7):
        f = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        k = f.detect(img)
        out["fast_%d" % thr] = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)
        roi = np.ascontiguousarray(img[19:19 + 70, 23:23 + 70])
        k = f.detect(roi)
        out["fast_roi_%d" % thr] = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)
    rng = np.random.default_rng(7)
    yx = rng.integers(-300000, 300000, (4000, 2)).astype(np.float32)
    yx[:50] = rng.integers(-3, 4, (50, 2))
    out["atan2_in"] = yx
    out["atan2_out"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    # ORB descriptors of given keypoints; cv2.ORB blurs a sub-matrix of its pyramid, which takes
    # OpenCV's non-bit-exact separable filter path, reproduced by sepFilter2D with the float kernel.
    k32 = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    out["orb_blur"] = cv2.sepFilter2D(img, -1, k32, k32, borderType=cv2.BORDER_REFLECT_101)
    orb = cv2.ORB_create(nfeatures=500, scaleFactor=1.2, nlevels=8, edgeThreshold=19, patchSize=31)
    kps = []
    for _ in range(400):
        x = int(rng.integers(25, 175)); y = int(rng.integers(25, 125))
        kps.append(cv2.KeyPoint(float(x), float(y), 31.0, float(rng.uniform(0, 360)), 1.0, 0))
    k2, d = orb.compute(img, kps)
    out["orb_kps"] = np.array([(p.pt[0], p.pt[1], p.angle) for p in k2], np.float32)
    out["orb_desc"] = d
    # util::convert_to_grayscale = cv::cvtColor(..., *2GRAY)
    col = rng.integers(0, 256, (61, 83, 4), dtype=np.uint8)
    out["color_in"] = col
    out["gray_bgr"] = cv2.cvtColor(np.ascontiguousarray(col[..., :3]), cv2.COLOR_BGR2GRAY)
    out["gray_rgb"] = cv2.cvtColor(np.ascontiguousarray(col[..., :3]), cv2.COLOR_RGB2GRAY)
    out["gray_bgra"] = cv2.cvtColor(col, cv2.COLOR_BGRA2GRAY)
    out["gray_rgba"] = cv2.cvtColor(col, cv2.COLOR_RGBA2GRAY)
    # camera::perspective::undistort_keypoints = cv::undistortPoints(pts, K, dist, R=I, P=K, MAX_ITER 20) (EuRoC-like intrinsics)
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]])
    dist = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.011])
    pts = np.stack([rng.uniform(0, 752, 500), rng.uniform(0, 480, 500)], 1).astype(np.float32)
    out["undist_K"] = K; out["undist_dist"] = dist; out["undist_in"] = pts
    out["undist_out_20"] = cv2.undistortPointsIter(pts.reshape(-1, 1, 2), K, dist, None, K, (cv2.TERM_CRITERIA_MAX_ITER, 20, 1e-6)).reshape(-1, 2)
    out["undist_out_5"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cv2_primitives.npz"), **out)
    print("written", {k: getattr(v, "shape", None) for k, v in out.items()})

if __name__ == "__main__":
    main()
