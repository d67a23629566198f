/* camera_oracle.c -- CPU restatement of the camera-model steps either side of the extractor:
 * camera::perspective::undistort_keypoints (cv::undistortPoints with the camera matrix as new projection and a fixed
 * iteration count) and camera::{perspective,equirectangular}::convert_keypoints_to_bearings (camera/perspective.cc,
 * camera/equirectangular.cc, as recalled).  TEST INFRASTRUCTURE ONLY -- the product never links this file.
 * PARITY STATUS: parity unpinned against the real reference (no source; the choice of 20 iterations and of K as the new
 * projection is as recalled).  Pinned: the undistortion is bit-exact (float32 output) against cv2 4.13.0 undistortPoints / undistortPointsIter
 * (tests/test_oracle_cv2.py, tests/golden).  GPU counterpart: k_undistort_bearings (orb_extractor.cu). */
#include <math.h>
#include <stddef.h>
#include "camera_oracle.h"

/* cv::undistortPoints(src, dst, K, dist = {k1, k2, p1, p2, k3}, R = I, P = K, criteria = MAX_ITER iters): the fixed-point
 * iteration of cvUndistortPointsInternal in double precision, result stored as float like the CV_32FC2 destination. */
void oc_undistort_points(const float* xy, int n, double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2,
                         double k3, int iters, float* out_xy) {
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    for (int i = 0; i < n; ++i) {
        double x = ((double)xy[2 * i] - cx) * ifx, y = ((double)xy[2 * i + 1] - cy) * ify;
        const double x0 = x, y0 = y;
        for (int it = 0; it < iters; ++it) {
            const double r2 = x * x + y * y;
            const double icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
            const double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
            const double dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
            x = (x0 - dx) * icdist;
            y = (y0 - dy) * icdist;
        }
        out_xy[2 * i] = (float)(x * fx + cx);
        out_xy[2 * i + 1] = (float)(y * fy + cy);
    }
}

/* camera::perspective::convert_keypoints_to_bearings: normalise ((x - cx) / fx, (y - cy) / fy, 1). */
void oc_bearings_perspective(const float* xy, int n, double fx, double fy, double cx, double cy, double* out3) {
    for (int i = 0; i < n; ++i) {
        const double xn = ((double)xy[2 * i] - cx) / fx, yn = ((double)xy[2 * i + 1] - cy) / fy;
        const double l2 = sqrt(xn * xn + yn * yn + 1.0);
        out3[3 * i] = xn / l2; out3[3 * i + 1] = yn / l2; out3[3 * i + 2] = 1.0 / l2;
    }
}

/* camera::equirectangular::convert_keypoints_to_bearings: longitude / latitude of the pixel -> unit vector. */
void oc_bearings_equirectangular(const float* xy, int n, double cols, double rows, double* out3) {
    for (int i = 0; i < n; ++i) {
        const double lon = ((double)xy[2 * i] / cols - 0.5) * (2.0 * M_PI);
        const double lat = -((double)xy[2 * i + 1] / rows - 0.5) * M_PI;
        out3[3 * i] = cos(lat) * sin(lon);
        out3[3 * i + 1] = -sin(lat);
        out3[3 * i + 2] = cos(lat) * cos(lon);
    }
}

/* camera::equirectangular::reproject_to_image of a bearing (the inverse used by the matchers' callers and by the BA edge):
 * x = cols (0.5 + atan2(bx, bz) / 2 pi), y = rows (0.5 - asin(-by / |b|) / pi). */
void oc_project_equirectangular(const double* b3, int n, double cols, double rows, double* out_xy) {
    for (int i = 0; i < n; ++i) {
        const double bx = b3[3 * i], by = b3[3 * i + 1], bz = b3[3 * i + 2];
        const double l = sqrt(bx * bx + by * by + bz * bz);
        const double lat = -asin(by / l), lon = atan2(bx, bz);
        out_xy[2 * i] = cols * (0.5 + lon / (2.0 * M_PI));
        out_xy[2 * i + 1] = rows * (0.5 - lat / M_PI);
    }
}

/* camera::fisheye::undistort_keypoints = cv::fisheye::undistortPoints(src, dst, K, D = {k1, k2, k3, k4}, R = I, P = K) with
 * its default criteria (COUNT + EPS: at most `max_count` = 10 Newton steps, stop at |theta_fix| < eps = 1e-8), in double
 * precision, float32 output: theta_d = |((x - cx) / fx, (y - cy) / fy)| clipped to pi / 2, Newton on
 * theta (1 + k1 theta^2 + k2 theta^4 + k3 theta^6 + k4 theta^8) = theta_d, scale = tan(theta) / theta_d; a point whose
 * iteration does not converge or flips sign becomes (-1000000, -1000000) (OpenCV >= 4.5).  Pinned bit-for-bit against
 * cv2 4.13.0 (tests/test_oracle_cv2.py). */
void oc_fisheye_undistort_points(const float* xy, int n, double fx, double fy, double cx, double cy, double k1, double k2, double k3, double k4,
                                 int max_count, double eps, float* out_xy) {
    for (int i = 0; i < n; ++i) {
        const double pwx = ((double)xy[2 * i] - cx) / fx, pwy = ((double)xy[2 * i + 1] - cy) / fy;
        double theta_d = sqrt(pwx * pwx + pwy * pwy);
        theta_d = fmin(fmax(-M_PI / 2., theta_d), M_PI / 2.);
        int converged = 0;
        double theta = theta_d, scale = 0.0;
        if (fabs(theta_d) > eps) {
            for (int j = 0; j < max_count; ++j) {
                const double theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta6 * theta2;
                const double k0_theta2 = k1 * theta2, k1_theta4 = k2 * theta4, k2_theta6 = k3 * theta6, k3_theta8 = k4 * theta8;
                const double theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                                         (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
                theta = theta - theta_fix;
                if (fabs(theta_fix) < eps) { converged = 1; break; }
            }
            scale = tan(theta) / theta_d;
        } else {
            converged = 1;
        }
        const int theta_flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
        if (converged && !theta_flipped) {
            const double pux = pwx * scale, puy = pwy * scale;
            /* P = K: pr = K (pu, 1), fi = pr.xy / pr.z with pr.z = 1 */
            const double prx = fx * pux + cx, pry = fy * puy + cy;
            out_xy[2 * i] = (float)(prx / 1.0); out_xy[2 * i + 1] = (float)(pry / 1.0);
        } else {
            out_xy[2 * i] = -1000000.0f; out_xy[2 * i + 1] = -1000000.0f;
        }
    }
}

/* camera::radial_division::undistort_keypoints (camera/radial_division.cc, as recalled): the one-parameter division model,
 * closed form -- normalised distorted point p_d = ((x - cx) / fx, (y - cy) / fy), p_u = p_d / (1 + distortion |p_d|^2),
 * back to pixels with the same intrinsics; double precision, float32 output.  Unpinned (no third-party counterpart). */
void oc_radial_division_undistort_points(const float* xy, int n, double fx, double fy, double cx, double cy, double distortion, float* out_xy) {
    for (int i = 0; i < n; ++i) {
        const double xd = ((double)xy[2 * i] - cx) / fx, yd = ((double)xy[2 * i + 1] - cy) / fy;
        const double r2 = xd * xd + yd * yd;
        const double s = 1.0 / (1.0 + distortion * r2);
        out_xy[2 * i] = (float)(fx * (xd * s) + cx);
        out_xy[2 * i + 1] = (float)(fy * (yd * s) + cy);
    }
}
