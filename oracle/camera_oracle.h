/* camera_oracle.h -- see camera_oracle.c.  TEST INFRASTRUCTURE ONLY. */
#ifndef CAMERA_ORACLE_H
#define CAMERA_ORACLE_H
void oc_undistort_points(const float* xy, int n, double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2,
                         double k3, int iters, float* out_xy);
void oc_bearings_perspective(const float* xy, int n, double fx, double fy, double cx, double cy, double* out3);
void oc_bearings_equirectangular(const float* xy, int n, double cols, double rows, double* out3);
void oc_project_equirectangular(const double* b3, int n, double cols, double rows, double* out_xy);
void oc_fisheye_undistort_points(const float* xy, int n, double fx, double fy, double cx, double cy, double k1, double k2, double k3, double k4,
                                 int max_count, double eps, float* out_xy);
void oc_radial_division_undistort_points(const float* xy, int n, double fx, double fy, double cx, double cy, double distortion, float* out_xy);
#endif
