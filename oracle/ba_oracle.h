/* oracle/ba_oracle.h -- CPU oracle for pose_optimizer / local_bundle_adjuster (test infrastructure only). */
#ifndef BA_ORACLE_H
#define BA_ORACLE_H
#include <stdint.h>

#define OB_CAM_PERSPECTIVE 0
#define OB_CAM_EQUIRECTANGULAR 1
#define OB_MAX_ROUNDS 8

typedef struct {
    int model;
    double fx, fy, cx, cy, focal_x_baseline;
    double cols, rows;
} ob_camera;

typedef struct {
    int num_rounds;                       /* optimizer.optimize() calls */
    int num_iterations;                   /* LM iterations executed in total */
    int num_trials;                       /* linear solves in total */
    int round_iterations[OB_MAX_ROUNDS];
    double lambda_init[OB_MAX_ROUNDS];
    double last_lambda, last_chi2;        /* of the last round */
    double final_chi2;                    /* sum of chi2 over the inlier edges (stored errors) */
} ob_stats;

void ob_se3_exp(const double* u, double* R, double* t);
void ob_pose_oplus(const double* pose, const double* u, double* out);
int ob_edge_eval(const ob_camera* cam, const double* pose, const double* pw, const double* obs, int stereo,
                 double* e, double* Jp, double* Jl, double* pc_out);
int ob_pose_optimize(const ob_camera* cam, int setup_is_mono, int n, const double* pts_w, const float* obs_xy,
                     const float* obs_xr, const float* inv_sigma_sq, double* pose_cw, uint8_t* outlier_flags,
                     int num_trials, int num_each_iter, ob_stats* st);
int ob_local_ba(const ob_camera* cam, int setup_is_mono, int K, double* poses, const uint8_t* fixed, int L, double* points,
                int M, const int* obs_kf, const int* obs_lm, const float* obs_xy, const float* obs_xr,
                const float* inv_sigma_sq, int num_first_iter, int num_second_iter, const volatile int* force_stop,
                uint8_t* outlier_out, ob_stats* st);
int ob_global_ba(const ob_camera* cam, int setup_is_mono, int K, double* poses, const uint8_t* fixed, int L, double* points,
                 int M, const int* obs_kf, const int* obs_lm, const float* obs_xy, const float* obs_xr,
                 const float* inv_sigma_sq, int num_iter, int use_huber_kernel, const volatile int* force_stop, ob_stats* st);
#endif
