/* oracle/match_oracle.h -- CPU oracle for the Hamming matchers (test infrastructure only). */
#ifndef MATCH_ORACLE_H
#define MATCH_ORACLE_H
#include <stdint.h>
#define OM_HAMMING_DIST_THR_LOW 50
#define OM_HAMMING_DIST_THR_HIGH 100
#define OM_MAX_HAMMING_DIST 256
unsigned om_hamming(const uint8_t* a, const uint8_t* b);
void om_bruteforce(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2,
                   int32_t* best_idx, int32_t* best_dist, int32_t* second_dist);
int om_robust_brute_force_match(const uint8_t* desc_frm, int n1, const uint8_t* desc_keyfrm, int n2,
                                const uint8_t* lm_valid_2, float lowe_ratio, int32_t* pairs_out);
#endif

/* ---- windowed (grid) search: data::frame grid + match::projection / match::area / match::stereo ---- */
typedef struct {
    float min_x, min_y;                 /* camera_->img_bounds_.min_x_/min_y_ */
    float inv_cell_width, inv_cell_height;
    int num_grid_cols, num_grid_rows;   /* camera::base: 64 x 48 */
} om_grid;

typedef struct {
    int n;
    const float* x; const float* y;     /* undist_keypts_[i].pt */
    const int* octave;                  /* undist_keypts_[i].octave */
    const float* angle;                 /* undist_keypts_[i].angle */
    const float* x_right;               /* stereo_x_right_ (NULL = all -1) */
    const uint8_t* desc;                /* descriptors_ (n x 32) */
    om_grid grid;
} om_frame;

int om_get_keypoints_in_cell(const om_frame* f, float ref_x, float ref_y, float margin, int min_level, int max_level, int* out);
/* the same, rebuilding the cell lists on every call (slow, obviously right): checker of the cached version */
int om_get_keypoints_in_cell_literal(const om_frame* f, float ref_x, float ref_y, float margin, int min_level, int max_level, int* out);
int om_projection_match_frame_and_landmarks(const om_frame* frm, const float* scale_factors, int nlm, const uint8_t* lm_usable,
                                            const float* reproj_xy, const float* x_right_in_tracking, const int* pred_scale_level,
                                            const uint8_t* lm_desc, const uint8_t* kp_has_observed_lm, float margin, float lowe_ratio,
                                            int* matched_lm_of_kp);
int om_projection_match_current_and_last(const om_frame* curr, const float* scale_factors, int num_scale_levels, int n_last,
                                         const uint8_t* last_usable, const float* reproj_xy, const float* reproj_x_right,
                                         const int* last_scale_level, const float* last_angle, const uint8_t* lm_desc,
                                         const uint8_t* kp_has_observed_lm, float margin, int assume_forward, int assume_backward,
                                         int check_orientation, int* matched_last_of_kp);
int om_projection_match_best(const om_frame* f, int nq, const uint8_t* usable, const float* ref_xy, const float* ref_x_right,
                             const float* margin, const int* min_level, const int* max_level, const float* q_angle, const uint8_t* q_desc,
                             const uint8_t* kp_unavailable, unsigned hamm_dist_thr, int check_orientation, int* matched_query_of_kp);
int om_bow_tree_match_frame_and_keyframe(int n_kf, const uint8_t* desc_kf, const float* angle_kf, const uint8_t* lm_valid_kf, const int* bow_node_kf,
                                         int n_frm, const uint8_t* desc_frm, const float* angle_frm, const int* bow_node_frm,
                                         float lowe_ratio, int check_orientation, int* matched_keyfrm_idx_of_frm);
int om_bow_tree_match_keyframes(int n1, const uint8_t* desc_1, const float* angle_1, const uint8_t* lm_valid_1, const int* bow_node_1,
                                int n2, const uint8_t* desc_2, const float* angle_2, const uint8_t* lm_valid_2, const int* bow_node_2,
                                float lowe_ratio, int check_orientation, int* matched_idx_2_of_1);
int om_fuse_best_keypoints(const om_frame* f, int nq, const uint8_t* usable, const float* reproj_xy, const float* reproj_x_right,
                           const int* pred_level, const uint8_t* lm_desc, const float* scale_factors, const float* inv_level_sigma_sq,
                           float margin, int* best_idx_of_lm);
int om_check_epipolar_constraint(const double* bearing_1, const double* bearing_2, const double* E_12, float bearing_1_scale_factor);
int om_robust_match_for_triangulation(int n1, const uint8_t* desc_1, const double* bearing_1, const int* octave_1, const float* angle_1,
                                      const uint8_t* has_lm_1, const uint8_t* is_stereo_1, const int* bow_node_1,
                                      int n2, const uint8_t* desc_2, const double* bearing_2, const float* angle_2,
                                      const uint8_t* has_lm_2, const uint8_t* is_stereo_2, const int* bow_node_2,
                                      const double* E_12, const double* epipole_in_2, const float* scale_factors_1,
                                      int check_orientation, int* matched_idx_2_of_1);
int om_projection_match_keyframes_mutually(const om_frame* f1, const om_frame* f2, const float* scale_factors, const uint8_t* usable_1,
                                           const float* reproj_1_in_2, const int* pred_level_1_in_2, const uint8_t* lm_desc_1,
                                           const uint8_t* usable_2, const float* reproj_2_in_1, const int* pred_level_2_in_1,
                                           const uint8_t* lm_desc_2, float margin, int* matched_idx_2_of_kp_1);
int om_area_match_in_consistent_area(const om_frame* f1, const om_frame* f2, float* prev_matched_xy, int* matched_idx_2_in_1,
                                     int margin, float lowe_ratio, int check_orientation);
void om_angle_checker_invalid(const float* delta_angles, int n, int histogram_length, int num_bins_to_retain, uint8_t* invalid);
int om_stereo_compute(const uint8_t* const* left_pyr, const uint8_t* const* right_pyr, const int* pyr_w, const int* pyr_h, const int* pyr_stride,
                      int num_levels, const float* scale_factors, const float* inv_scale_factors,
                      int n_left, const float* lx, const float* ly, const int* loct, const uint8_t* ldesc,
                      int n_right, const float* rx, const float* ry, const int* roct, const uint8_t* rdesc,
                      float focal_x_baseline, float true_baseline, float* stereo_x_right, float* depths);
