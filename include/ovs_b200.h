/*
 * ovs_b200.h -- C ABI of the B200-native OpenVSLAM hot path (libovs_b200.so).
 *
 * This is the drop-in boundary of SURVEY.md section 8(b): plain pointers and sizes, int
 * return codes, no exceptions, no torch / OpenCV / Eigen types.  The reference has no
 * FFI for this path -- its boundary is the C++ class surface
 *   openvslam::feature::orb_extractor                       (src/openvslam/feature/orb_extractor.h)
 *   openvslam::match::{area,projection,robust,stereo}       (src/openvslam/match/*.h)
 *   openvslam::optimize::{pose_optimizer,local_bundle_adjuster} (src/openvslam/optimize/*.h)
 * [file names as recalled in SURVEY.md 8(a); /root/reference holds no source, so no line
 * numbers can be cited].  include/openvslam_b200/*.h re-declares those classes on top of the
 * entry points below; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *  - every function returns OVS_OK (0) or a negative OVS_ERR_* code; ovs_last_error()
 *    returns a thread-local message for the last failure.
 *  - handles own their CUDA stream, device buffers and pinned staging; a handle must not be
 *    used from two threads at once (the reference uses one extractor per camera, and the
 *    stereo frame constructor runs two extractor instances on two threads).
 *  - *_host entry points take HOST buffers (copies are inside the call); *_device entry points
 *    take DEVICE buffers and leave results on the device.
 *  - there is no CPU fallback: if no sm_100 device is present, *_create fails.
 */
#ifndef OVS_B200_H
#define OVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVS_OK 0
#define OVS_ERR_INVALID_ARG (-1)
#define OVS_ERR_CUDA (-2)
#define OVS_ERR_NO_DEVICE (-3)
#define OVS_ERR_CAPACITY (-4)      /* caller-provided output capacity too small */
#define OVS_ERR_OVERFLOW (-5)      /* internal candidate buffer overflow (pathological image) */
#define OVS_ERR_UNSUPPORTED (-6)
#define OVS_ERR_NUMERIC (-7)       /* linear solve failed (not positive definite) */

const char* ovs_last_error(void);
/* Library / build identification: "ovs_b200 <version> sm_100a". */
const char* ovs_version(void);
/* Number of CUDA kernels launched by this library in the calling process so far. */
uint64_t ovs_kernel_launch_count(void);
/* How host threads wait for the device inside the calls below.  0 (default): spin -- lowest latency, right for one
 * camera stream per GPU.  1: sleep on a blocking CUDA event (frees the core; ~100 us per wake-up).  2: poll with
 * sched_yield() -- near-spin latency while cores are free, fair sharing once host threads outnumber cores.
 * Process-wide; call it before creating the handles it should apply to. */
int ovs_set_wait_mode(int mode);

/* ------------------------------------------------------------------ feature::orb_extractor */

/* openvslam::feature::orb_params (feature/orb_params.h): same field names. */
typedef struct {
    uint32_t max_num_keypts;   /* Feature.max_num_keypoints */
    float scale_factor;        /* Feature.scale_factor */
    uint32_t num_levels;       /* Feature.num_levels (1..16) */
    uint32_t ini_fast_thr;     /* Feature.ini_fast_threshold */
    uint32_t min_fast_thr;     /* Feature.min_fast_threshold */
} ovs_orb_params;

/* Binary-compatible with cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave,
 * class_id -- so a std::vector<cv::KeyPoint>::data() can be passed directly. */
typedef struct {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} ovs_keypoint;

typedef struct ovs_extractor ovs_extractor;

/* orb_extractor::orb_extractor(const orb_params&) + mask_rects_ ({x_min,x_max,y_min,y_max} in
 * [0,1] each, num_mask_rects*4 floats, may be NULL).  `device` is the CUDA ordinal. */
int ovs_extractor_create(const ovs_orb_params* params, const float* mask_rects, int num_mask_rects,
                         int device, ovs_extractor** out);
void ovs_extractor_destroy(ovs_extractor* h);

/* Upper bound on the number of keypoints one extract() can return for this handle
 * (max_num_keypts + 3 per level: the tree distribution may overshoot by < 4 per level). */
int ovs_extractor_max_keypoints(const ovs_extractor* h);

/* orb_extractor::extract(in_image, in_image_mask, keypts, out_descriptors), HOST buffers.
 *  image: CV_8UC1, `pitch` bytes per row.  mask: CV_8UC1 same size or NULL (0 = masked out).
 *  keypts_out[capacity], descriptors_out[capacity*32]; *num_out receives the count. */
int ovs_extract_host(ovs_extractor* h, const uint8_t* image, int width, int height, size_t pitch,
                     const uint8_t* mask, size_t mask_pitch,
                     ovs_keypoint* keypts_out, uint8_t* descriptors_out, int capacity, int* num_out);

/* util::convert_to_grayscale(img, color_order) + extract() in one call: `image` is CV_8UC3 or CV_8UC4
 * (`channels`), `pitch` bytes per row, channel order OVS_COLOR_ORDER_BGR (BGR / BGRA) or OVS_COLOR_ORDER_RGB (RGB / RGBA);
 * the gray conversion is cv::cvtColor's 15-bit fixed point (OpenCV 4), bit-exact, done on the device. */
#define OVS_COLOR_ORDER_BGR 0
#define OVS_COLOR_ORDER_RGB 1
int ovs_extract_host_color(ovs_extractor* h, const uint8_t* image, int width, int height, size_t pitch, int channels, int color_order,
                           const uint8_t* mask, size_t mask_pitch,
                           ovs_keypoint* keypts_out, uint8_t* descriptors_out, int capacity, int* num_out);

/* Same, DEVICE image in / DEVICE keypoints + descriptors out (they stay resident for the
 * matchers).  The mask, if any, is still a HOST buffer (it only drives host-side cell and
 * keypoint filtering, as in the reference). */
int ovs_extract_device(ovs_extractor* h, const uint8_t* d_image, int width, int height, size_t pitch,
                       const uint8_t* mask, size_t mask_pitch,
                       ovs_keypoint* d_keypts_out, uint8_t* d_descriptors_out, int capacity, int* num_out);

/* orb_extractor::image_pyramid_ (public member read by match::stereo): geometry and device
 * pointer of level `level` of the last extract(); and a host copy. */
int ovs_extractor_pyramid_level(const ovs_extractor* h, int level, const uint8_t** d_ptr, size_t* pitch,
                                int* width, int* height);
int ovs_extractor_copy_pyramid_level(ovs_extractor* h, int level, uint8_t* out, size_t out_pitch);
/* orb_extractor::scale_factors_ etc. (filled by orb_params::calc_scale_factors). */
int ovs_extractor_scale_factors(const ovs_extractor* h, float* scale_factors, float* inv_scale_factors,
                                float* level_sigma_sq, float* inv_level_sigma_sq);

/* Stage taps for the parity tests (device -> host copies of intermediate results of the last
 * extract()): FAST score map of a level (0 where score < min_fast_thr), and the candidate list
 * (x, y relative to the 19 px border, score) that went into the tree distribution. */
int ovs_extractor_debug_score_map(ovs_extractor* h, int level, uint8_t* out, size_t out_pitch);
int ovs_extractor_debug_candidates(ovs_extractor* h, int level, int32_t* xys_out /* [cap*3] */, int cap, int* n_out);
/* Per-stage device time of the last extract() in microseconds (CUDA events):
 * [0] upload [1] pyramid [2] fast score [3] cell nms+compact [4] host tree distribution (wall)
 * [5] orientation+descriptor [6] download [7] total wall. */
int ovs_extractor_last_timings(const ovs_extractor* h, float* out_us /* [8] */);

/* ------------------------------------------------------------------------------- match::* */

/* match::base constants (match/base.h). */
#define OVS_HAMMING_DIST_THR_LOW 50
#define OVS_HAMMING_DIST_THR_HIGH 100
#define OVS_MAX_HAMMING_DIST 256

/* Length of the per-query candidate lists of the brute-force search. */
#define OVS_MATCH_TOPK 8

typedef struct ovs_matcher ovs_matcher;
int ovs_matcher_create(int device, ovs_matcher** out);
void ovs_matcher_destroy(ovs_matcher* h);

/* Hamming brute force (the inner double loop of match::robust::brute_force_match, match/robust.cc,
 * with match::compute_descriptor_distance_32, match/base.h): for each of the nq query descriptors
 * the OVS_MATCH_TOPK (8) smallest keys (distance << 16 | train index) over the nt train descriptors, ascending, i.e.
 * exactly the order a sequential `<` scan ranks them (lowest index wins ties).  Missing entries
 * (nt < 8) are 0xFFFFFFFF.  nt must be < 65536.  keys_out[nq * 8].  Descriptors are 32 bytes each,
 * device pointers 16-byte aligned. */
int ovs_match_bruteforce_topk_host(ovs_matcher* h, const uint8_t* query, int nq, const uint8_t* train, int nt,
                                   uint32_t* keys_out);
int ovs_match_bruteforce_topk_device(ovs_matcher* h, const uint8_t* d_query, int nq, const uint8_t* d_train, int nt,
                                     uint32_t* d_keys_out);
/* match::robust::match_for_triangulation(keyfrm_1, keyfrm_2, E_12, matched_idx_pairs) (match/robust.cc) on plain arrays.
 * The BoW feature vectors are inputs: bow_node_k[i] = vocabulary node of keypoint i of keyframe k (keyfrm->bow_feat_vec_
 * inverted; < 0 = none).  bearing_k[n*3] = keyfrm->bearings_ (f64), octave_1 / angle_k = undist_keypts_, has_lm_k[i] =
 * keyfrm->get_landmark(i) != nullptr, is_stereo_k[i] = stereo_x_right_[i] >= 0 (NULL = monocular), E_12[9] row-major,
 * epipole_in_2[3] = camera centre of keyframe 1 reprojected to a bearing of keyframe 2, scale_factors_1 = keyfrm_1->
 * scale_factors_.  Candidates are the keyframe-2 keypoints of the same node without a landmark, distance <=
 * HAMMING_DIST_THR_LOW, not within cos 0.998 of the epipole (monocular pairs), inside 0.2 deg x scale of the epipolar
 * plane (check_epipolar_constraint); first taker keeps a keyframe-2 keypoint; orientation histogram if requested.
 * matched_idx_2_of_1[n1] = keypoint of keyframe 2 or -1; *num_matches = the reference's return value. */
int ovs_robust_match_for_triangulation_host(ovs_matcher* m, int n1, const uint8_t* desc_1, const double* bearing_1, const int32_t* octave_1,
                                            const float* angle_1, const uint8_t* has_lm_1, const uint8_t* is_stereo_1, const int32_t* bow_node_1,
                                            int n2, const uint8_t* desc_2, const double* bearing_2, const float* angle_2, const uint8_t* has_lm_2,
                                            const uint8_t* is_stereo_2, const int32_t* bow_node_2, const double* E_12, const double* epipole_in_2,
                                            const float* scale_factors_1, int num_scale_levels, int check_orientation,
                                            int32_t* matched_idx_2_of_1, int* num_matches);

/* Convenience view of the same search: best index (-1 if none), best and second-best distance
 * (OVS_MAX_HAMMING_DIST when absent) of desc1[i] over desc2. */
int ovs_match_bruteforce_host(ovs_matcher* h, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2,
                              int32_t* best_idx, int32_t* best_dist, int32_t* second_dist);

/* match::robust::brute_force_match(frm, keyfrm, matches) (match/robust.cc) on plain arrays:
 *  desc_frm[n1*32]: frm.descriptors_;  desc_keyfrm[n2*32]: keyfrm->descriptors_;
 *  lm_valid_2[n2]: 1 where keyfrm->get_landmarks()[idx_2] is non-null and not will_be_erased()
 *  (NULL = all valid);  lowe_ratio: robust::lowe_ratio_.
 * Output pairs (idx_1 in frame, idx_2 in keyframe) in the reference's emission order (ascending
 * idx_2), with its greedy "a frame keypoint is matched at most once" rule.  Returns the count in
 * *num_matches (the reference's return value). */
int ovs_robust_brute_force_match_host(ovs_matcher* h, const uint8_t* desc_frm, int n1, const uint8_t* desc_keyfrm, int n2,
                                      const uint8_t* lm_valid_2, float lowe_ratio,
                                      int32_t* pairs_out, int capacity, int* num_matches);
/* The same with both descriptor sets resident in device memory (16-byte aligned; e.g. the output of ovs_extract_device):
 * only the per-keypoint candidate lists travel to the host for the sequential replay.  lm_valid_2 / pairs_out: host. */
int ovs_robust_brute_force_match_device(ovs_matcher* h, const uint8_t* d_desc_frm, int n1, const uint8_t* d_desc_keyfrm, int n2,
                                        const uint8_t* lm_valid_2, float lowe_ratio,
                                        int32_t* pairs_out, int capacity, int* num_matches);
/* Diagnostic: how many single-query GPU re-searches the greedy replays of this handle have needed. */
int ovs_matcher_num_requeries(const ovs_matcher* h, int* out);
/* Device time (CUDA events, microseconds) of the Hamming kernels of the last call. */
int ovs_matcher_last_kernel_us(const ovs_matcher* h, float* out_us);

/* ---- windowed search: match::projection / match::area (match/projection.cc, match/area.cc) ---- */

/* The part of camera::base that data::frame::get_keypoints_in_cell reads (camera/base.h):
 * img_bounds_.min_x_/min_y_, inv_cell_width_, inv_cell_height_, num_grid_cols_ (64), num_grid_rows_ (48). */
typedef struct {
    float min_x, min_y;
    float inv_cell_width, inv_cell_height;
    int32_t num_grid_cols, num_grid_rows;
} ovs_grid;

/* A frame's keypoints uploaded once and indexed by grid cell (data::assign_keypoints_to_grid,
 * data/common.cc): x/y/octave/angle = undist_keypts_[i].{pt, octave, angle}, x_right =
 * stereo_x_right_ (NULL: monocular), desc = descriptors_.  Owned by the matcher `m` (its stream). */
typedef struct ovs_frame_index ovs_frame_index;
int ovs_frame_index_create(ovs_matcher* m, int n, const float* x, const float* y, const int32_t* octave, const float* angle,
                           const float* x_right, const uint8_t* desc, const ovs_grid* grid, ovs_frame_index** out);
/* The same index built from the DEVICE output of ovs_extract_device (keypoint records + descriptors, optionally the
 * stereo x_right array, all device pointers; descriptors 16-byte aligned): the descriptors are never copied to the host
 * (data::frame grid + device-resident descriptors).  The matcher's stream must be ordered after the
 * extraction that produced the arrays (ovs_extract_device returns with its stream drained). */
int ovs_frame_index_create_device(ovs_matcher* m, int n, const ovs_keypoint* d_keypts, const uint8_t* d_desc, const float* d_x_right,
                                  const ovs_grid* grid, ovs_frame_index** out);
void ovs_frame_index_destroy(ovs_frame_index* f);

/* frame::get_keypoints_in_cell(ref_x, ref_y, margin, min_level, max_level) followed by the nearest
 * descriptor search every projection matcher performs, for nq queries at once: the 4 best candidates
 * of each query (index into the frame's keypoints and Hamming distance; -1 / 256 where absent) in
 * the reference's candidate order (ties: first visited wins).  x_right_q may be NULL. */
int ovs_match_window_topk_host(ovs_frame_index* f, int nq, const float* ref_xy, const float* margin, const int32_t* min_level,
                               const int32_t* max_level, const float* x_right_q, const uint8_t* qdesc,
                               int32_t* idx_out, int32_t* dist_out);

/* match::projection::match_frame_and_landmarks(frm, local_landmarks, margin):
 *  lm_usable[l] = is_observable_in_tracking_ && !will_be_erased();  reproj_xy / x_right_in_tracking /
 *  pred_scale_level = the landmark's reproj_in_tracking_, x_right_in_tracking_, scale_level_in_tracking_;
 *  lm_desc = get_descriptor();  kp_has_observed_lm[i] = frm.landmarks_[i] && has_observation().
 * matched_lm_of_kp[i] receives the landmark index assigned to frm.landmarks_[i] by this call (-1: none). */
int ovs_projection_match_frame_and_landmarks_host(ovs_frame_index* f, const float* scale_factors, int num_scale_levels, int nlm, const uint8_t* lm_usable,
                                                  const float* reproj_xy, const float* x_right_in_tracking,
                                                  const int32_t* pred_scale_level, const uint8_t* lm_desc,
                                                  const uint8_t* kp_has_observed_lm, float margin, float lowe_ratio,
                                                  int32_t* matched_lm_of_kp, int* num_matches);

/* match::projection::match_current_and_last_frames(curr_frm, last_frm, margin): one entry per keypoint
 * of the last frame.  last_usable[i] = it has a landmark, is not an outlier and reprojects into the
 * current image (camera->reproject_to_image, evaluated by the caller as in the reference);
 * reproj_xy / reproj_x_right = that reprojection; last_scale_level / last_angle = last_frm keypoint
 * octave / undistorted angle; lm_desc = landmark descriptors.  assume_forward / assume_backward as
 * computed from trans_lc.  matched_last_of_kp[i] = index in the last frame or -1. */
int ovs_projection_match_current_and_last_host(ovs_frame_index* curr, const float* scale_factors, int num_scale_levels, int n_last,
                                               const uint8_t* last_usable, const float* reproj_xy, const float* reproj_x_right,
                                               const int32_t* last_scale_level, const float* last_angle, const uint8_t* lm_desc,
                                               const uint8_t* kp_has_observed_lm, float margin, int assume_forward, int assume_backward,
                                               int check_orientation, int32_t* matched_last_of_kp, int* num_matches);

/* The search loop shared by projection::match_current_and_last_frames, match_frame_and_keyframe,
 * match_by_Sim3_transform and each direction of match_keyframes_mutually (match/projection.cc): one
 * query per reprojected landmark -- usable[i] (NULL = all), reprojection ref_xy, optional reprojected
 * x_right (NULL: the x_right test is not part of the matcher), search margin (already multiplied by
 * the scale factor of the predicted level), level range [min_level, max_level] (max < 0: unbounded),
 * descriptor, keypoint/landmark angle (for the orientation check).  kp_unavailable[i] marks frame
 * keypoints that must not be matched (already associated).  A query takes its nearest available
 * keypoint when the distance is <= hamm_dist_thr; queries are served in index order and a keypoint is
 * given to the first taker, as in the reference loops.  matched_query_of_kp[i] = query index or -1. */
int ovs_projection_match_best_host(ovs_frame_index* f, int nq, const uint8_t* usable, const float* ref_xy, const float* ref_x_right,
                                   const float* margin, const int32_t* min_level, const int32_t* max_level, const float* q_angle,
                                   const uint8_t* q_desc, const uint8_t* kp_unavailable, unsigned hamm_dist_thr, int check_orientation,
                                   int32_t* matched_query_of_kp, int* num_matches);

/* match::projection::match_keyframes_mutually(keyfrm_1, keyfrm_2, matched_lms_in_keyfrm_2, Sim3_12, Sim3_21, margin)
 * (match/projection.cc, loop closure).  Landmark arrays are indexed by the keypoint of the keyframe they belong to
 * (keyfrm->landmarks_): usable_k[i] = landmark present, not bad, not already matched, reprojection inside the other
 * image; reproj_a_in_b / pred_level_a_in_b = its reprojection with the Sim3 and predict_scale_level(); lm_desc = the
 * landmark's representative descriptor.  Each landmark takes its nearest keypoint of the other keyframe in the window
 * margin * scale_factors[level], levels [level - 1, level], distance <= HAMMING_DIST_THR_HIGH; pairs on which both
 * directions agree are returned: matched_idx_2_of_kp_1[i1] = keypoint of keyframe 2 or -1 (f1->n entries). */
int ovs_projection_match_keyframes_mutually_host(ovs_frame_index* f1, ovs_frame_index* f2, const float* scale_factors,
                                                 const uint8_t* usable_1, const float* reproj_1_in_2, const int32_t* pred_level_1_in_2,
                                                 const uint8_t* lm_desc_1, const uint8_t* usable_2, const float* reproj_2_in_1,
                                                 const int32_t* pred_level_2_in_1, const uint8_t* lm_desc_2, float margin,
                                                 int32_t* matched_idx_2_of_kp_1, int* num_matches);

/* match::area::match_in_consistent_area(frm_1, frm_2, prev_matched_pts, matched_indices_2_in_frm_1, margin):
 * f2 indexes frm_2; octave_1 / angle_1 / desc_1 describe frm_1's keypoints; prev_matched_xy[n1*2] is
 * updated in place. */
int ovs_area_match_in_consistent_area_host(ovs_frame_index* f2, int n1, const int32_t* octave_1, const float* angle_1, const uint8_t* desc_1,
                                           float* prev_matched_xy, int32_t* matched_idx_2_in_1, int margin, float lowe_ratio,
                                           int check_orientation, int* num_matches);

/* match::stereo(left_image_pyramid, right_image_pyramid, keypts_left, keypts_right, descs_left,
 * descs_right, scale_factors, inv_scale_factors, focal_x_baseline, true_baseline)
 *   .compute(stereo_x_right, depths)  (match/stereo.cc).
 * The pyramids are read from the two extractors (device resident, after their extract() of the pair). */
int ovs_stereo_compute_host(ovs_matcher* m, const ovs_extractor* left, const ovs_extractor* right,
                            int n_left, const float* lx, const float* ly, const int32_t* loct, const uint8_t* ldesc,
                            int n_right, const float* rx, const float* ry, const int32_t* roct, const uint8_t* rdesc,
                            float focal_x_baseline, float true_baseline, float* stereo_x_right, float* depths, int* num_matched);

/* ------------------------------------------------------------------------------ optimize::* */

#define OVS_CAMERA_PERSPECTIVE 0       /* camera::perspective (and fisheye: same edges on undistorted keypoints) */
#define OVS_CAMERA_EQUIRECTANGULAR 1   /* camera::equirectangular */
#define OVS_CAMERA_FISHEYE 2           /* camera::fisheye -- ovs_undistort_keypoints_* only (matchers / optimisers see it as perspective) */
#define OVS_CAMERA_RADIAL_DIVISION 3   /* camera::radial_division -- ovs_undistort_keypoints_* only */

/* The camera parameters the reprojection edges read (optimize/g2o/se3/ *_edge.h): fx_, fy_, cx_, cy_,
 * focal_x_baseline_ (stereo / RGBD), cols_, rows_ (equirectangular). */
typedef struct {
    int32_t model;
    double fx, fy, cx, cy, focal_x_baseline;
    double cols, rows;
} ovs_camera;

/* data::frame's constructor right after extract(): camera->undistort_keypoints(keypts_, undist_keypts_) +
 * camera->convert_keypoints_to_bearings(undist_keypts_, bearings_) (camera/perspective.cc, camera/equirectangular.cc; SURVEY 8f
 * rank 3).  Perspective: cv::undistortPoints(pts, K, dist, R = I, P = K, MAX_ITER num_iterations) -- OpenVSLAM uses 20 --
 * with dist = {k1, k2, p1, p2, k3} (NULL = no distortion), bit-exact with OpenCV in the float keypoints; bearings[n*3] f64.
 * Equirectangular: keypoints unchanged, bearings from longitude / latitude.  Only pt changes in the keypoint records.
 * Fisheye (camera/fisheye.cc): cv::fisheye::undistortPoints(pts, K, D, R = I, P = K) with its default criteria -- the `dist`
 * argument then holds D = (k1, k2, k3, k4) and num_iterations the Newton step limit (OpenCV: 10); points that do not converge
 * become (-1e6, -1e6) as in OpenCV >= 4.5.  Radial division (camera/radial_division.cc): p_u = p_d / (1 + dist[0] |p_d|^2) on
 * normalised coordinates.  Both take perspective bearings of the undistorted keypoints.
 * The _device variant works on the extractor's device output (d_undist_out may alias d_keypts_in; outputs may be NULL). */
int ovs_undistort_keypoints_device(ovs_extractor* h, const ovs_camera* cam, const double* dist_k1k2p1p2k3, int num_iterations, int n,
                                   const ovs_keypoint* d_keypts_in, ovs_keypoint* d_undist_out, double* d_bearings_out);
int ovs_undistort_keypoints_host(ovs_extractor* h, const ovs_camera* cam, const double* dist_k1k2p1p2k3, int num_iterations, int n,
                                 const ovs_keypoint* keypts_in, ovs_keypoint* undist_out, double* bearings_out);

typedef struct {
    int32_t num_rounds;            /* optimizer.optimize() calls made */
    int32_t num_iterations;        /* Levenberg iterations executed in total */
    int32_t num_trials;            /* linear solves (LM trials) in total */
    int32_t round_iterations[8];
    double lambda_init[8];         /* computeLambdaInit() of each round */
    double last_lambda, last_chi2; /* of the last round */
    double final_chi2;             /* robust chi2 of the active edges at the returned state */
    float device_us;               /* CUDA-event time of the call's device work */
    float solver_us;               /* local BA: CUDA-event time (launching stream) of the reduced-system solver launches, summed */
    int32_t solver_launches;       /* local BA: launches of the reduced-system solver (one per Levenberg iteration) */
    int32_t solver_trials;         /* local BA: systems factorised (speculative damping trials, <= 4 per launch) */
    int32_t reduced_dim;           /* local BA: dimension of the reduced camera system (6 x free keyframes) */
    float schur_us;                /* local BA: CUDA-event time of the Schur-complement launches (chunk + final) of the batches that ran, summed */
    int32_t co_observations;       /* local BA: (landmark, keyframe pair a <= b) records the Schur complement sums over */
} ovs_ba_stats;

typedef struct ovs_optimizer ovs_optimizer;
int ovs_optimizer_create(int device, ovs_optimizer** out);
void ovs_optimizer_destroy(ovs_optimizer* h);

/* pose_optimizer::optimize(data::frame& frm) (optimize/pose_optimizer.cc) on plain arrays: one SE3
 * vertex, one unary reprojection edge per matched landmark (frm.landmarks_[idx] valid).
 *  setup_is_mono: camera->setup_type_ == Monocular (selects the Huber delta sqrt(5.991) / sqrt(7.815));
 *  pts_w[n*3]: lm->get_pos_in_world();  obs_xy[n*2]: frm.undist_keypts_[idx].pt;
 *  obs_x_right[n]: frm.stereo_x_right_[idx] (< 0 = monocular edge; NULL = all monocular);
 *  inv_sigma_sq[n]: frm.inv_level_sigma_sq_[octave];
 *  pose_cw[12]: frm.cam_pose_cw_ as {R row-major (9), t (3)}, updated in place (frm.set_cam_pose);
 *  outlier_flags[n]: frm.outlier_flags_;  *num_inliers: the return value (num_init_obs - num_bad_obs).
 * num_trials / num_each_iter: the constructor arguments (4, 10).
 * A handle holds ONE problem at a time: this call reuses the handle's device buffers, so a local-BA problem prepared on
 * the same handle (ovs_local_ba_prepare) is invalidated by it -- ovs_local_ba_run / _fetch then fail with
 * OVS_ERR_INVALID_ARG until the problem is prepared again.  Use separate handles to keep both resident. */
int ovs_pose_optimize_host(ovs_optimizer* h, const ovs_camera* cam, int setup_is_mono, int n, const double* pts_w,
                           const float* obs_xy, const float* obs_x_right, const float* inv_sigma_sq,
                           double* pose_cw, uint8_t* outlier_flags, int num_trials, int num_each_iter,
                           int* num_inliers, ovs_ba_stats* stats);

/* local_bundle_adjuster::optimize(curr_keyfrm, force_stop_flag) (optimize/local_bundle_adjuster.cc) on
 * the graph the reference builds: K keyframe vertices (local keyframes free, "fixed" keyframes and
 * keyframe id 0 fixed), L landmark vertices (marginalised), M reprojection edges.
 *  poses[K*12] ({R row-major, t} of cam_pose_cw), points[L*3]: updated in place;
 *  observations grouped by landmark (obs_lm non-decreasing), the order the reference adds them;
 *  outlier_out[M]: 1 where the reference would erase the observation (chi2 over the 5% bound or
 *  non-positive depth after the second round).  force_stop_flag (the reference's `bool* const`, read as
 *  one byte) may be NULL; it is polled between LM trials like g2o's terminate().  num_first_iter / num_second_iter: constructor arguments (5, 10).
 * Up to 114 free keyframes the reduced camera system is factorised by one cluster kernel out of shared memory; larger
 * local maps (up to 1000 free keyframes) take a multi-launch path with the panel in global memory. */
int ovs_local_ba_host(ovs_optimizer* h, const ovs_camera* cam, int setup_is_mono, int K, double* poses, const uint8_t* fixed,
                      int L, double* points, int M, const int32_t* obs_kf, const int32_t* obs_lm, const float* obs_xy,
                      const float* obs_x_right, const float* inv_sigma_sq, int num_first_iter, int num_second_iter,
                      const volatile uint8_t* force_stop_flag, uint8_t* outlier_out, ovs_ba_stats* stats);

/* The same call in three phases, for callers that keep the problem resident on the device:
 * prepare = graph bookkeeping + upload + co-observation lists (enqueued, returns without waiting; the input arrays
 * are copied before it returns); run = the two Levenberg rounds, restarting from the uploaded estimates each time
 * (state stays in HBM; the whole Levenberg loop, accept / reject decisions included, runs on the device: the host enqueues
 * a static launch sequence and waits once); fetch = download of the poses, points and outlier flags of the last run (any
 * output pointer may be NULL). */
int ovs_local_ba_prepare(ovs_optimizer* h, const ovs_camera* cam, int setup_is_mono, int K, const double* poses,
                         const uint8_t* fixed, int L, const double* points, int M, const int32_t* obs_kf,
                         const int32_t* obs_lm, const float* obs_xy, const float* obs_x_right, const float* inv_sigma_sq);
int ovs_local_ba_run(ovs_optimizer* h, int num_first_iter, int num_second_iter, const volatile uint8_t* force_stop_flag,
                     ovs_ba_stats* stats);
int ovs_local_ba_fetch(ovs_optimizer* h, double* poses, double* points, uint8_t* outlier_out);
/* The same with the graph already resident in device memory (all array arguments are device pointers, d_obs_x_right may be
 * NULL): what a caller that keeps its map on the GPU uses -- no host loop touches the observations, the graph bookkeeping
 * (free-keyframe ids, per-landmark edge ranges, validation) runs on the device.  The arrays are copied before return. */
int ovs_local_ba_prepare_device(ovs_optimizer* h, const ovs_camera* cam, int setup_is_mono, int K, const double* d_poses,
                                const uint8_t* d_fixed, int L, const double* d_points, int M, const int32_t* d_obs_kf,
                                const int32_t* d_obs_lm, const float* d_obs_xy, const float* d_obs_x_right, const float* d_inv_sigma_sq);
int ovs_local_ba_fetch_device(ovs_optimizer* h, double* d_poses, double* d_points, uint8_t* d_outlier_out);
/* optimize::global_bundle_adjuster::optimize(lead_keyfrm_id_in_global_BA, force_stop_flag) (optimize/global_bundle_adjuster.cc)
 * on the graph the reference builds: every keyframe (the origin keyframe fixed: fixed[k] != 0) and every landmark of the map,
 * one reprojection edge per observation, ONE Levenberg round of num_iter (constructor argument, 10) iterations, Huber kernel
 * on every edge when use_huber_kernel (constructor argument, true); no outlier classification.  poses / points are updated in
 * place (the reference stores them as pose_cw_after_loop_BA_ / pos_w_after_global_BA_).  Same array conventions as
 * ovs_local_ba_host; up to 1000 free keyframes (beyond 114 the reduced system takes the multi-launch solver). */
int ovs_global_ba_host(ovs_optimizer* h, const ovs_camera* cam, int setup_is_mono, int K, double* poses, const uint8_t* fixed,
                       int L, double* points, int M, const int32_t* obs_kf, const int32_t* obs_lm, const float* obs_xy,
                       const float* obs_x_right, const float* inv_sigma_sq, int num_iter, int use_huber_kernel,
                       const volatile uint8_t* force_stop_flag, ovs_ba_stats* stats);
/* Development aid: SM clock stamps of the phases of the last reduced-system factorisation (192 values). */
int ovs_optimizer_debug_clocks(ovs_optimizer* h, long long* out192);
/* Test hook: the stable radix sort of the BA graph preparation (co-observations by keyframe pair: k_sort_hist /
 * k_sort_tile_prefix / k_sort_scatter) on host arrays, by the low end_bit bits of the keys. */
int ovs_debug_sort_pairs(int device, const uint32_t* keys, const uint64_t* vals, int n, int end_bit, uint32_t* keys_out, uint64_t* vals_out);
/* CTAs per thread-block cluster of the reduced-system solver on this device (8, or 16 when 4 such clusters can be co-resident). */
int ovs_optimizer_cluster_width(const ovs_optimizer* h);
/* Local / global BA: CTAs per cluster of the reduced-system solver (1, 2, 4 or 8; default 8).  8 gives the lowest latency of a
 * single call; 2 gives the most calls per second when several optimisers share the GPU (measured on B200, 8 concurrent local
 * BAs: +13 % calls/s, +8 % latency per call).  The result does not depend on the width. */
int ovs_optimizer_set_cluster_width(ovs_optimizer* h, int width);
/* Local BA: the launch sequence of one Levenberg iteration is static (damping values, ring slots and the accept / reject
 * walk live in device memory), so it can be captured once per run and replayed as ONE CUDA graph per iteration.  Trims
 * the inter-kernel gaps of a single stream; off by default. */
int ovs_optimizer_set_graphs(ovs_optimizer* h, int enable);
/* Local BA: the Levenberg loop runs on the device without host round trips.  mode 1 makes the host read the device's
 * decision after every trial batch and skip the launches that are not needed (default only for reduced systems beyond
 * the cluster solver, whose batches are ~100 launches each); mode 0 never synchronises; -1 = automatic.  Same results. */
int ovs_optimizer_set_host_sync(ovs_optimizer* h, int mode);
/* Local BA: number of Levenberg damping trials evaluated speculatively per launch sequence (1..4, default 4).  The
 * result is the sequential algorithm's for every width; 4 minimises the latency of one session (a rejected trial costs
 * no extra round trip) and, measured on B200, is also the fastest setting with 8 sessions per GPU; smaller widths
 * trade latency for less speculative GPU work. */
int ovs_optimizer_set_speculation(ovs_optimizer* h, int width);
/* Local BA: a second trial batch of `width` (1..4) damping values enqueued statically behind the first batch of every
 * iteration (0 = off, the default).  Its kernels return at their first instruction when the first batch decided the
 * iteration, so e.g. speculation 2 + second batch 2 evaluates 2 trials where g2o's loop needs <= 2 and 4 where it needs
 * 3 or 4, still without a host round trip: less speculative GPU work per call (throughput when several sessions share the
 * GPU) for one more dependent launch sequence on the iterations that reject their first trials (latency).  Same results. */
int ovs_optimizer_set_second_batch(ovs_optimizer* h, int width);

/* match::bow_tree::match_frame_and_keyframe(keyfrm, frm, matched_lms_in_frm) (match/bow_tree.cc) on plain arrays.  The BoW
 * feature vectors are inputs: bow_node_x[i] = vocabulary node of keypoint i (< 0 = none).  Nodes ascending, keypoints of a
 * node in index order (the reference's lock-step walk over the two feature vectors): every keyframe keypoint with a valid
 * landmark (lm_valid_kf) takes its nearest still-unmatched frame keypoint of the same node when the distance is <=
 * HAMMING_DIST_THR_LOW and lowe_ratio * second_best >= best; orientation histogram if requested.
 * matched_keyfrm_idx_of_frm[n_frm] = keyframe keypoint whose landmark frame keypoint i receives, or -1. */
int ovs_bow_tree_match_frame_and_keyframe_host(ovs_matcher* m, int n_kf, const uint8_t* desc_kf, const float* angle_kf, const uint8_t* lm_valid_kf,
                                               const int32_t* bow_node_kf, int n_frm, const uint8_t* desc_frm, const float* angle_frm,
                                               const int32_t* bow_node_frm, float lowe_ratio, int check_orientation,
                                               int32_t* matched_keyfrm_idx_of_frm, int* num_matches);
/* match::bow_tree::match_keyframes(keyfrm_1, keyfrm_2, matched_lms_in_keyfrm_1): both keypoints need a valid landmark, a
 * keyframe-2 keypoint is matched at most once.  matched_idx_2_of_1[n1] = keypoint of keyframe 2 or -1. */
int ovs_bow_tree_match_keyframes_host(ovs_matcher* m, int n1, const uint8_t* desc_1, const float* angle_1, const uint8_t* lm_valid_1,
                                      const int32_t* bow_node_1, int n2, const uint8_t* desc_2, const float* angle_2, const uint8_t* lm_valid_2,
                                      const int32_t* bow_node_2, float lowe_ratio, int check_orientation,
                                      int32_t* matched_idx_2_of_1, int* num_matches);
/* match::fuse (match/fuse.cc): the matching core of replace_duplication / detect_duplication.  Every usable landmark
 * (reprojected into the keyframe `f` by the caller: reproj_xy[nq*2], reproj_x_right[nq] or NULL, pred_level[nq] from
 * predict_scale_level, lm_desc[nq*32]) searches the window margin * scale_factors[level], levels [level - 1, level];
 * candidates whose reprojection error times inv_level_sigma_sq[own octave] exceeds 5.99 (7.8 with the x_right term) are
 * skipped; nearest descriptor, first in visiting order on ties, accepted at <= HAMMING_DIST_THR_LOW.  best_idx_of_lm[nq] =
 * keypoint index or -1 (what happens to a keypoint that already holds a landmark is the caller's data-model decision). */
int ovs_fuse_best_keypoints_host(ovs_frame_index* f, int nq, const uint8_t* usable, const float* reproj_xy, const float* reproj_x_right,
                                 const int32_t* pred_level, const uint8_t* lm_desc, const float* scale_factors,
                                 const float* inv_level_sigma_sq, int num_scale_levels, float margin,
                                 int32_t* best_idx_of_lm, int* num_matches);

/* Measured FP64 peaks of `device` (whole chip, TFLOP/s counting 2 per FMA): independent mma.sync.m8n8k4.f64 (DMMA) and
 * independent DFMA.  bench.py quotes the Cholesky roofline against the DMMA figure measured in the same run. */
int ovs_probe_fp64_peaks(int device, double* dmma_tflops, double* dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* OVS_B200_H */
