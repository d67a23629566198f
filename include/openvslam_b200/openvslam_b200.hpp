// openvslam_b200.hpp -- the reference's hot-path class surfaces, re-declared over the C ABI of
// libovs_b200.so (include/ovs_b200.h).  Header only.
//
//   openvslam::feature::orb_params / orb_extractor        (src/openvslam/feature/orb_params.h, orb_extractor.h)
//   openvslam::match::robust / projection / area / stereo (src/openvslam/match/*.h)
//   openvslam::optimize::pose_optimizer / local_bundle_adjuster (src/openvslam/optimize/*.h)
// [file names as recalled in SURVEY.md 8(a); /root/reference holds no source, so no line numbers].
//
// The reference's methods take cv::Mat / cv::KeyPoint / Eigen / data::frame / data::keyframe.  None of
// those headers exist in this build environment, so every method is declared on plain views
// (pointers + sizes, `ovs_keypoint` which is layout-compatible with cv::KeyPoint).  When
// OVS_B200_WITH_OPENCV is defined (a tree that has OpenCV), the cv::_InputArray overloads with the
// reference's exact signatures are compiled as well and forward to the same code.
// The data::frame / data::keyframe flattening the matchers and optimisers need is the job of the
// adapter shown in INTEGRATION.md; the classes here keep the reference's constructor arguments,
// member names and return values.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../ovs_b200.h"

#ifdef OVS_B200_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace openvslam {

// The reference's data model (data/frame.h, data/keyframe.h, data/landmark.h).  With OVS_B200_WITH_REFERENCE_TYPES the classes
// below also declare the reference's own method signatures on these types; their bodies -- the flattening of the data model
// into the array views -- are in adapters.hpp, which a tree that has those headers includes instead of this file.
namespace data { class frame; class keyframe; class landmark; }

namespace detail {
inline void check(int rc) {
    if (rc != OVS_OK) throw std::runtime_error(std::string("ovs_b200: ") + ovs_last_error());
}
}  // namespace detail

namespace feature {

struct orb_params {
    orb_params() = default;
    orb_params(const unsigned int max_num_keypts, const float scale_factor, const unsigned int num_levels,
               const unsigned int ini_fast_thr, const unsigned int min_fast_thr,
               const std::vector<std::vector<float>>& mask_rects = {})
        : max_num_keypts_(max_num_keypts), scale_factor_(scale_factor), num_levels_(num_levels),
          ini_fast_thr_(ini_fast_thr), min_fast_thr(min_fast_thr), mask_rects_(mask_rects) {
        for (const auto& v : mask_rects_) {
            if (v.size() != 4) throw std::runtime_error("Each of mask rectangles must contain four parameters");
            if (v.at(0) >= v.at(1)) throw std::runtime_error("x_max must be greater than x_min");
            if (v.at(2) >= v.at(3)) throw std::runtime_error("y_max must be greater than x_min");
        }
    }
    unsigned int max_num_keypts_ = 2000;
    float scale_factor_ = 1.2f;
    unsigned int num_levels_ = 8;
    unsigned int ini_fast_thr_ = 20;
    unsigned int min_fast_thr = 7;
    //! A vector of keypoint area represents mask area: {x_min / cols, x_max / cols, y_min / rows, y_max / rows}
    std::vector<std::vector<float>> mask_rects_;
};

class orb_extractor {
public:
    orb_extractor() = delete;
    explicit orb_extractor(const orb_params& orb_params, const int device = 0) : orb_params_(orb_params) {
        ovs_orb_params p{orb_params.max_num_keypts_, orb_params.scale_factor_, orb_params.num_levels_, orb_params.ini_fast_thr_,
                         orb_params.min_fast_thr};
        std::vector<float> rects;
        for (const auto& r : orb_params.mask_rects_) rects.insert(rects.end(), r.begin(), r.end());
        detail::check(ovs_extractor_create(&p, rects.empty() ? nullptr : rects.data(), static_cast<int>(rects.size() / 4), device, &h_));
        const unsigned int L = orb_params.num_levels_;
        scale_factors_.resize(L); inv_scale_factors_.resize(L); level_sigma_sq_.resize(L); inv_level_sigma_sq_.resize(L);
        detail::check(ovs_extractor_scale_factors(h_, scale_factors_.data(), inv_scale_factors_.data(), level_sigma_sq_.data(),
                                                  inv_level_sigma_sq_.data()));
    }
    orb_extractor(const unsigned int max_num_keypts, const float scale_factor, const unsigned int num_levels,
                  const unsigned int ini_fast_thr, const unsigned int min_fast_thr,
                  const std::vector<std::vector<float>>& mask_rects = {})
        : orb_extractor(orb_params{max_num_keypts, scale_factor, num_levels, ini_fast_thr, min_fast_thr, mask_rects}) {}
    ~orb_extractor() { ovs_extractor_destroy(h_); }
    orb_extractor(const orb_extractor&) = delete;
    orb_extractor& operator=(const orb_extractor&) = delete;

    //! Extract keypoints and each descriptor of them (image: CV_8UC1 rows x cols, `step` bytes per row;
    //! mask: same size or nullptr).  descriptors: keypts.size() x 32 bytes.
    void extract(const std::uint8_t* image, const int rows, const int cols, const std::size_t step, const std::uint8_t* mask,
                 const std::size_t mask_step, std::vector<ovs_keypoint>& keypts, std::vector<std::uint8_t>& descriptors) {
        keypts.clear(); descriptors.clear();
        if (!image || rows <= 0 || cols <= 0) return;
        const int cap = ovs_extractor_max_keypoints(h_);
        keypts.resize(cap); descriptors.resize(static_cast<std::size_t>(cap) * 32);
        int n = 0;
        detail::check(ovs_extract_host(h_, image, cols, rows, step, mask, mask_step, keypts.data(), descriptors.data(), cap, &n));
        keypts.resize(n); descriptors.resize(static_cast<std::size_t>(n) * 32);
    }

    //! util::convert_to_grayscale(img, color_order) + extract() in one call: image is CV_8UC3 / CV_8UC4 (`channels`),
    //! color_order = OVS_COLOR_ORDER_BGR or OVS_COLOR_ORDER_RGB (camera::color_order_t)
    void extract_color(const std::uint8_t* image, const int rows, const int cols, const std::size_t step, const int channels, const int color_order,
                       const std::uint8_t* mask, const std::size_t mask_step, std::vector<ovs_keypoint>& keypts,
                       std::vector<std::uint8_t>& descriptors) {
        keypts.clear(); descriptors.clear();
        if (!image || rows <= 0 || cols <= 0) return;
        const int cap = ovs_extractor_max_keypoints(h_);
        keypts.resize(cap); descriptors.resize(static_cast<std::size_t>(cap) * 32);
        int n = 0;
        detail::check(ovs_extract_host_color(h_, image, cols, rows, step, channels, color_order, mask, mask_step, keypts.data(), descriptors.data(),
                                             cap, &n));
        keypts.resize(n); descriptors.resize(static_cast<std::size_t>(n) * 32);
    }

    //! camera->undistort_keypoints(keypts, undist_keypts) + camera->convert_keypoints_to_bearings(undist_keypts, bearings):
    //! dist = {k1, k2, p1, p2, k3} or nullptr; bearings = 3 doubles per keypoint
    void undistort_keypoints(const ovs_camera& camera, const double* dist, const std::vector<ovs_keypoint>& keypts,
                             std::vector<ovs_keypoint>& undist_keypts, std::vector<double>& bearings, const int num_iterations = 20) const {
        undist_keypts.resize(keypts.size()); bearings.resize(keypts.size() * 3);
        detail::check(ovs_undistort_keypoints_host(h_, &camera, dist, num_iterations, static_cast<int>(keypts.size()), keypts.data(),
                                                   undist_keypts.data(), bearings.data()));
    }

#ifdef OVS_B200_WITH_OPENCV
    //! The reference's signature.
    void extract(const cv::_InputArray& in_image, const cv::_InputArray& in_image_mask, std::vector<cv::KeyPoint>& keypts,
                 const cv::_OutputArray& out_descriptors) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(ovs_keypoint), "cv::KeyPoint layout changed");
        if (in_image.empty()) return;
        const cv::Mat image = in_image.getMat();
        CV_Assert(image.type() == CV_8UC1);
        const cv::Mat mask = in_image_mask.empty() ? cv::Mat() : in_image_mask.getMat();
        const int cap = ovs_extractor_max_keypoints(h_);
        keypts.resize(cap);
        cv::Mat desc(cap, 32, CV_8U);
        int n = 0;
        detail::check(ovs_extract_host(h_, image.data, image.cols, image.rows, image.step, mask.empty() ? nullptr : mask.data,
                                       mask.empty() ? 0 : mask.step, reinterpret_cast<ovs_keypoint*>(keypts.data()), desc.data, cap, &n));
        keypts.resize(n);
        if (n == 0) out_descriptors.release(); else desc.rowRange(0, n).copyTo(out_descriptors);
    }
#endif

    unsigned int get_max_num_keypoints() const { return orb_params_.max_num_keypts_; }
    float get_scale_factor() const { return orb_params_.scale_factor_; }
    unsigned int get_num_scale_levels() const { return orb_params_.num_levels_; }
    unsigned int get_initial_fast_threshold() const { return orb_params_.ini_fast_thr_; }
    unsigned int get_minimum_fast_threshold() const { return orb_params_.min_fast_thr; }
    std::vector<float> get_scale_factors() const { return scale_factors_; }
    std::vector<float> get_inv_scale_factors() const { return inv_scale_factors_; }
    std::vector<float> get_level_sigma_sq() const { return level_sigma_sq_; }
    std::vector<float> get_inv_level_sigma_sq() const { return inv_level_sigma_sq_; }

    //! image_pyramid_ (read by match::stereo): host copy of one level of the last extract()
    std::vector<std::uint8_t> image_pyramid(const int level, int& rows, int& cols) const {
        detail::check(ovs_extractor_pyramid_level(h_, level, nullptr, nullptr, &cols, &rows));
        std::vector<std::uint8_t> out(static_cast<std::size_t>(rows) * cols);
        detail::check(ovs_extractor_copy_pyramid_level(h_, level, out.data(), static_cast<std::size_t>(cols)));
        return out;
    }

    ovs_extractor* handle() const { return h_; }

private:
    orb_params orb_params_;
    ovs_extractor* h_ = nullptr;
    std::vector<float> scale_factors_, inv_scale_factors_, level_sigma_sq_, inv_level_sigma_sq_;
};

}  // namespace feature

namespace match {

static constexpr unsigned int HAMMING_DIST_THR_LOW = OVS_HAMMING_DIST_THR_LOW;
static constexpr unsigned int HAMMING_DIST_THR_HIGH = OVS_HAMMING_DIST_THR_HIGH;
static constexpr unsigned int MAX_HAMMING_DIST = OVS_MAX_HAMMING_DIST;

class base {
public:
    base(const float lowe_ratio, const bool check_orientation, const int device = 0)
        : lowe_ratio_(lowe_ratio), check_orientation_(check_orientation) {
        detail::check(ovs_matcher_create(device, &h_));
    }
    virtual ~base() { ovs_matcher_destroy(h_); }
    base(const base&) = delete;
    base& operator=(const base&) = delete;
    ovs_matcher* handle() const { return h_; }

protected:
    const float lowe_ratio_;
    const bool check_orientation_;
    ovs_matcher* h_ = nullptr;
};

//! A data::frame as the matchers see it: undist_keypts_, stereo_x_right_, descriptors_ and the grid.
struct frame_view {
    int num_keypts = 0;
    const float* x = nullptr; const float* y = nullptr; const std::int32_t* octave = nullptr; const float* angle = nullptr;
    const float* stereo_x_right = nullptr;   // nullptr: monocular
    const std::uint8_t* descriptors = nullptr;
    ovs_grid grid{};
};

class frame_index {
public:
    frame_index(const base& m, const frame_view& f) : n_(f.num_keypts) {
        detail::check(ovs_frame_index_create(m.handle(), f.num_keypts, f.x, f.y, f.octave, f.angle, f.stereo_x_right, f.descriptors, &f.grid, &h_));
    }
    //! over the device output of orb_extractor::extract_device: keypoint records and descriptors stay on the GPU
    frame_index(const base& m, const int num_keypts, const ovs_keypoint* d_keypts, const std::uint8_t* d_descriptors, const float* d_stereo_x_right,
                const ovs_grid& grid) : n_(num_keypts) {
        detail::check(ovs_frame_index_create_device(m.handle(), num_keypts, d_keypts, d_descriptors, d_stereo_x_right, &grid, &h_));
    }
    ~frame_index() { ovs_frame_index_destroy(h_); }
    frame_index(const frame_index&) = delete;
    frame_index& operator=(const frame_index&) = delete;
    ovs_frame_index* handle() const { return h_; }
    int num_keypts() const { return n_; }
private:
    ovs_frame_index* h_ = nullptr;
    int n_ = 0;
};

class robust final : public base {
public:
    explicit robust(const float lowe_ratio = 0.6, const bool check_orientation = true, const int device = 0)
        : base(lowe_ratio, check_orientation, device) {}
    //! brute_force_match(frm, keyfrm, matches): matches = (idx in frame, idx in keyframe)
    unsigned int brute_force_match(const std::uint8_t* descs_frm, const int num_keypts_frm, const std::uint8_t* descs_keyfrm,
                                   const int num_keypts_keyfrm, const std::uint8_t* keyfrm_lm_valid,
                                   std::vector<std::pair<int, int>>& matches) const {
        std::vector<std::int32_t> pairs(2 * static_cast<std::size_t>(std::max(1, std::min(num_keypts_frm, num_keypts_keyfrm))));
        int n = 0;
        detail::check(ovs_robust_brute_force_match_host(h_, descs_frm, num_keypts_frm, descs_keyfrm, num_keypts_keyfrm, keyfrm_lm_valid,
                                                        lowe_ratio_, pairs.data(), static_cast<int>(pairs.size() / 2), &n));
        matches.clear();
        for (int i = 0; i < n; ++i) matches.emplace_back(pairs[2 * i], pairs[2 * i + 1]);
        return static_cast<unsigned int>(n);
    }
#ifdef OVS_B200_WITH_REFERENCE_TYPES
    //! The reference's signature (match/robust.h); body in adapters.hpp.
    unsigned int brute_force_match(data::frame& frm, data::keyframe* keyfrm, std::vector<std::pair<int, int>>& matches) const;
#endif
    //! match_for_triangulation(keyfrm_1, keyfrm_2, E_12, matched_idx_pairs): the keyframes' BoW feature vectors come in as
    //! per-keypoint node ids; matched pairs = (idx in keyframe 1, idx in keyframe 2)
    struct triangulation_view {
        int num_keypts; const std::uint8_t* descriptors; const double* bearings; const std::int32_t* octave; const float* angle;
        const std::uint8_t* has_landmark; const std::uint8_t* is_stereo; const std::int32_t* bow_node;
    };
    unsigned int match_for_triangulation(const triangulation_view& keyfrm_1, const triangulation_view& keyfrm_2, const double* E_12,
                                         const double* epipole_in_keyfrm_2, const std::vector<float>& scale_factors_1,
                                         std::vector<std::pair<unsigned int, unsigned int>>& matched_idx_pairs) const {
        std::vector<std::int32_t> m(static_cast<std::size_t>(std::max(1, keyfrm_1.num_keypts)), -1);
        int n = 0;
        detail::check(ovs_robust_match_for_triangulation_host(h_, keyfrm_1.num_keypts, keyfrm_1.descriptors, keyfrm_1.bearings, keyfrm_1.octave,
                                                              keyfrm_1.angle, keyfrm_1.has_landmark, keyfrm_1.is_stereo, keyfrm_1.bow_node,
                                                              keyfrm_2.num_keypts, keyfrm_2.descriptors, keyfrm_2.bearings, keyfrm_2.angle,
                                                              keyfrm_2.has_landmark, keyfrm_2.is_stereo, keyfrm_2.bow_node, E_12, epipole_in_keyfrm_2,
                                                              scale_factors_1.data(), static_cast<int>(scale_factors_1.size()), check_orientation_,
                                                              m.data(), &n));
        matched_idx_pairs.clear();
        for (int i = 0; i < keyfrm_1.num_keypts; ++i)
            if (m[i] >= 0) matched_idx_pairs.emplace_back(static_cast<unsigned int>(i), static_cast<unsigned int>(m[i]));
        return static_cast<unsigned int>(n);
    }
};

class projection final : public base {
public:
    explicit projection(const float lowe_ratio = 0.6, const bool check_orientation = true, const int device = 0)
        : base(lowe_ratio, check_orientation, device) {}
    //! match_frame_and_landmarks(frm, local_landmarks, margin)
    unsigned int match_frame_and_landmarks(const frame_index& frm, const std::vector<float>& scale_factors, const int num_landmarks,
                                           const std::uint8_t* lm_usable, const float* reproj_in_tracking, const float* x_right_in_tracking,
                                           const std::int32_t* scale_level_in_tracking, const std::uint8_t* lm_descriptors,
                                           const std::uint8_t* kp_has_observed_lm, std::vector<std::int32_t>& matched_lm_of_kp,
                                           const float margin = 5.0) const {
        matched_lm_of_kp.assign(std::max(1, frm.num_keypts()), -1);
        int n = 0;
        detail::check(ovs_projection_match_frame_and_landmarks_host(frm.handle(), scale_factors.data(), static_cast<int>(scale_factors.size()), num_landmarks, lm_usable, reproj_in_tracking,
                                                                    x_right_in_tracking, scale_level_in_tracking, lm_descriptors, kp_has_observed_lm,
                                                                    margin, lowe_ratio_, matched_lm_of_kp.data(), &n));
        matched_lm_of_kp.resize(frm.num_keypts());
        return static_cast<unsigned int>(n);
    }
#ifdef OVS_B200_WITH_REFERENCE_TYPES
    //! The reference's signature (match/projection.h); body in adapters.hpp.
    unsigned int match_frame_and_landmarks(data::frame& frm, const std::vector<data::landmark*>& local_landmarks, const float margin = 5.0) const;
#endif
    //! match_current_and_last_frames(curr_frm, last_frm, margin)
    unsigned int match_current_and_last_frames(const frame_index& curr, const std::vector<float>& scale_factors, const int num_last_keypts,
                                               const std::uint8_t* last_usable, const float* reproj, const float* reproj_x_right,
                                               const std::int32_t* last_scale_level, const float* last_angle, const std::uint8_t* lm_descriptors,
                                               const std::uint8_t* kp_has_observed_lm, std::vector<std::int32_t>& matched_last_of_kp,
                                               const float margin, const bool assume_forward, const bool assume_backward) const {
        matched_last_of_kp.assign(std::max(1, curr.num_keypts()), -1);
        int n = 0;
        detail::check(ovs_projection_match_current_and_last_host(curr.handle(), scale_factors.data(), static_cast<int>(scale_factors.size()),
                                                                 num_last_keypts, last_usable, reproj, reproj_x_right, last_scale_level, last_angle,
                                                                 lm_descriptors, kp_has_observed_lm, margin, assume_forward, assume_backward,
                                                                 check_orientation_, matched_last_of_kp.data(), &n));
        matched_last_of_kp.resize(curr.num_keypts());
        return static_cast<unsigned int>(n);
    }
    //! match_frame_and_keyframe(curr_frm, keyfrm, already_matched_lms, margin, hamm_dist_thr): the keyframe's landmarks
    //! reprojected into curr_frm by the caller (reproj, pred_scale_level, usable); window levels [pred - 1, pred + 1]
    unsigned int match_frame_and_keyframe(const frame_index& curr, const std::vector<float>& scale_factors, const int num_landmarks,
                                          const std::uint8_t* usable, const float* reproj, const std::int32_t* pred_scale_level,
                                          const float* keyfrm_angle, const std::uint8_t* lm_descriptors, const std::uint8_t* kp_has_lm,
                                          std::vector<std::int32_t>& matched_lm_of_kp, const float margin, const unsigned int hamm_dist_thr) const {
        return match_best_(curr, scale_factors, num_landmarks, usable, reproj, pred_scale_level, 1, keyfrm_angle, lm_descriptors, kp_has_lm,
                           matched_lm_of_kp, margin, hamm_dist_thr, check_orientation_);
    }
    //! match_by_Sim3_transform(keyfrm, Sim3_cw, landmarks, matched_lms_in_keyfrm, margin): window levels [pred - 1, pred],
    //! distance <= HAMMING_DIST_THR_LOW, no orientation check
    unsigned int match_by_Sim3_transform(const frame_index& keyfrm, const std::vector<float>& scale_factors, const int num_landmarks,
                                         const std::uint8_t* usable, const float* reproj, const std::int32_t* pred_scale_level,
                                         const std::uint8_t* lm_descriptors, const std::uint8_t* kp_already_matched,
                                         std::vector<std::int32_t>& matched_lm_of_kp, const float margin) const {
        const std::vector<float> zero(static_cast<std::size_t>(std::max(1, num_landmarks)), 0.0f);
        return match_best_(keyfrm, scale_factors, num_landmarks, usable, reproj, pred_scale_level, 0, zero.data(), lm_descriptors, kp_already_matched,
                           matched_lm_of_kp, margin, OVS_HAMMING_DIST_THR_LOW, false);
    }
    //! match_keyframes_mutually(keyfrm_1, keyfrm_2, matched_lms_in_keyfrm_2, Sim3s, margin): landmark arrays indexed by the
    //! keypoint of their keyframe, reprojected into the other keyframe by the caller
    unsigned int match_keyframes_mutually(const frame_index& keyfrm_1, const frame_index& keyfrm_2, const std::vector<float>& scale_factors,
                                          const std::uint8_t* usable_1, const float* reproj_1_in_2, const std::int32_t* pred_level_1_in_2,
                                          const std::uint8_t* lm_descriptors_1, const std::uint8_t* usable_2, const float* reproj_2_in_1,
                                          const std::int32_t* pred_level_2_in_1, const std::uint8_t* lm_descriptors_2,
                                          std::vector<std::int32_t>& matched_idx_2_of_kp_1, const float margin) const {
        matched_idx_2_of_kp_1.assign(std::max(1, keyfrm_1.num_keypts()), -1);
        int n = 0;
        detail::check(ovs_projection_match_keyframes_mutually_host(keyfrm_1.handle(), keyfrm_2.handle(), scale_factors.data(), usable_1, reproj_1_in_2,
                                                                   pred_level_1_in_2, lm_descriptors_1, usable_2, reproj_2_in_1, pred_level_2_in_1,
                                                                   lm_descriptors_2, margin, matched_idx_2_of_kp_1.data(), &n));
        matched_idx_2_of_kp_1.resize(keyfrm_1.num_keypts());
        return static_cast<unsigned int>(n);
    }

private:
    unsigned int match_best_(const frame_index& frm, const std::vector<float>& scale_factors, const int nq, const std::uint8_t* usable,
                             const float* reproj, const std::int32_t* pred_scale_level, const int levels_above, const float* q_angle,
                             const std::uint8_t* q_desc, const std::uint8_t* kp_unavailable, std::vector<std::int32_t>& matched_query_of_kp,
                             const float margin, const unsigned int hamm_dist_thr, const bool check_orientation) const {
        std::vector<float> mg(static_cast<std::size_t>(std::max(1, nq)));
        std::vector<std::int32_t> lo(mg.size()), hi(mg.size());
        for (int q = 0; q < nq; ++q) {
            const int l = pred_scale_level[q];
            mg[q] = margin * scale_factors[static_cast<std::size_t>(l < 0 ? 0 : l)];
            lo[q] = l - 1; hi[q] = l + levels_above;
        }
        matched_query_of_kp.assign(std::max(1, frm.num_keypts()), -1);
        int n = 0;
        detail::check(ovs_projection_match_best_host(frm.handle(), nq, usable, reproj, nullptr, mg.data(), lo.data(), hi.data(), q_angle, q_desc,
                                                     kp_unavailable, hamm_dist_thr, check_orientation, matched_query_of_kp.data(), &n));
        matched_query_of_kp.resize(frm.num_keypts());
        return static_cast<unsigned int>(n);
    }
};

class area final : public base {
public:
    explicit area(const float lowe_ratio = 0.9, const bool check_orientation = true, const int device = 0)
        : base(lowe_ratio, check_orientation, device) {}
    //! match_in_consistent_area(frm_1, frm_2, prev_matched_pts, matched_indices_2_in_frm_1, margin)
    unsigned int match_in_consistent_area(const frame_view& frm_1, const frame_index& frm_2, std::vector<float>& prev_matched_pts_xy,
                                          std::vector<int>& matched_indices_2_in_frm_1, const int margin = 20) const {
        matched_indices_2_in_frm_1.assign(std::max(1, frm_1.num_keypts), -1);
        int n = 0;
        detail::check(ovs_area_match_in_consistent_area_host(frm_2.handle(), frm_1.num_keypts, frm_1.octave, frm_1.angle, frm_1.descriptors,
                                                             prev_matched_pts_xy.data(), matched_indices_2_in_frm_1.data(), margin, lowe_ratio_,
                                                             check_orientation_, &n));
        matched_indices_2_in_frm_1.resize(frm_1.num_keypts);
        return static_cast<unsigned int>(n);
    }
};

class stereo final : public base {
public:
    //! The reference constructor takes the two image pyramids; here they stay on the device inside
    //! the two extractors that produced the keypoints.
    stereo(const feature::orb_extractor& extractor_left, const feature::orb_extractor& extractor_right,
           const float focal_x_baseline, const float true_baseline, const int device = 0)
        : base(0.0f, false, device), left_(extractor_left.handle()), right_(extractor_right.handle()),
          focal_x_baseline_(focal_x_baseline), true_baseline_(true_baseline) {}
    //! compute(stereo_x_right, depths)
    void compute(const frame_view& left, const frame_view& right, std::vector<float>& stereo_x_right, std::vector<float>& depths) const {
        stereo_x_right.assign(std::max(1, left.num_keypts), -1.0f); depths.assign(std::max(1, left.num_keypts), -1.0f);
        detail::check(ovs_stereo_compute_host(h_, left_, right_, left.num_keypts, left.x, left.y, left.octave, left.descriptors, right.num_keypts,
                                              right.x, right.y, right.octave, right.descriptors, focal_x_baseline_, true_baseline_,
                                              stereo_x_right.data(), depths.data(), nullptr));
        stereo_x_right.resize(left.num_keypts); depths.resize(left.num_keypts);
    }
private:
    const ovs_extractor* left_; const ovs_extractor* right_;
    const float focal_x_baseline_, true_baseline_;
};

}  // namespace match

namespace optimize {

class pose_optimizer {
public:
    explicit pose_optimizer(const unsigned int num_trials = 4, const unsigned int num_each_iter = 10, const int device = 0)
        : num_trials_(num_trials), num_each_iter_(num_each_iter) { detail::check(ovs_optimizer_create(device, &h_)); }
    ~pose_optimizer() { ovs_optimizer_destroy(h_); }
    pose_optimizer(const pose_optimizer&) = delete;
    pose_optimizer& operator=(const pose_optimizer&) = delete;
    //! optimize(frm): pose_cw (R row-major, t) is updated in place, outlier_flags filled; returns the inlier count.
    unsigned int optimize(const ovs_camera& camera, const bool setup_is_monocular, const int num_obs, const double* pos_w, const float* undist_xy,
                          const float* stereo_x_right, const float* inv_level_sigma_sq, double* cam_pose_cw, std::vector<std::uint8_t>& outlier_flags) const {
        outlier_flags.assign(std::max(1, num_obs), 0);
        int n = 0;
        detail::check(ovs_pose_optimize_host(h_, &camera, setup_is_monocular, num_obs, pos_w, undist_xy, stereo_x_right, inv_level_sigma_sq, cam_pose_cw,
                                             outlier_flags.data(), static_cast<int>(num_trials_), static_cast<int>(num_each_iter_), &n, nullptr));
        outlier_flags.resize(num_obs);
        return static_cast<unsigned int>(n);
    }
#ifdef OVS_B200_WITH_REFERENCE_TYPES
    //! The reference's signature (optimize/pose_optimizer.h); body in adapters.hpp.
    unsigned int optimize(data::frame& frm) const;
#endif
private:
    const unsigned int num_trials_, num_each_iter_;
    ovs_optimizer* h_ = nullptr;
};

class local_bundle_adjuster {
public:
    explicit local_bundle_adjuster(const unsigned int num_first_iter = 5, const unsigned int num_second_iter = 10, const int device = 0)
        : num_first_iter_(num_first_iter), num_second_iter_(num_second_iter) { detail::check(ovs_optimizer_create(device, &h_)); }
    ~local_bundle_adjuster() { ovs_optimizer_destroy(h_); }
    local_bundle_adjuster(const local_bundle_adjuster&) = delete;
    local_bundle_adjuster& operator=(const local_bundle_adjuster&) = delete;
    //! optimize(curr_keyfrm, force_stop_flag) on the flattened local graph (see include/ovs_b200.h).
    void optimize(const ovs_camera& camera, const bool setup_is_monocular, const int num_keyfrms, double* cam_poses_cw, const std::uint8_t* is_fixed,
                  const int num_landmarks, double* pos_w, const int num_obs, const std::int32_t* obs_keyfrm, const std::int32_t* obs_landmark,
                  const float* undist_xy, const float* stereo_x_right, const float* inv_level_sigma_sq, bool* const force_stop_flag,
                  std::vector<std::uint8_t>& outlier_observations) const {
        outlier_observations.assign(std::max(1, num_obs), 0);
        static_assert(sizeof(bool) == 1, "the C ABI polls the flag as one byte");
        detail::check(ovs_local_ba_host(h_, &camera, setup_is_monocular, num_keyfrms, cam_poses_cw, is_fixed, num_landmarks, pos_w, num_obs, obs_keyfrm,
                                        obs_landmark, undist_xy, stereo_x_right, inv_level_sigma_sq, static_cast<int>(num_first_iter_),
                                        static_cast<int>(num_second_iter_), reinterpret_cast<const volatile std::uint8_t*>(force_stop_flag), outlier_observations.data(), nullptr));
        outlier_observations.resize(num_obs);
    }
#ifdef OVS_B200_WITH_REFERENCE_TYPES
    //! The reference's signature (optimize/local_bundle_adjuster.h); body in adapters.hpp.
    void optimize(data::keyframe* curr_keyfrm, bool* const force_stop_flag) const;
#endif
private:
    const unsigned int num_first_iter_, num_second_iter_;
    ovs_optimizer* h_ = nullptr;
};

}  // namespace optimize
}  // namespace openvslam
