// adapters.hpp -- the reference's own method signatures on the reference's own data model, implemented over the class layer
// of openvslam_b200.hpp (which works on array views).  This is the binding a maintainer adds to the reference tree:
//
//   feature::orb_extractor::extract(const cv::_InputArray&, const cv::_InputArray&, std::vector<cv::KeyPoint>&, const cv::_OutputArray&)
//   match::robust::brute_force_match(data::frame&, data::keyframe*, std::vector<std::pair<int, int>>&)           (match/robust.h)
//   match::projection::match_frame_and_landmarks(data::frame&, const std::vector<data::landmark*>&, float)      (match/projection.h)
//   optimize::pose_optimizer::optimize(data::frame&)                                                             (optimize/pose_optimizer.h)
//   optimize::local_bundle_adjuster::optimize(data::keyframe*, bool* const)                                      (optimize/local_bundle_adjuster.h)
//
// Include it INSTEAD of openvslam_b200.hpp in a translation unit that can see the reference's headers (here: the stand-ins
// under tests/cpp/standin, which declare the members used below with the names recalled in SURVEY.md section 2 / 8b;
// /root/reference holds no source, so no file:line can be cited).  tests/test_class_layer.py compiles this file with g++ and
// runs one call of each method on the GPU (tests/cpp/test_adapters.cpp).
#pragma once

#ifndef OVS_B200_WITH_REFERENCE_TYPES
#define OVS_B200_WITH_REFERENCE_TYPES
#endif
#ifndef OVS_B200_WITH_OPENCV
#define OVS_B200_WITH_OPENCV
#endif

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include <opencv2/core.hpp>
#include "openvslam/camera/base.h"
#include "openvslam/data/frame.h"
#include "openvslam/data/keyframe.h"
#include "openvslam/data/landmark.h"
#include "openvslam/type.h"

#include "openvslam_b200.hpp"

namespace openvslam {
namespace adapters {

//! camera::base -> the parameters the reprojection edges read.  Fisheye / radial-division cameras optimise on undistorted
//! keypoints with the perspective edges, as in the reference.
inline ovs_camera to_camera(const camera::base* cam) {
    ovs_camera c{};
    c.focal_x_baseline = cam->focal_x_baseline_; c.cols = cam->cols_; c.rows = cam->rows_;
    switch (cam->model_type_) {
        case camera::model_type_t::Perspective: {
            auto p = static_cast<const camera::perspective*>(cam);
            c.model = OVS_CAMERA_PERSPECTIVE; c.fx = p->fx_; c.fy = p->fy_; c.cx = p->cx_; c.cy = p->cy_; break;
        }
        case camera::model_type_t::Fisheye: {
            auto p = static_cast<const camera::fisheye*>(cam);
            c.model = OVS_CAMERA_PERSPECTIVE; c.fx = p->fx_; c.fy = p->fy_; c.cx = p->cx_; c.cy = p->cy_; break;
        }
        case camera::model_type_t::RadialDivision: {
            auto p = static_cast<const camera::radial_division*>(cam);
            c.model = OVS_CAMERA_PERSPECTIVE; c.fx = p->fx_; c.fy = p->fy_; c.cx = p->cx_; c.cy = p->cy_; break;
        }
        case camera::model_type_t::Equirectangular: c.model = OVS_CAMERA_EQUIRECTANGULAR; break;
    }
    return c;
}

inline void to_Rt(const Mat44_t& T, double* pose12) {
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) pose12[3 * r + c] = T(r, c); pose12[9 + r] = T(r, 3); }
}
inline Mat44_t from_Rt(const double* pose12) {
    Mat44_t T = Mat44_t::Identity();
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T(r, c) = pose12[3 * r + c]; T(r, 3) = pose12[9 + r]; }
    return T;
}

//! undist_keypts_ / stereo_x_right_ / descriptors_ + the camera's grid constants as a match::frame_view (owns the SoA copies)
struct frame_arrays {
    std::vector<float> x, y, angle, x_right;
    std::vector<std::int32_t> octave;
    std::vector<std::uint8_t> desc;
    match::frame_view view;
    template <class FrameLike>
    explicit frame_arrays(const FrameLike& f) {
        const std::size_t n = f.undist_keypts_.size();
        x.resize(n); y.resize(n); angle.resize(n); octave.resize(n);
        for (std::size_t i = 0; i < n; ++i) {
            const cv::KeyPoint& k = f.undist_keypts_[i];
            x[i] = k.pt.x; y[i] = k.pt.y; angle[i] = k.angle; octave[i] = k.octave;
        }
        x_right.assign(f.stereo_x_right_.begin(), f.stereo_x_right_.end());
        desc.resize(n * 32);
        for (std::size_t i = 0; i < n; ++i) std::memcpy(&desc[32 * i], f.descriptors_.ptr(static_cast<int>(i)), 32);   // rows may be strided
        view.num_keypts = static_cast<int>(n);
        view.x = x.data(); view.y = y.data(); view.octave = octave.data(); view.angle = angle.data();
        view.stereo_x_right = (x_right.size() == n && n > 0) ? x_right.data() : nullptr;
        view.descriptors = desc.data();
        const camera::base* cam = f.camera_;
        view.grid.min_x = cam->img_bounds_.min_x_; view.grid.min_y = cam->img_bounds_.min_y_;
        view.grid.inv_cell_width = cam->inv_cell_width_; view.grid.inv_cell_height = cam->inv_cell_height_;
        view.grid.num_grid_cols = static_cast<std::int32_t>(cam->num_grid_cols_); view.grid.num_grid_rows = static_cast<std::int32_t>(cam->num_grid_rows_);
    }
};

}  // namespace adapters

// ---------------------------------------------------------------------------------------------------- match::robust
inline unsigned int match::robust::brute_force_match(data::frame& frm, data::keyframe* keyfrm, std::vector<std::pair<int, int>>& matches) const {
    const auto lms_2 = keyfrm->get_landmarks();
    const int n1 = static_cast<int>(frm.num_keypts_), n2 = static_cast<int>(keyfrm->num_keypts_);
    std::vector<std::uint8_t> valid(static_cast<std::size_t>(std::max(n2, 1)), 0);
    for (int i = 0; i < n2 && i < static_cast<int>(lms_2.size()); ++i) valid[i] = lms_2[i] && !lms_2[i]->will_be_erased();
    std::vector<std::uint8_t> d1(static_cast<std::size_t>(n1) * 32), d2(static_cast<std::size_t>(n2) * 32);
    for (int i = 0; i < n1; ++i) std::memcpy(&d1[32 * static_cast<std::size_t>(i)], frm.descriptors_.ptr(i), 32);
    for (int i = 0; i < n2; ++i) std::memcpy(&d2[32 * static_cast<std::size_t>(i)], keyfrm->descriptors_.ptr(i), 32);
    return brute_force_match(d1.data(), n1, d2.data(), n2, valid.data(), matches);
}

// ------------------------------------------------------------------------------------------------ match::projection
inline unsigned int match::projection::match_frame_and_landmarks(data::frame& frm, const std::vector<data::landmark*>& local_landmarks,
                                                                 const float margin) const {
    const adapters::frame_arrays arrays(frm);
    const frame_index idx(*this, arrays.view);               // upload + cell index (reusable by every matcher call on this frame)
    const std::size_t L = local_landmarks.size();
    const unsigned int n = frm.num_keypts_;
    std::vector<std::uint8_t> usable(std::max<std::size_t>(L, 1), 0), has(std::max<unsigned int>(n, 1), 0), desc(32 * std::max<std::size_t>(L, 1));
    std::vector<float> reproj(2 * std::max<std::size_t>(L, 1)), xr(std::max<std::size_t>(L, 1));
    std::vector<std::int32_t> lvl(std::max<std::size_t>(L, 1), 0);
    for (std::size_t l = 0; l < L; ++l) {
        const data::landmark* lm = local_landmarks[l];
        usable[l] = lm && lm->is_observable_in_tracking_ && !lm->will_be_erased();
        if (!usable[l]) continue;
        reproj[2 * l] = static_cast<float>(lm->reproj_in_tracking_(0)); reproj[2 * l + 1] = static_cast<float>(lm->reproj_in_tracking_(1));
        xr[l] = lm->x_right_in_tracking_; lvl[l] = lm->scale_level_in_tracking_;
        const cv::Mat d = lm->get_descriptor();
        std::memcpy(&desc[32 * l], d.data, 32);
    }
    for (unsigned int i = 0; i < n; ++i) has[i] = frm.landmarks_.at(i) && frm.landmarks_.at(i)->has_observation();
    std::vector<std::int32_t> matched;
    const unsigned int num = match_frame_and_landmarks(idx, frm.scale_factors_, static_cast<int>(L), usable.data(), reproj.data(), xr.data(), lvl.data(),
                                                       desc.data(), has.data(), matched, margin);
    for (unsigned int i = 0; i < n; ++i) if (matched[i] >= 0) frm.landmarks_.at(i) = local_landmarks[static_cast<std::size_t>(matched[i])];
    return num;
}

// --------------------------------------------------------------------------------------------- optimize::pose_optimizer
inline unsigned int optimize::pose_optimizer::optimize(data::frame& frm) const {
    // one edge per keypoint with a valid landmark (the reference's loop over frm.landmarks_)
    std::vector<unsigned int> idxs;
    std::vector<double> pos_w; std::vector<float> xy, x_right, inv_sigma_sq;
    const unsigned int num_keypts = frm.num_keypts_;
    for (unsigned int idx = 0; idx < num_keypts; ++idx) {
        data::landmark* lm = frm.landmarks_.at(idx);
        if (!lm) continue;
        if (lm->will_be_erased()) continue;
        frm.outlier_flags_.at(idx) = false;
        const cv::KeyPoint& undist_keypt = frm.undist_keypts_.at(idx);
        const Vec3_t p = lm->get_pos_in_world();
        idxs.push_back(idx);
        pos_w.push_back(p(0)); pos_w.push_back(p(1)); pos_w.push_back(p(2));
        xy.push_back(undist_keypt.pt.x); xy.push_back(undist_keypt.pt.y);
        x_right.push_back(idx < frm.stereo_x_right_.size() ? frm.stereo_x_right_.at(idx) : -1.0f);
        inv_sigma_sq.push_back(frm.inv_level_sigma_sq_.at(static_cast<std::size_t>(undist_keypt.octave)));
    }
    const int num_init_obs = static_cast<int>(idxs.size());
    if (num_init_obs < 5) return 0;
    double pose[12];
    adapters::to_Rt(frm.cam_pose_cw_, pose);
    const ovs_camera cam = adapters::to_camera(frm.camera_);
    std::vector<std::uint8_t> outlier;
    const unsigned int num_inliers = optimize(cam, frm.camera_->setup_type_ == camera::setup_type_t::Monocular, num_init_obs, pos_w.data(), xy.data(),
                                              x_right.data(), inv_sigma_sq.data(), pose, outlier);
    for (int k = 0; k < num_init_obs; ++k) frm.outlier_flags_.at(idxs[static_cast<std::size_t>(k)]) = outlier[static_cast<std::size_t>(k)] != 0;
    frm.set_cam_pose(adapters::from_Rt(pose));
    return num_inliers;
}

// ------------------------------------------------------------------------------------- optimize::local_bundle_adjuster
inline void optimize::local_bundle_adjuster::optimize(data::keyframe* curr_keyfrm, bool* const force_stop_flag) const {
    // 1. local keyframes (the current one and its covisibilities), local landmarks (seen by them), fixed keyframes (other observers
    //    of the local landmarks).  The reference collects them in unordered_maps keyed by id; ordered maps here, so that vertex and
    //    edge order -- and with it the rounding of the sums -- is reproducible.
    std::map<unsigned int, data::keyframe*> local_keyfrms, fixed_keyfrms;
    std::map<unsigned int, data::landmark*> local_lms;
    local_keyfrms[curr_keyfrm->id_] = curr_keyfrm;
    for (data::keyframe* k : curr_keyfrm->graph_node_->get_covisibilities()) {
        if (!k || k->will_be_erased()) continue;
        local_keyfrms[k->id_] = k;
    }
    for (const auto& id_kf : local_keyfrms)
        for (data::landmark* lm : id_kf.second->get_landmarks()) {
            if (!lm || lm->will_be_erased()) continue;
            local_lms[lm->id_] = lm;
        }
    for (const auto& id_lm : local_lms)
        for (const auto& obs : id_lm.second->get_observations()) {
            data::keyframe* k = obs.first;
            if (!k || k->will_be_erased()) continue;
            if (local_keyfrms.count(k->id_)) continue;
            fixed_keyfrms[k->id_] = k;
        }
    if (local_lms.empty()) return;
    // 2. the graph as arrays: keyframe vertices (local ones free, keyframe id 0 and the fixed ones fixed), landmark vertices,
    //    one edge per observation, landmark by landmark
    std::vector<data::keyframe*> kfs; std::map<data::keyframe*, int> kf_index;
    std::vector<std::uint8_t> is_fixed;
    for (const auto& id_kf : local_keyfrms) { kf_index[id_kf.second] = static_cast<int>(kfs.size()); kfs.push_back(id_kf.second); is_fixed.push_back(id_kf.first == 0); }
    for (const auto& id_kf : fixed_keyfrms) { kf_index[id_kf.second] = static_cast<int>(kfs.size()); kfs.push_back(id_kf.second); is_fixed.push_back(1); }
    const int K = static_cast<int>(kfs.size());
    std::vector<double> poses(static_cast<std::size_t>(K) * 12);
    for (int k = 0; k < K; ++k) adapters::to_Rt(kfs[static_cast<std::size_t>(k)]->get_cam_pose(), &poses[static_cast<std::size_t>(k) * 12]);
    std::vector<data::landmark*> lms;
    std::vector<double> points;
    std::vector<std::int32_t> obs_kf, obs_lm; std::vector<float> obs_xy, obs_xr, obs_w;
    std::vector<std::pair<data::keyframe*, data::landmark*>> obs_pairs;
    for (const auto& id_lm : local_lms) {
        data::landmark* lm = id_lm.second;
        const int l = static_cast<int>(lms.size());
        bool any = false;
        // observers in keyframe-id order (std::map<keyframe*, unsigned> iterates by address in the reference)
        std::map<unsigned int, std::pair<data::keyframe*, unsigned int>> by_id;
        for (const auto& obs : lm->get_observations()) if (obs.first && !obs.first->will_be_erased() && kf_index.count(obs.first)) by_id[obs.first->id_] = {obs.first, obs.second};
        for (const auto& e : by_id) {
            data::keyframe* k = e.second.first; const unsigned int idx = e.second.second;
            const cv::KeyPoint& undist_keypt = k->undist_keypts_.at(idx);
            obs_kf.push_back(kf_index[k]); obs_lm.push_back(l);
            obs_xy.push_back(undist_keypt.pt.x); obs_xy.push_back(undist_keypt.pt.y);
            obs_xr.push_back(idx < k->stereo_x_right_.size() ? k->stereo_x_right_.at(idx) : -1.0f);
            obs_w.push_back(k->inv_level_sigma_sq_.at(static_cast<std::size_t>(undist_keypt.octave)));
            obs_pairs.emplace_back(k, lm);
            any = true;
        }
        if (!any) continue;
        lms.push_back(lm);
        const Vec3_t p = lm->get_pos_in_world();
        points.push_back(p(0)); points.push_back(p(1)); points.push_back(p(2));
    }
    const int L = static_cast<int>(lms.size()), M = static_cast<int>(obs_kf.size());
    if (L == 0 || M == 0) return;
    // 3.-7. the two Levenberg rounds with the outlier cut in between (replaces the g2o call)
    const ovs_camera cam = adapters::to_camera(curr_keyfrm->camera_);
    std::vector<std::uint8_t> outlier;
    optimize(cam, curr_keyfrm->camera_->setup_type_ == camera::setup_type_t::Monocular, K, poses.data(), is_fixed.data(), L, points.data(), M,
             obs_kf.data(), obs_lm.data(), obs_xy.data(), obs_xr.data(), obs_w.data(), force_stop_flag, outlier);
    // 8. under the map lock: erase the outlier observations, write the estimates back
    {
        std::lock_guard<std::mutex> lock(data::map_database::mtx_database_);
        for (int i = 0; i < M; ++i) {
            if (!outlier[static_cast<std::size_t>(i)]) continue;
            data::keyframe* k = obs_pairs[static_cast<std::size_t>(i)].first; data::landmark* lm = obs_pairs[static_cast<std::size_t>(i)].second;
            k->erase_landmark(lm);
            lm->erase_observation(k);
        }
        for (int k = 0; k < K; ++k)
            if (!is_fixed[static_cast<std::size_t>(k)] && local_keyfrms.count(kfs[static_cast<std::size_t>(k)]->id_))
                kfs[static_cast<std::size_t>(k)]->set_cam_pose(adapters::from_Rt(&poses[static_cast<std::size_t>(k) * 12]));
        for (int l = 0; l < L; ++l) {
            Vec3_t p; p(0) = points[3 * static_cast<std::size_t>(l)]; p(1) = points[3 * static_cast<std::size_t>(l) + 1]; p(2) = points[3 * static_cast<std::size_t>(l) + 2];
            lms[static_cast<std::size_t>(l)]->set_pos_in_world(p);
            lms[static_cast<std::size_t>(l)]->update_normal_and_depth();
        }
    }
}

}  // namespace openvslam
