// Stand-in for openvslam/data/{keyframe,graph_node,map_database}.h.  See ../../README.md.
#pragma once
#include <memory>
#include <mutex>
#include <vector>
#include <opencv2/core.hpp>
#include "openvslam/camera/base.h"
#include "openvslam/data/landmark.h"
#include "openvslam/type.h"
namespace openvslam { namespace data {
class keyframe;
class graph_node {
public:
    std::vector<keyframe*> get_covisibilities() const { return covisibilities_; }
    std::vector<keyframe*> covisibilities_;
};
class map_database { public: static std::mutex mtx_database_; };
inline std::mutex map_database::mtx_database_;
class keyframe {
public:
    keyframe(unsigned id, camera::base* cam) : id_(id), camera_(cam), graph_node_(new graph_node()) {}
    unsigned int id_;
    camera::base* camera_;
    unsigned int num_keypts_ = 0;
    std::vector<cv::KeyPoint> undist_keypts_;
    std::vector<float> stereo_x_right_;
    cv::Mat descriptors_;
    std::vector<float> scale_factors_, inv_level_sigma_sq_;
    const std::unique_ptr<graph_node> graph_node_;
    Mat44_t get_cam_pose() const { return cam_pose_cw_; }
    void set_cam_pose(const Mat44_t& p) { cam_pose_cw_ = p; }
    std::vector<landmark*> get_landmarks() const { return landmarks_; }
    landmark* get_landmark(unsigned idx) const { return landmarks_.at(idx); }
    void add_landmark(landmark* lm, unsigned idx) { if (landmarks_.size() <= idx) landmarks_.resize(idx + 1, nullptr); landmarks_[idx] = lm; }
    void erase_landmark_with_index(unsigned idx) { landmarks_.at(idx) = nullptr; ++num_erased_; }
    void erase_landmark(landmark* lm) { const int idx = lm->get_index_in_keyframe(this); if (0 <= idx) erase_landmark_with_index((unsigned)idx); }
    bool will_be_erased() const { return will_be_erased_; }
    bool will_be_erased_ = false;
    int num_erased_ = 0;
private:
    Mat44_t cam_pose_cw_ = Mat44_t::Identity();
    std::vector<landmark*> landmarks_;
};
}}  // namespace openvslam::data
