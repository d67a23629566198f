// Stand-in for openvslam/data/landmark.h.  See ../../README.md.
#pragma once
#include <map>
#include <opencv2/core.hpp>
#include "openvslam/type.h"
namespace openvslam { namespace data {
class keyframe;
class landmark {
public:
    landmark(unsigned id, const Vec3_t& pos_w) : id_(id), pos_w_(pos_w) {}
    unsigned id_;
    // tracking information (filled by frame::can_observe / the tracker before projection::match_frame_and_landmarks)
    Vec2_t reproj_in_tracking_;
    float x_right_in_tracking_ = -1.0f;
    bool is_observable_in_tracking_ = false;
    int scale_level_in_tracking_ = 0;
    Vec3_t get_pos_in_world() const { return pos_w_; }
    void set_pos_in_world(const Vec3_t& p) { pos_w_ = p; }
    cv::Mat get_descriptor() const { return descriptor_.clone(); }
    void set_descriptor(const cv::Mat& d) { descriptor_ = d.clone(); }
    std::map<keyframe*, unsigned int> get_observations() const { return observations_; }
    void add_observation(keyframe* k, unsigned idx) { observations_[k] = idx; }
    void erase_observation(keyframe* k) { observations_.erase(k); ++num_erased_; }
    bool has_observation() const { return !observations_.empty(); }
    int get_index_in_keyframe(keyframe* k) const { auto it = observations_.find(k); return it == observations_.end() ? -1 : (int)it->second; }
    bool will_be_erased() const { return will_be_erased_; }
    void update_normal_and_depth() { ++num_updates_; }
    int num_erased_ = 0, num_updates_ = 0;
    bool will_be_erased_ = false;
private:
    Vec3_t pos_w_;
    cv::Mat descriptor_;
    std::map<keyframe*, unsigned int> observations_;
};
}}  // namespace openvslam::data
