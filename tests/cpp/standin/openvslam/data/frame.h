// Stand-in for openvslam/data/frame.h.  See ../../README.md.
#pragma once
#include <vector>
#include <opencv2/core.hpp>
#include "openvslam/camera/base.h"
#include "openvslam/type.h"
namespace openvslam { namespace data {
class landmark;
class frame {
public:
    camera::base* camera_ = nullptr;
    unsigned int num_keypts_ = 0;
    std::vector<cv::KeyPoint> keypts_, undist_keypts_;
    std::vector<float> stereo_x_right_, depths_;
    cv::Mat descriptors_;
    std::vector<landmark*> landmarks_;
    std::vector<bool> outlier_flags_;
    std::vector<float> scale_factors_, inv_scale_factors_, level_sigma_sq_, inv_level_sigma_sq_;
    unsigned int num_scale_levels_ = 0;
    bool cam_pose_cw_is_valid_ = false;
    Mat44_t cam_pose_cw_ = Mat44_t::Identity();
    void set_cam_pose(const Mat44_t& cam_pose_cw) { cam_pose_cw_ = cam_pose_cw; cam_pose_cw_is_valid_ = true; }
};
}}  // namespace openvslam::data
