// Stand-in for openvslam/camera/{base,perspective,fisheye,equirectangular,radial_division}.h.  See ../../README.md.
#pragma once
namespace openvslam { namespace camera {
enum class setup_type_t { Monocular = 0, Stereo = 1, RGBD = 2 };
enum class model_type_t { Perspective = 0, Fisheye = 1, Equirectangular = 2, RadialDivision = 3 };
struct image_bounds { float min_x_ = 0, max_x_ = 0, min_y_ = 0, max_y_ = 0; };
class base {
public:
    base(setup_type_t s, model_type_t m, unsigned cols, unsigned rows, double fxb) : setup_type_(s), model_type_(m), cols_(cols), rows_(rows), focal_x_baseline_(fxb) {
        img_bounds_.min_x_ = 0; img_bounds_.max_x_ = (float)cols; img_bounds_.min_y_ = 0; img_bounds_.max_y_ = (float)rows;
        inv_cell_width_ = (float)num_grid_cols_ / (img_bounds_.max_x_ - img_bounds_.min_x_);
        inv_cell_height_ = (float)num_grid_rows_ / (img_bounds_.max_y_ - img_bounds_.min_y_);
    }
    virtual ~base() = default;
    const setup_type_t setup_type_;
    const model_type_t model_type_;
    const unsigned cols_, rows_;
    const double focal_x_baseline_;
    const unsigned num_grid_cols_ = 64, num_grid_rows_ = 48;
    image_bounds img_bounds_;
    float inv_cell_width_ = 0, inv_cell_height_ = 0;
};
class perspective final : public base {
public:
    perspective(setup_type_t s, unsigned cols, unsigned rows, double fx, double fy, double cx, double cy, double fxb)
        : base(s, model_type_t::Perspective, cols, rows, fxb), fx_(fx), fy_(fy), cx_(cx), cy_(cy) {}
    const double fx_, fy_, cx_, cy_;
};
class fisheye final : public base {
public:
    fisheye(setup_type_t s, unsigned cols, unsigned rows, double fx, double fy, double cx, double cy, double fxb)
        : base(s, model_type_t::Fisheye, cols, rows, fxb), fx_(fx), fy_(fy), cx_(cx), cy_(cy) {}
    const double fx_, fy_, cx_, cy_;
};
class radial_division final : public base {
public:
    radial_division(setup_type_t s, unsigned cols, unsigned rows, double fx, double fy, double cx, double cy, double fxb)
        : base(s, model_type_t::RadialDivision, cols, rows, fxb), fx_(fx), fy_(fy), cx_(cx), cy_(cy) {}
    const double fx_, fy_, cx_, cy_;
};
class equirectangular final : public base {
public:
    equirectangular(unsigned cols, unsigned rows) : base(setup_type_t::Monocular, model_type_t::Equirectangular, cols, rows, 0.0) {}
};
}}  // namespace openvslam::camera
