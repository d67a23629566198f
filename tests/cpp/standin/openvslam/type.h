// Stand-in for openvslam/type.h (Eigen typedefs in the reference).  See ../README.md.
#pragma once
namespace openvslam {
struct Vec2_t { double v[2] = {0, 0}; double& operator()(int i) { return v[i]; } double operator()(int i) const { return v[i]; } };
struct Vec3_t { double v[3] = {0, 0, 0}; double& operator()(int i) { return v[i]; } double operator()(int i) const { return v[i]; } };
struct Mat33_t { double m[3][3] = {}; double& operator()(int r, int c) { return m[r][c]; } double operator()(int r, int c) const { return m[r][c]; } };
struct Mat44_t {
    double m[4][4] = {};
    double& operator()(int r, int c) { return m[r][c]; }
    double operator()(int r, int c) const { return m[r][c]; }
    static Mat44_t Identity() { Mat44_t I; for (int i = 0; i < 4; ++i) I.m[i][i] = 1.0; return I; }
};
}  // namespace openvslam
