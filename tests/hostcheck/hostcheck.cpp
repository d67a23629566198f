// tests/hostcheck/hostcheck.cpp -- TEST SHIM (not a product path, not a CPU fallback).
// Compiles the __host__ __device__ arithmetic of csrc/orb_math.cuh / ba_math.cuh and the list-based host tree
// distribution (keypoint_tree.cpp: the product's code until the tree moved to the GPU, kept as a second
// independent restatement) with g++ so the CPU test-suite can compare them with the oracle without a GPU.  The kernels' indexing/tiling is checked on the GPU (-m gpu tests).
#include <stdint.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "../../openvslam_b200/csrc/orb_math.cuh"
#include "keypoint_tree.h"

extern "C" {

int hc_distribute(const uint32_t* cand, int n, int min_x, int max_x, int min_y, int max_y, unsigned num_keypts, int* out) {
    ovs::TreeScratch s;
    std::vector<int> o((size_t)n + 8);
    int m = ovs::distribute_keypoints_via_tree(cand, n, min_x, max_x, min_y, max_y, num_keypts, o.data(), s);
    memcpy(out, o.data(), sizeof(int) * m);
    return m;
}

// timing aid: `reps` runs with one persistent scratch (as the extractor's level workers keep theirs); returns ns per run
double hc_distribute_time_ns(const uint32_t* cand, int n, int min_x, int max_x, int min_y, int max_y, unsigned num_keypts, int reps) {
    ovs::TreeScratch s;
    std::vector<int> o((size_t)n + 8);
    ovs::distribute_keypoints_via_tree(cand, n, min_x, max_x, min_y, max_y, num_keypts, o.data(), s);
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) ovs::distribute_keypoints_via_tree(cand, n, min_x, max_x, min_y, max_y, num_keypts, o.data(), s);
    return std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / reps;
}

void hc_fast_score_map(const uint8_t* img, int w, int h, int stride, int min_thr, uint8_t* score) {
    const int dx[16] = OVS_FAST_RING_DX, dy[16] = OVS_FAST_RING_DY;
    memset(score, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = img + (size_t)y * stride + x;
            int d[16];
            for (int k = 0; k < 16; ++k) d[k] = (int)p[0] - (int)p[dy[k] * stride + dx[k]];
            if (!ovs::fast9_maybe(d[0], d[4], d[8], d[12], min_thr)) continue;
            const int s = ovs::fast9_score(d);
            score[(size_t)y * w + x] = (uint8_t)(s >= min_thr ? s : 0);
        }
}

float hc_fast_atan2(float y, float x) { return ovs::fast_atan2_deg(y, x); }

void hc_descriptor_offsets(float angle_deg, const int8_t* pattern /*[512][2]*/, int* drow, int* dcol) {
    float s, c;
    ovs::angle_sincos(angle_deg, &s, &c);
    for (int i = 0; i < 512; ++i) ovs::brief_offset(pattern[2 * i], pattern[2 * i + 1], s, c, &drow[i], &dcol[i]);
}

uint8_t hc_resize_px(int s00, int s01, int s10, int s11, int a0, int a1, int b0, int b1) {
    return ovs::resize_px(s00, s01, s10, s11, a0, a1, b0, b1);
}
}

// ---- FP64 BA arithmetic (csrc/ba_math.cuh) -------------------------------------------------
#include "../../openvslam_b200/csrc/ba_math.cuh"
extern "C" {
int hc_edge_eval(int model, const double* camp /* fx fy cx cy fb cols rows */, const double* pose, const double* pw,
                 const double* obs, int stereo, double* e, double* Jp, double* Jl) {
    ovs::CameraD c;
    c.model = model; c.fx = camp[0]; c.fy = camp[1]; c.cx = camp[2]; c.cy = camp[3]; c.fb = camp[4]; c.cols = camp[5]; c.rows = camp[6];
    return ovs::edge_eval(c, pose, pw, obs, stereo != 0, e, Jp, Jl);
}
void hc_pose_oplus(const double* pose, const double* u, double* out) { ovs::pose_oplus(pose, u, out); }
int hc_inv3_sym(const double* D, double* Di) { return ovs::inv3_sym(D, Di) ? 1 : 0; }
}
