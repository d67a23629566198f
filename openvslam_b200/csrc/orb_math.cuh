// orb_math.cuh -- per-pixel / per-keypoint arithmetic of the ORB front-end, shared by the
// CUDA kernels (orb_extractor.cu).  Everything here is integer or strictly-ordered IEEE
// float arithmetic so the results are bit-identical to the CPU path of the reference
// (feature/orb_extractor.cc and the OpenCV primitives it calls; names as in SURVEY.md 8a).
//
// The functions are __host__ __device__ so tests/hostcheck can run the same arithmetic on
// the CPU against the oracle without a GPU (kernel indexing is then checked on the GPU).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define OVS_HD __host__ __device__ __forceinline__
#else
#define OVS_HD static inline
#endif

namespace ovs {

// ---- strictly rounded float ops (no FMA contraction, round-to-nearest-even) -------------
OVS_HD float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
OVS_HD float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
OVS_HD float fsub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
OVS_HD float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
// cvRound(float): round half to even.
OVS_HD int cv_round(float v) {
#ifdef __CUDA_ARCH__
    return __float2int_rn(v);
#else
    return (int)lrintf(v);
#endif
}

// ---- FAST-9/16 ------------------------------------------------------------------------
// Ring offsets in cv::FAST order: pixel[0] = (0,+3), clockwise in image coordinates.
#define OVS_FAST_RING_DX {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1}
#define OVS_FAST_RING_DY {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3}

OVS_HD int imin(int a, int b) { return a < b ? a : b; }
OVS_HD int imax(int a, int b) { return a > b ? a : b; }

// Corner score of cv::cornerScore<16>: S = max over the 16 arcs of 9 contiguous ring pixels of
// min(v - p) (brighter centre) or min(p - v) (darker centre), minus 1; i.e. the largest
// threshold t for which the pixel passes the FAST-9 segment test.  Returns 0 if S < 1.
// d[k] = v - ring[k], k = 0..15.
OVS_HD int fast9_score(const int* d) {
    // sliding minimum / maximum over windows of 9 on a circular array of 16, by doubling
    int mn2[16], mx2[16], mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn2[k] = imin(d[k], d[(k + 1) & 15]); mx2[k] = imax(d[k], d[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn4[k] = imin(mn2[k], mn2[(k + 2) & 15]); mx4[k] = imax(mx2[k], mx2[(k + 2) & 15]); }
    int best_bright = -256, best_dark = 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int mn9 = imin(imin(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int mx9 = imax(imax(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
        best_bright = imax(best_bright, mn9);  // max over arcs of min(d)
        best_dark = imin(best_dark, mx9);      // min over arcs of max(d)  (= -max min(-d))
    }
    const int s = imax(best_bright, -best_dark) - 1;
    return s > 0 ? s : 0;
}

// Necessary condition for a FAST-9 corner at threshold t: every arc of 9 contains at least
// one pixel of each opposite pair, so for a bright (dark) arc every pair has a bright (dark) one.
OVS_HD bool fast9_maybe(int d0, int d4, int d8, int d12, int t) {
    const bool bright = (d0 > t || d8 > t) && (d4 > t || d12 > t);
    const bool dark = (d0 < -t || d8 < -t) && (d4 < -t || d12 < -t);
    return bright || dark;
}

// ---- cv::resize INTER_LINEAR, CV_8UC1: one output pixel -----------------------------------
// s00,s01 = top row at xofs, xofs+1; s10,s11 = bottom row.  a0,a1 / b0,b1: 11-bit weights.
OVS_HD uint8_t resize_px(int s00, int s01, int s10, int s11, int a0, int a1, int b0, int b1) {
    const int r0 = s00 * a0 + s01 * a1;
    const int r1 = s10 * a0 + s11 * a1;
    int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    return (uint8_t)v;
}

// ---- cv::fastAtan2 (degrees), scalar path, strict float ------------------------------------
OVS_HD float fast_atan2_deg(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    const float eps = (float)2.2204460492503131e-16;  // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = fdiv(ay, fadd(ax, eps));
        const float c2 = fmul(c, c);
        a = fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = fdiv(ax, fadd(ay, eps));
        const float c2 = fmul(c, c);
        a = fsub(90.f, fmul(fadd(fmul(fadd(fmul(fadd(fmul(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = fsub(180.f, a);
    if (y < 0) a = fsub(360.f, a);
    return a;
}

// ---- 7x7 sigma-2 Gaussian, OpenCV's bit-exact 8.8 fixed-point kernel -------------------------
#define OVS_GAUSS7 {18, 34, 48, 56, 48, 34, 18}

// BORDER_REFLECT_101 index.
OVS_HD int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// ---- steered BRIEF: rotated sample offset -------------------------------------------------
// Row/column offsets of pattern point (px, py) for an angle with the given sin/cos, exactly as
// compute_orb_descriptor does: cvRound(px*sin + py*cos), cvRound(px*cos - py*sin).
OVS_HD void brief_offset(int px, int py, float sin_a, float cos_a, int* drow, int* dcol) {
    *drow = cv_round(fadd(fmul((float)px, sin_a), fmul((float)py, cos_a)));
    *dcol = cv_round(fsub(fmul((float)px, cos_a), fmul((float)py, sin_a)));
}

// keypt.angle (degrees, float) -> float radians as `keypt.angle * M_PI / 180.0` narrowed to
// float, then sin/cos of that float, rounded to float from a double evaluation.
OVS_HD void angle_sincos(float angle_deg, float* s, float* c) {
    const float a = (float)((double)angle_deg * 3.14159265358979323846 / 180.0);
    *s = (float)sin((double)a);
    *c = (float)cos((double)a);
}

}  // namespace ovs
