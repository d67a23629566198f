/* oracle/orb_oracle.h -- declarations for the CPU oracle (test infrastructure only;
 * see orb_oracle.c for the parity statement and the reference files restated). */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define OO_MAX_LEVELS 16

typedef struct { int x, y, score; } oo_fast_pt;

/* cv::KeyPoint fields the reference fills (class_id is always -1), plus the integer
 * level coordinates the keypoint was detected at (debug aid for tests). */
typedef struct {
    float x, y, size, angle, response;
    int octave;
    int lx, ly;
} oo_keypoint;

/* orb_params (feature/orb_params.h) */
typedef struct {
    uint32_t max_num_keypts;
    float scale_factor;
    uint32_t num_levels;
    uint32_t ini_fast_thr;
    uint32_t min_fast_thr;
} oo_params;

typedef struct {
    int level_w[OO_MAX_LEVELS], level_h[OO_MAX_LEVELS];
    int num_candidates[OO_MAX_LEVELS], num_selected[OO_MAX_LEVELS];
} oo_debug;

void oo_scale_factors(float scale_factor, int num_levels, float* out);
void oo_level_size(int w0, int h0, float scale, int* w, int* h);
void oo_keypts_per_level(unsigned max_num_keypts, float scale_factor, int num_levels, unsigned* out);
int oo_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
int oo_fast_detect(const uint8_t* img, int w, int h, int stride, int threshold, int nonmax, oo_fast_pt* out, int max_out);
void oo_fast_score_map(const uint8_t* img, int w, int h, int stride, uint8_t* score, int sstride);
int oo_distribute_via_tree(const oo_fast_pt* cand, int ncand, int min_x, int max_x, int min_y, int max_y,
                           unsigned num_keypts, int* out_idx);
float oo_fast_atan2(float y, float x);
void oo_umax(int* u_max);
float oo_ic_angle(const uint8_t* img, int stride, int x, int y, int* m01_out, int* m10_out);
void oo_gaussian7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
void oo_sincosf(float a, float* s, float* c);
void oo_orb_descriptor(const uint8_t* blurred, int stride, int x, int y, float angle_deg, uint8_t* desc);
void oo_rect_mask(int cols, int rows, const float* rects, int nrects, uint8_t* mask);
int oo_level_candidates(const oo_params* P, const uint8_t* level_img, int lw, int lh, int lstride, float scale,
                        const uint8_t* mask, int w0, int h0, int mstride, oo_fast_pt** out);
void oo_free(void* p);
int oo_extract(const oo_params* P, const uint8_t* image, int w, int h, int stride,
               const uint8_t* mask, int mstride,
               oo_keypoint* kps, uint8_t* desc, int max_out, oo_debug* dbg);
int oo_build_pyramid(const oo_params* P, const uint8_t* image, int w, int h, int stride, uint8_t** levels);

#ifdef __cplusplus
}
#endif
void oo_color_to_gray(const uint8_t* src, int w, int h, int src_pitch, int channels, int rgb_order, uint8_t* dst, int dst_pitch);
#endif
