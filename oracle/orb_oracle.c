/*
 * oracle/orb_oracle.c -- CPU restatement of OpenVSLAM's ORB front-end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under openvslam_b200/ may include, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS: **parity unpinned** against the real reference.  /root/reference holds
 * no OpenVSLAM source, so "reference file" citations below are to
 * the file names as recalled in SURVEY.md section 8(a) (no line numbers exist to cite).
 * What IS pinned: every OpenCV primitive the reference calls (cv::resize INTER_LINEAR,
 * cv::FAST, cv::GaussianBlur 7x7 sigma 2, cv::fastAtan2, steered-BRIEF bit order) is
 * checked bit-for-bit against cv2 4.13.0 by tests/test_oracle_cv2.py and the committed
 * fixtures in tests/golden/.
 *
 * Functions and the reference code they restate:
 *   oo_resize_linear_u8      cv::resize(..., INTER_LINEAR) for CV_8UC1, as called by
 *                            orb_extractor::compute_image_pyramid (feature/orb_extractor.cc)
 *   oo_fast_detect           cv::FAST(roi, kps, thr, nonmax=true) TYPE_9_16, as called per
 *                            cell by orb_extractor::compute_fast_keypoints
 *   oo_distribute_via_tree   orb_extractor::distribute_keypoints_via_tree /
 *                            initialize_nodes / assign_child_nodes /
 *                            find_keypoints_with_max_response and
 *                            orb_extractor_node::divide_node (feature/orb_extractor_node.cc)
 *   oo_ic_angle              orb_extractor::ic_angle + cv::fastAtan2
 *   oo_gaussian7             cv::GaussianBlur(7x7, 2, 2, BORDER_REFLECT_101) for CV_8UC1
 *   oo_orb_descriptor        orb_extractor::compute_orb_descriptor (+ orb_point_pairs.h)
 *   oo_extract               orb_extractor::extract
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#include "orb_oracle.h"

/* ------------------------------------------------------------------------- helpers */

static int cv_round_f(float v) { return (int)lrintf(v); } /* cvRound: round-half-even */
static int cv_floor_f(float v) {
    int i = (int)v;
    return i - (v < (float)i);
}

/* ------------------------------------------------------------------ scale pyramid  */

/* orb_params::calc_scale_factors (feature/orb_params.cc): float running product. */
void oo_scale_factors(float scale_factor, int num_levels, float* out) {
    out[0] = 1.0f;
    for (int l = 1; l < num_levels; ++l) out[l] = scale_factor * out[l - 1];
}

/* orb_extractor::compute_image_pyramid: size = round(cols * 1.0 / scale). */
void oo_level_size(int w0, int h0, float scale, int* w, int* h) {
    const double s = (double)scale;
    *w = (int)round(w0 * 1.0 / s);
    *h = (int)round(h0 * 1.0 / s);
}

/* orb_extractor::initialize(): geometric split of max_num_keypts over levels. */
void oo_keypts_per_level(unsigned max_num_keypts, float scale_factor, int num_levels, unsigned* out) {
    double desired = max_num_keypts * (1.0 - 1.0 / scale_factor)
                     / (1.0 - pow(1.0 / scale_factor, (double)num_levels));
    unsigned total = 0;
    for (int l = 0; l < num_levels - 1; ++l) {
        out[l] = (unsigned)round(desired);
        total += out[l];
        desired *= 1.0 / scale_factor;
    }
    int rest = (int)max_num_keypts - (int)total;
    out[num_levels - 1] = rest > 0 ? (unsigned)rest : 0u;
}

/* cv::resize, INTER_LINEAR, CV_8UC1, general (non-integer-ratio) path.
 * Fixed point: INTER_RESIZE_COEF_BITS = 11.  Horizontal pass keeps 8.11 ints,
 * vertical pass is ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2. */
int oo_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                        uint8_t* dst, int dw, int dh, int dstride) {
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return -1;
    const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    int* xofs = (int*)malloc(sizeof(int) * dw);
    short* ialpha = (short*)malloc(sizeof(short) * 2 * dw);
    int* rows0 = (int*)malloc(sizeof(int) * dw);
    int* rows1 = (int*)malloc(sizeof(int) * dw);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor_f(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        /* saturate_cast<short>(float) == cvRound */
        ialpha[2 * dx] = (short)cv_round_f((1.f - fx) * 2048.f);
        ialpha[2 * dx + 1] = (short)cv_round_f(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor_f(fy);
        fy -= sy;
        short b0 = (short)cv_round_f((1.f - fy) * 2048.f);
        short b1 = (short)cv_round_f(fy * 2048.f);
        int sy0 = sy, sy1 = sy + 1;
        /* rows are clipped (replicated) at the image border */
        if (sy0 < 0) sy0 = 0;
        if (sy0 > sh - 1) sy0 = sh - 1;
        if (sy1 < 0) sy1 = 0;
        if (sy1 > sh - 1) sy1 = sh - 1;
        const uint8_t* S0 = src + (size_t)sy0 * sstride;
        const uint8_t* S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = xofs[dx];
            const int sx1 = sx + 1 < sw ? sx + 1 : sw - 1; /* alpha1 == 0 there */
            rows0[dx] = S0[sx] * ialpha[2 * dx] + S0[sx1] * ialpha[2 * dx + 1];
            rows1[dx] = S1[sx] * ialpha[2 * dx] + S1[sx1] * ialpha[2 * dx + 1];
        }
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (rows0[dx] >> 4)) >> 16) + ((b1 * (rows1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xofs); free(ialpha); free(rows0); free(rows1);
    return 0;
}

/* ------------------------------------------------------------------------ FAST-9  */

/* Bresenham circle of radius 3, in cv::FAST's order (pixel[0] = (0,3)). */
static const int kCircle[16][2] = {
    {0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
    {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

/* cv::cornerScore<16>: largest t such that the pixel still passes the 9-of-16 arc test
 * with threshold t (given it passes with `threshold`). */
static int corner_score16(const uint8_t* ptr, const int* pixel, int threshold) {
    int d[25];
    const int v = ptr[0];
    for (int k = 0; k < 25; ++k) d[k] = v - ptr[pixel[k % 16]];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = d[k + 1] < d[k + 2] ? d[k + 1] : d[k + 2];
        if (d[k + 3] < a) a = d[k + 3];
        if (a <= a0) continue;
        for (int j = 4; j <= 8; ++j) if (d[k + j] < a) a = d[k + j];
        int t = a < d[k] ? a : d[k];
        if (t > a0) a0 = t;
        t = a < d[k + 9] ? a : d[k + 9];
        if (t > a0) a0 = t;
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = d[k + 1] > d[k + 2] ? d[k + 1] : d[k + 2];
        for (int j = 3; j <= 5; ++j) if (d[k + j] > b) b = d[k + j];
        if (b >= b0) continue;
        for (int j = 6; j <= 8; ++j) if (d[k + j] > b) b = d[k + j];
        int t = b > d[k] ? b : d[k];
        if (t < b0) b0 = t;
        t = b > d[k + 9] ? b : d[k + 9];
        if (t < b0) b0 = t;
    }
    return -b0 - 1;
}

/* The 9-contiguous-of-16 segment test itself (cv::FAST_t<16> inner loop, scalar tail). */
static int is_corner16(const uint8_t* ptr, const int* pixel, int threshold) {
    const int v = ptr[0];
    /* brighter-than-centre run and darker-than-centre run, each over 16+8 wrapped pixels */
    int count = 0;
    for (int k = 0; k < 25; ++k) {
        if (ptr[pixel[k % 16]] < v - threshold) { if (++count > 8) return 1; }
        else count = 0;
    }
    count = 0;
    for (int k = 0; k < 25; ++k) {
        if (ptr[pixel[k % 16]] > v + threshold) { if (++count > 8) return 1; }
        else count = 0;
    }
    return 0;
}

/* cv::FAST(img(roi), keypoints, threshold, nonmax_suppression=true), TYPE_9_16.
 * Only rows/cols [3, size-3) of the ROI are examined; scores of pixels outside that
 * band (or that are not corners at `threshold`) count as 0 in the 3x3 strict-max NMS.
 * Output in cv::FAST's order (row-major).  Returns the number of keypoints written. */
int oo_fast_detect(const uint8_t* img, int w, int h, int stride, int threshold,
                   int nonmax, oo_fast_pt* out, int max_out) {
    if (threshold < 0) threshold = 0;
    if (threshold > 255) threshold = 255;
    int pixel[16];
    for (int k = 0; k < 16; ++k) pixel[k] = kCircle[k][0] + kCircle[k][1] * stride;
    int n = 0;
    if (w < 7 || h < 7) return 0;
    int* score = (int*)calloc((size_t)w * h, sizeof(int));
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = img + (size_t)y * stride + x;
            if (is_corner16(p, pixel, threshold))
                score[y * w + x] = nonmax ? corner_score16(p, pixel, threshold) : 1;
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = score[y * w + x];
            if (!s) continue;
            if (nonmax) {
                const int* c = score + y * w + x;
                if (!(s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1]
                      && s > c[w - 1] && s > c[w] && s > c[w + 1]))
                    continue;
            }
            if (n < max_out) { out[n].x = x; out[n].y = y; out[n].score = nonmax ? s : 0; }
            ++n;
        }
    free(score);
    return n;
}

/* Whole-image score map S(p) = corner_score16 with threshold 0 floor: the largest t>=1
 * for which p is a FAST-9 corner, or 0.  Used by tests to check the CUDA score kernel. */
void oo_fast_score_map(const uint8_t* img, int w, int h, int stride, uint8_t* score, int sstride) {
    int pixel[16];
    for (int k = 0; k < 16; ++k) pixel[k] = kCircle[k][0] + kCircle[k][1] * stride;
    for (int y = 0; y < h; ++y) memset(score + (size_t)y * sstride, 0, w);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = img + (size_t)y * stride + x;
            if (is_corner16(p, pixel, 1)) score[(size_t)y * sstride + x] = (uint8_t)corner_score16(p, pixel, 1);
        }
}

/* ----------------------------------------------------- quadtree keypoint distribution */

typedef struct oo_node {
    int bx, by, ex, ey;           /* pt_begin_, pt_end_ (cv::Point2i) */
    int* idx; int n;              /* indices into the candidate array, insertion order kept */
    int is_leaf;
    long serial;                  /* creation order: stands in for the heap address used by
                                     std::sort on pair<int, node*> to break ties */
    struct oo_node *prev, *next;  /* std::list links */
} oo_node;

typedef struct { oo_node* head; oo_node* tail; int size; long serial; } oo_list;

static oo_node* node_new(oo_list* L, int cap) {
    oo_node* nd = (oo_node*)calloc(1, sizeof(oo_node));
    nd->idx = (int*)malloc(sizeof(int) * (cap > 0 ? cap : 1));
    nd->serial = L->serial++;
    return nd;
}
static void list_push_back(oo_list* L, oo_node* nd) {
    nd->prev = L->tail; nd->next = NULL;
    if (L->tail) L->tail->next = nd; else L->head = nd;
    L->tail = nd; L->size++;
}
static void list_push_front(oo_list* L, oo_node* nd) {
    nd->next = L->head; nd->prev = NULL;
    if (L->head) L->head->prev = nd; else L->tail = nd;
    L->head = nd; L->size++;
}
static oo_node* list_erase(oo_list* L, oo_node* nd) {
    oo_node* nx = nd->next;
    if (nd->prev) nd->prev->next = nd->next; else L->head = nd->next;
    if (nd->next) nd->next->prev = nd->prev; else L->tail = nd->prev;
    L->size--;
    free(nd->idx); free(nd);
    return nx;
}

/* orb_extractor_node::divide_node */
static void divide_node(oo_list* L, const oo_node* nd, const oo_fast_pt* c, oo_node* child[4]) {
    const int half_x = (int)ceil((nd->ex - nd->bx) / 2.0);
    const int half_y = (int)ceil((nd->ey - nd->by) / 2.0);
    for (int k = 0; k < 4; ++k) child[k] = node_new(L, nd->n);
    const int cx = nd->bx + half_x, cy = nd->by + half_y;
    child[0]->bx = nd->bx; child[0]->by = nd->by; child[0]->ex = cx;     child[0]->ey = cy;
    child[1]->bx = cx;     child[1]->by = nd->by; child[1]->ex = nd->ex; child[1]->ey = cy;
    child[2]->bx = nd->bx; child[2]->by = cy;     child[2]->ex = cx;     child[2]->ey = nd->ey;
    child[3]->bx = cx;     child[3]->by = cy;     child[3]->ex = nd->ex; child[3]->ey = nd->ey;
    for (int i = 0; i < nd->n; ++i) {
        const oo_fast_pt* p = &c[nd->idx[i]];
        int k = 0;
        if (cx <= p->x) k += 1;
        if (cy <= p->y) k += 2;
        child[k]->idx[child[k]->n++] = nd->idx[i];
    }
    for (int k = 0; k < 4; ++k) child[k]->is_leaf = (child[k]->n == 1);
}

typedef struct { int count; oo_node* node; } oo_pool_item;

/* orb_extractor::assign_child_nodes */
static void assign_child_nodes(oo_list* L, oo_node* child[4], oo_pool_item* pool, int* npool) {
    for (int k = 0; k < 4; ++k) {
        if (child[k]->n == 0) { free(child[k]->idx); free(child[k]); continue; }
        list_push_front(L, child[k]);
        if (child[k]->n == 1) continue;
        pool[*npool].count = child[k]->n;
        pool[*npool].node = child[k];
        (*npool)++;
    }
}

/* std::sort(pool.rbegin(), pool.rend()) on pair<int, node*>: descending (count, address).
 * The address tie-break is modelled by creation serial (monotone heap). */
static int pool_cmp_desc(const void* a, const void* b) {
    const oo_pool_item* A = (const oo_pool_item*)a; const oo_pool_item* B = (const oo_pool_item*)b;
    if (A->count != B->count) return B->count - A->count;
    if (A->node->serial != B->node->serial) return A->node->serial > B->node->serial ? -1 : 1;
    return 0;
}

/* orb_extractor::distribute_keypoints_via_tree.  Candidate coordinates are relative to
 * (min_x, min_y), as in the reference.  Writes the surviving candidate indices in node
 * list order; returns how many. */
int oo_distribute_via_tree(const oo_fast_pt* cand, int ncand, int min_x, int max_x, int min_y, int max_y,
                           unsigned num_keypts, int* out_idx) {
    if (ncand == 0) return 0;
    oo_list L = {0, 0, 0, 0};
    /* initialize_nodes */
    const double ratio = (double)(max_x - min_x) / (max_y - min_y);
    double delta_x, delta_y; unsigned gx, gy;
    if (ratio > 1) { gx = (unsigned)round(ratio); gy = 1; delta_x = (double)(max_x - min_x) / gx; delta_y = max_y - min_y; }
    else { gx = 1; gy = (unsigned)round(1 / ratio); delta_x = max_x - min_x; delta_y = (double)(max_y - min_y) / gy; }
    const unsigned nini = gx * gy;
    oo_node** ini = (oo_node**)malloc(sizeof(oo_node*) * nini);
    for (unsigned i = 0; i < nini; ++i) {
        oo_node* nd = node_new(&L, ncand);
        const unsigned ix = i % gx, iy = i / gx;
        nd->bx = (int)(delta_x * ix); nd->by = (int)(delta_y * iy);
        nd->ex = (int)(delta_x * (ix + 1)); nd->ey = (int)(delta_y * (iy + 1));
        list_push_back(&L, nd); ini[i] = nd;
    }
    for (int i = 0; i < ncand; ++i) {
        /* keypt.pt is float; x / delta_x is evaluated in double, truncated to unsigned */
        unsigned ix = (unsigned)((float)cand[i].x / delta_x);
        unsigned iy = (unsigned)((float)cand[i].y / delta_y);
        unsigned k = ix + iy * gx;
        if (k >= nini) k = nini - 1; /* cannot happen for in-range candidates */
        ini[k]->idx[ini[k]->n++] = i;
    }
    free(ini);
    for (oo_node* it = L.head; it;) {
        if (it->n == 0) { it = list_erase(&L, it); continue; }
        it->is_leaf = (it->n == 1);
        it = it->next;
    }

    oo_pool_item* pool = (oo_pool_item*)malloc(sizeof(oo_pool_item) * ((size_t)ncand + 8));
    oo_pool_item* prev_pool = (oo_pool_item*)malloc(sizeof(oo_pool_item) * ((size_t)ncand + 8));
    int npool = 0, is_filled = 0;
    for (;;) {
        const int prev_size = L.size;
        npool = 0;
        for (oo_node* it = L.head; it;) {
            if (it->is_leaf) { it = it->next; continue; }
            oo_node* child[4];
            divide_node(&L, it, cand, child);
            assign_child_nodes(&L, child, pool, &npool);
            it = list_erase(&L, it);
        }
        if ((int)num_keypts <= L.size || L.size == prev_size) { is_filled = 1; break; }
        if ((long)num_keypts < (long)L.size + 3L * npool) { is_filled = 0; break; }
    }
    while (!is_filled) {
        const int prev_size = L.size;
        const int nprev = npool;
        memcpy(prev_pool, pool, sizeof(oo_pool_item) * nprev);
        npool = 0;
        qsort(prev_pool, nprev, sizeof(oo_pool_item), pool_cmp_desc);
        for (int i = 0; i < nprev; ++i) {
            oo_node* child[4];
            divide_node(&L, prev_pool[i].node, cand, child);
            assign_child_nodes(&L, child, pool, &npool);
            list_erase(&L, prev_pool[i].node);
            if ((int)num_keypts <= L.size) { is_filled = 1; break; }
        }
        if (is_filled || (int)num_keypts <= L.size || L.size == prev_size) { is_filled = 1; break; }
    }
    /* find_keypoints_with_max_response: first strict maximum per node, list order */
    int nout = 0;
    for (oo_node* it = L.head; it; it = it->next) {
        int best = it->idx[0];
        for (int k = 1; k < it->n; ++k)
            if (cand[it->idx[k]].score > cand[best].score) best = it->idx[k];
        out_idx[nout++] = best;
    }
    for (oo_node* it = L.head; it;) it = list_erase(&L, it);
    free(pool); free(prev_pool);
    return nout;
}

/* ----------------------------------------------------------------- orientation */

/* cv::fastAtan2 (scalar path, degrees).  All arithmetic in IEEE float, no FMA. */
float oo_fast_atan2(float y, float x) {
    static const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
    static const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    static const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
    static const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    volatile float ax = fabsf(x), ay = fabsf(y);
    volatile float a, c, c2, t;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
        a = t * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
        t = t * c;
        a = 90.f - t;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* orb_extractor::initialize(): u_max_ table of the circular patch, half size 15. */
void oo_umax(int* u_max /* [16] */) {
    const int hp = 15;
    const int vmax = (int)floor(hp * sqrt(2.0) / 2 + 1);
    const int vmin = (int)ceil(hp * sqrt(2.0) / 2);
    for (int v = 0; v <= vmax; ++v) u_max[v] = (int)round(sqrt((double)hp * hp - (double)v * v));
    for (int v = hp, v0 = 0; v >= vmin; --v) {
        while (u_max[v0] == u_max[v0 + 1]) ++v0;
        u_max[v] = v0;
        ++v0;
    }
}

/* orb_extractor::ic_angle: intensity-centroid moments over the circular patch. */
float oo_ic_angle(const uint8_t* img, int stride, int x, int y, int* m01_out, int* m10_out) {
    int u_max[16];
    oo_umax(u_max);
    int m_01 = 0, m_10 = 0;
    const uint8_t* center = img + (size_t)y * stride + x;
    for (int u = -15; u <= 15; ++u) m_10 += u * center[u];
    for (int v = 1; v <= 15; ++v) {
        int v_sum = 0;
        const int d = u_max[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * stride];
            const int val_minus = center[u - v * stride];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    if (m01_out) *m01_out = m_01;
    if (m10_out) *m10_out = m_10;
    return oo_fast_atan2((float)m_01, (float)m_10);
}

/* ----------------------------------------------------------------------- blur  */

/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1:
 * OpenCV's bit-exact fixed-point path, 8.8 kernel {18,34,48,56,48,34,18}/256,
 * exact 16.16 accumulation, one rounding at the end. */
static int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}
void oo_gaussian7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    uint16_t* hbuf = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int s = 0;
            for (int i = -3; i <= 3; ++i) s += k[i + 3] * src[(size_t)y * sstride + reflect101(x + i, w)];
            hbuf[(size_t)y * w + x] = (uint16_t)s;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint32_t s = 0;
            for (int j = -3; j <= 3; ++j) s += (uint32_t)k[j + 3] * hbuf[(size_t)reflect101(y + j, h) * w + x];
            dst[(size_t)y * dstride + x] = (uint8_t)((s + 32768u) >> 16);
        }
    free(hbuf);
}

/* --------------------------------------------------------------- descriptors  */

static const int8_t kPattern[256][4] = {
#include "orb_pattern.inc"
};

/* orb_extractor::compute_orb_descriptor.  angle in degrees; the reference converts with
 * `keypt.angle * M_PI / 180.0` (double) narrowed to float, then util::cos/sin on float
 * (std::cos(float) == cosf when USE_SSE_FP_MATH is off, the default build). */
void oo_orb_descriptor(const uint8_t* blurred, int stride, int x, int y, float angle_deg, uint8_t* desc) {
    const float angle = (float)((double)angle_deg * 3.14159265358979323846 / 180.0);
    float cos_angle, sin_angle;
    oo_sincosf(angle, &sin_angle, &cos_angle);
    const uint8_t* center = blurred + (size_t)y * stride + x;
    for (int i = 0; i < 32; ++i) {
        int byte = 0;
        for (int b = 0; b < 8; ++b) {
            const int8_t* p = kPattern[i * 8 + b];
            volatile float r0a = p[0] * sin_angle, r0b = p[1] * cos_angle;
            volatile float c0a = p[0] * cos_angle, c0b = p[1] * sin_angle;
            volatile float r1a = p[2] * sin_angle, r1b = p[3] * cos_angle;
            volatile float c1a = p[2] * cos_angle, c1b = p[3] * sin_angle;
            const int t0 = center[cv_round_f(r0a + r0b) * stride + cv_round_f(c0a - c0b)];
            const int t1 = center[cv_round_f(r1a + r1b) * stride + cv_round_f(c1a - c1b)];
            byte |= (t0 < t1) << b;
        }
        desc[i] = (uint8_t)byte;
    }
}

/* sin/cos of a float angle (radians), correctly rounded to float via double libm.
 * glibc's cosf/sinf are documented < 1 ULP, not always correctly rounded; the oracle
 * defines the value as float(cos((double)a)), and tests/test_oracle_cv2.py measures how
 * often that differs from this host's cosf/sinf (see DESIGN.md "sin/cos"). */
void oo_sincosf(float a, float* s, float* c) {
    *s = (float)sin((double)a);
    *c = (float)cos((double)a);
}

/* ---------------------------------------------------------------- full extract */

static int mask_is_zero(const uint8_t* mask, int mw, int mh, int mstride, unsigned y, unsigned x, float scale) {
    int my = (int)(y * scale), mx = (int)(x * scale);
    if (my >= mh) my = mh - 1;
    if (mx >= mw) mx = mw - 1;
    return mask[(size_t)my * mstride + mx] == 0;
}

/* orb_extractor::create_rectangle_mask: rects are {x_min, x_max, y_min, y_max} in [0,1]. */
void oo_rect_mask(int cols, int rows, const float* rects, int nrects, uint8_t* mask) {
    memset(mask, 255, (size_t)cols * rows);
    for (int r = 0; r < nrects; ++r) {
        const float* q = rects + 4 * r;
        const unsigned x0 = (unsigned)(cols * q[0]), x1 = (unsigned)(cols * q[1]);
        const unsigned y0 = (unsigned)(rows * q[2]), y1 = (unsigned)(rows * q[3]);
        for (unsigned y = y0; y < y1 && y < (unsigned)rows; ++y)
            for (unsigned x = x0; x < x1 && x < (unsigned)cols; ++x) mask[(size_t)y * cols + x] = 0;
    }
}

/* orb_extractor::compute_fast_keypoints, the per-level cell loop: cv::FAST per 64 px cell
 * (6 px overlap, 19 px border) with the ini -> min threshold fallback and the mask tests.
 * Returns the candidates (coordinates relative to the border) in the reference's order;
 * *out is malloc'd. */
int oo_level_candidates(const oo_params* P, const uint8_t* level_img, int lw, int lh, int lstride, float scale,
                        const uint8_t* mask, int w0, int h0, int mstride, oo_fast_pt** out) {
    enum { overlap = 6, cell = 64, border = 19 };
    *out = NULL;
    if (lw <= 2 * border || lh <= 2 * border) return 0;
    const unsigned min_bx = border, min_by = border;
    const unsigned max_bx = lw - border, max_by = lh - border;
    const unsigned width = max_bx - min_bx, height = max_by - min_by;
    /* std::ceil(width / cell_size) on unsigned operands: integer division first */
    const unsigned num_cols = (unsigned)ceil((double)(width / cell)) + 1;
    const unsigned num_rows = (unsigned)ceil((double)(height / cell)) + 1;
    size_t cap = 1024, nc = 0;
    oo_fast_pt* cand = (oo_fast_pt*)malloc(sizeof(oo_fast_pt) * cap);
    oo_fast_pt* cell_pts = (oo_fast_pt*)malloc(sizeof(oo_fast_pt) * 70 * 70);
    for (unsigned i = 0; i < num_rows; ++i) {
        const unsigned min_y = min_by + i * cell;
        if (max_by - overlap <= min_y) continue;
        unsigned max_y = min_y + cell + overlap;
        if (max_by < max_y) max_y = max_by;
        for (unsigned j = 0; j < num_cols; ++j) {
            const unsigned min_x = min_bx + j * cell;
            if (max_bx - overlap <= min_x) continue;
            unsigned max_x = min_x + cell + overlap;
            if (max_bx < max_x) max_x = max_bx;
            if (mask) {
                if (mask_is_zero(mask, w0, h0, mstride, min_y, min_x, scale) || mask_is_zero(mask, w0, h0, mstride, max_y, min_x, scale)
                    || mask_is_zero(mask, w0, h0, mstride, min_y, max_x, scale) || mask_is_zero(mask, w0, h0, mstride, max_y, max_x, scale))
                    continue;
            }
            const uint8_t* roi = level_img + (size_t)min_y * lstride + min_x;
            int n = oo_fast_detect(roi, (int)(max_x - min_x), (int)(max_y - min_y), lstride, (int)P->ini_fast_thr, 1, cell_pts, 70 * 70);
            if (n == 0)
                n = oo_fast_detect(roi, (int)(max_x - min_x), (int)(max_y - min_y), lstride, (int)P->min_fast_thr, 1, cell_pts, 70 * 70);
            for (int k = 0; k < n; ++k) {
                oo_fast_pt p = cell_pts[k];
                p.x += (int)(j * cell); p.y += (int)(i * cell);
                if (mask && mask_is_zero(mask, w0, h0, mstride, (unsigned)(float)(min_by + (float)p.y), (unsigned)(float)(min_bx + (float)p.x), scale))
                    continue;
                if (nc == cap) { cap *= 2; cand = (oo_fast_pt*)realloc(cand, sizeof(oo_fast_pt) * cap); }
                cand[nc++] = p;
            }
        }
    }
    free(cell_pts);
    *out = cand;
    return (int)nc;
}

void oo_free(void* p) { free(p); }

int oo_extract(const oo_params* P, const uint8_t* image, int w, int h, int stride,
               const uint8_t* mask, int mstride,
               oo_keypoint* kps, uint8_t* desc, int max_out, oo_debug* dbg) {
    const int L = (int)P->num_levels;
    if (L < 1 || L > OO_MAX_LEVELS) return -1;
    float sf[OO_MAX_LEVELS];
    unsigned per_level[OO_MAX_LEVELS];
    oo_scale_factors(P->scale_factor, L, sf);
    oo_keypts_per_level(P->max_num_keypts, P->scale_factor, L, per_level);

    uint8_t* pyr[OO_MAX_LEVELS]; int pw[OO_MAX_LEVELS], ph[OO_MAX_LEVELS];
    pw[0] = w; ph[0] = h;
    pyr[0] = (uint8_t*)malloc((size_t)w * h);
    for (int y = 0; y < h; ++y) memcpy(pyr[0] + (size_t)y * w, image + (size_t)y * stride, w);
    for (int l = 1; l < L; ++l) {
        oo_level_size(w, h, sf[l], &pw[l], &ph[l]);
        pyr[l] = (uint8_t*)malloc((size_t)pw[l] * ph[l]);
        oo_resize_linear_u8(pyr[l - 1], pw[l - 1], ph[l - 1], pw[l - 1], pyr[l], pw[l], ph[l], pw[l]);
    }

    enum { border = 19 };
    int total = 0;
    for (int l = 0; l < L; ++l) {
        if (dbg) { dbg->level_w[l] = pw[l]; dbg->level_h[l] = ph[l]; dbg->num_candidates[l] = 0; dbg->num_selected[l] = 0; }
        if (pw[l] <= 2 * border || ph[l] <= 2 * border) continue;
        const unsigned min_bx = border, min_by = border;
        const unsigned max_bx = pw[l] - border, max_by = ph[l] - border;
        oo_fast_pt* cand = NULL;
        const size_t nc = (size_t)oo_level_candidates(P, pyr[l], pw[l], ph[l], pw[l], sf[l], mask, w, h, mstride, &cand);
        int* sel = (int*)malloc(sizeof(int) * (nc + 1));
        const int ns = oo_distribute_via_tree(cand, (int)nc, (int)min_bx, (int)max_bx, (int)min_by, (int)max_by, per_level[l], sel);
        if (dbg) { dbg->num_candidates[l] = (int)nc; dbg->num_selected[l] = ns; }

        uint8_t* blurred = NULL;
        if (ns > 0 && desc) {
            blurred = (uint8_t*)malloc((size_t)pw[l] * ph[l]);
            oo_gaussian7(pyr[l], pw[l], ph[l], pw[l], blurred, pw[l]);
        }
        const unsigned scaled_patch_size = (unsigned)(31 * sf[l]);
        for (int k = 0; k < ns; ++k) {
            const oo_fast_pt* p = &cand[sel[k]];
            const int lx = p->x + (int)min_bx, ly = p->y + (int)min_by;
            const float angle = oo_ic_angle(pyr[l], pw[l], lx, ly, NULL, NULL);
            if (total < max_out) {
                oo_keypoint* q = &kps[total];
                /* correct_keypoint_scale: pt *= scale_factor (float), level 0 untouched */
                q->x = l == 0 ? (float)lx : (float)lx * sf[l];
                q->y = l == 0 ? (float)ly : (float)ly * sf[l];
                q->size = (float)scaled_patch_size;
                q->angle = angle;
                q->response = (float)p->score;
                q->octave = l;
                q->lx = lx; q->ly = ly;
                if (desc) oo_orb_descriptor(blurred, pw[l], lx, ly, angle, desc + (size_t)total * 32);
            }
            ++total;
        }
        free(blurred); free(sel); free(cand);
    }
    for (int l = 0; l < L; ++l) free(pyr[l]);
    return total;
}

/* Build only the pyramid (tests compare each level with the CUDA pyramid). */
int oo_build_pyramid(const oo_params* P, const uint8_t* image, int w, int h, int stride, uint8_t** levels /* caller-allocated */) {
    const int L = (int)P->num_levels;
    float sf[OO_MAX_LEVELS];
    oo_scale_factors(P->scale_factor, L, sf);
    int pw = w, ph = h;
    for (int y = 0; y < h; ++y) memcpy(levels[0] + (size_t)y * w, image + (size_t)y * stride, w);
    for (int l = 1; l < L; ++l) {
        int nw, nh;
        oo_level_size(w, h, sf[l], &nw, &nh);
        oo_resize_linear_u8(levels[l - 1], pw, ph, pw, levels[l], nw, nh, nw);
        pw = nw; ph = nh;
    }
    return 0;
}


/* util::convert_to_grayscale (util/image_converter.cc) = cv::cvtColor(img, {BGR,RGB,BGRA,RGBA}2GRAY) on CV_8U: OpenCV's
 * 15-bit fixed-point weights (B 3735, G 19235, R 9798, rounding 1 << 14), pinned bit-for-bit against cv2 4.13.0.
 * (OpenCV 3.x used 14-bit weights 1868 / 9617 / 4899, which differ by one grey level on 0.27 % of random pixels.)
 * channels: 3 or 4 (alpha ignored); rgb_order: 0 = B first (BGR, BGRA), 1 = R first (RGB, RGBA). */
void oo_color_to_gray(const uint8_t* src, int w, int h, int src_pitch, int channels, int rgb_order, uint8_t* dst, int dst_pitch) {
    for (int y = 0; y < h; ++y) {
        const uint8_t* p = src + (size_t)y * src_pitch;
        uint8_t* d = dst + (size_t)y * dst_pitch;
        for (int x = 0; x < w; ++x, p += channels) {
            const int b = rgb_order ? p[2] : p[0], g = p[1], r = rgb_order ? p[0] : p[2];
            d[x] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15);
        }
    }
}
