// match_window.cu -- grid-windowed Hamming search for openvslam::match::{projection, area} and the
// rectified-row search + SAD sub-pixel refinement of openvslam::match::stereo
// (match/projection.cc, match/area.cc, match/stereo.cc, data/frame.cc get_keypoints_in_cell,
// data/common.cc assign_keypoints_to_grid; names as in SURVEY.md 8a a9, a10, a12).
//
// Frame index (ovs_frame_index): the frame's keypoints re-ordered by (cell_x, cell_y, index) -- the
// order get_keypoints_in_cell visits them -- with a CSR of cell starts, resident on the device.
//
// k_window_topk     one warp per query (landmark / keypoint): the cell range and the box, level,
//                   x_right and per-keypoint distance-cap tests exactly as the reference, 256-bit
//                   Hamming by __popc, per-lane sorted top-4 of (distance << 16 | rank), merged
//                   across the warp with __reduce_min_sync.  Rank order == candidate order, so the
//                   sorted keys reproduce the reference's first-wins tie-breaking.
// k_stereo_match    one thread per left keypoint, right keypoints staged through shared memory:
//                   row band / octave / disparity tests + Hamming -> best right keypoint.
// k_stereo_subpixel one warp per left keypoint: 11 SAD windows (11 x 11, centre-normalised) on the
//                   extractor's device pyramids, warp-reduced with __reduce_add_sync, parabola fit.
//
// The sequential parts of the reference (greedy "a keypoint is matched once" bookkeeping, the angle
// histogram, the median test of stereo) run on the host over these results.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "match_common.h"

// ------------------------------------------------------------------------------- frame index
struct ovs_frame_index {
    ovs_matcher* m = nullptr;
    int n = 0, nranked = 0;
    ovs_grid grid{};
    std::vector<int> rank_to_idx, idx_to_rank;   // host
    std::vector<float> hx, hy, hxr, hangle; std::vector<int> hoct;
    // device, rank order
    float* d_x = nullptr; float* d_y = nullptr; float* d_xr = nullptr; signed char* d_oct = nullptr;
    uint4* d_desc = nullptr; int* d_cell_start = nullptr; unsigned short* d_cap = nullptr; int* d_rank = nullptr;
    ovs_index_buf buf;          // all of the above are carved from this one allocation (recycled through the matcher's pool)
    bool has_xr = false;
};

namespace {

constexpr int kTopK = 4;

__device__ __forceinline__ void topk_insert(unsigned (&k)[kTopK], unsigned key) {
    if (key < k[3]) {
        k[3] = key;
        if (k[3] < k[2]) { const unsigned t = k[2]; k[2] = k[3]; k[3] = t; }
        if (k[2] < k[1]) { const unsigned t = k[1]; k[1] = k[2]; k[2] = t; }
        if (k[1] < k[0]) { const unsigned t = k[0]; k[0] = k[1]; k[1] = t; }
    }
}

__device__ __forceinline__ int cv_floor_f(float v) { return __float2int_rd(v); }
__device__ __forceinline__ int cv_ceil_f(float v) { return __float2int_ru(v); }

struct WindowFrame {
    float min_x, min_y, inv_w, inv_h;
    int cols, rows;
    const float* x; const float* y; const float* xr; const signed char* oct; const uint4* desc;
    const int* cell_start; const unsigned short* cap;
};

struct WindowQueries {
    int nq;
    const float2* ref; const float* margin; const int* min_level; const int* max_level; const float* xr; const uint4* desc;
    // match::fuse: instead of the x_right window test, a candidate is skipped when its reprojection error exceeds the
    // chi-square bound of its own octave (5.99 monocular, 7.8 with the x_right term)
    int fuse_gate;
    float inv_sigma_sq[16];
};

__global__ void __launch_bounds__(128) k_window_topk(WindowFrame F, WindowQueries Q, unsigned* __restrict__ out) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= Q.nq) return;
    const float2 ref = Q.ref[q];
    const float margin = Q.margin[q];
    const int min_level = Q.min_level[q], max_level = Q.max_level[q];
    unsigned best[kTopK] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    // data::frame::get_keypoints_in_cell: cell range (float arithmetic as in the reference)
    const int min_cx = max(0, cv_floor_f(__fmul_rn(__fsub_rn(__fsub_rn(ref.x, F.min_x), margin), F.inv_w)));
    const int max_cx = min(F.cols - 1, cv_ceil_f(__fmul_rn(__fadd_rn(__fsub_rn(ref.x, F.min_x), margin), F.inv_w)));
    const int min_cy = max(0, cv_floor_f(__fmul_rn(__fsub_rn(__fsub_rn(ref.y, F.min_y), margin), F.inv_h)));
    const int max_cy = min(F.rows - 1, cv_ceil_f(__fmul_rn(__fadd_rn(__fsub_rn(ref.y, F.min_y), margin), F.inv_h)));
    if (F.cols > min_cx && max_cx >= 0 && F.rows > min_cy && max_cy >= 0) {
        const uint4 qa = __ldg(Q.desc + 2 * (size_t)q), qb = __ldg(Q.desc + 2 * (size_t)q + 1);
        const bool check_level = (0 < min_level) || (0 <= max_level);
        const float xr_q = Q.xr ? Q.xr[q] : -1.0f;
        for (int cx = min_cx; cx <= max_cx; ++cx) {
            const int r0 = F.cell_start[cx * F.rows + min_cy], r1 = F.cell_start[cx * F.rows + max_cy + 1];
            for (int r = r0 + lane; r < r1; r += 32) {
                if (check_level) {
                    const int o = F.oct[r];
                    if (o < min_level) continue;
                    if (0 <= max_level && max_level < o) continue;
                }
                const float dist_x = __fsub_rn(F.x[r], ref.x), dist_y = __fsub_rn(F.y[r], ref.y);
                if (!(fabsf(dist_x) < margin && fabsf(dist_y) < margin)) continue;
                if (Q.fuse_gate) {
                    const float ex = __fsub_rn(ref.x, F.x[r]), ey = __fsub_rn(ref.y, F.y[r]);
                    const float w = Q.inv_sigma_sq[F.oct[r] & 15];
                    const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                    const float kxr = F.xr ? F.xr[r] : -1.0f;
                    if (kxr >= 0 && Q.xr) {
                        const float er = __fsub_rn(xr_q, kxr);
                        if (__fmul_rn(__fadd_rn(e2, __fmul_rn(er, er)), w) > 7.8f) continue;
                    } else {
                        if (__fmul_rn(e2, w) > 5.99f) continue;
                    }
                } else if (F.xr) {
                    const float kxr = F.xr[r];
                    if (0 < kxr) {
                        const float reproj_error = fabsf(__fsub_rn(xr_q, kxr));
                        if (margin < reproj_error) continue;
                    }
                }
                const uint4 ta = __ldg(F.desc + 2 * (size_t)r), tb = __ldg(F.desc + 2 * (size_t)r + 1);
                const int d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w)
                              + __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
                if (F.cap && !(d < (int)F.cap[r])) continue;
                topk_insert(best, ((unsigned)d << 16) | (unsigned)r);
            }
        }
    }
    // merge the 32 sorted lists: 4 rounds of warp-wide minimum
#pragma unroll
    for (int k = 0; k < kTopK; ++k) {
        const unsigned m = __reduce_min_sync(0xffffffffu, best[0]);
        if (best[0] == m && m != 0xffffffffu) { best[0] = best[1]; best[1] = best[2]; best[2] = best[3]; best[3] = 0xffffffffu; }
        if (lane == 0) out[(size_t)q * kTopK + k] = m;
    }
}

// ------------------------------------------------------------------------------------ stereo
struct StereoArgs {
    int n_left, n_right;
    const float* lx; const float* ly; const int* loct; const uint4* ldesc;
    const float* rx; const float* ry; const int* roct; const uint4* rdesc;
    float min_disp, max_disp; int hamm_thr; int rows0;
    float scale[16], inv_scale[16];
    const uint8_t* lpyr[16]; const uint8_t* rpyr[16]; int pw[16], ph[16], pitch[16];
    float focal_x_baseline;
};

// best[il] = (dist << 16 | idx_right) or 0xFFFFFFFF; flag bit 31 of any[il] unused.
__global__ void __launch_bounds__(128) k_stereo_match(StereoArgs A, unsigned* __restrict__ best_out) {
    __shared__ float s_rx[256], s_lo[256], s_hi[256];
    __shared__ int s_oct[256];
    __shared__ uint4 s_desc[512];
    const int il = blockIdx.x * 128 + threadIdx.x;
    const bool act = il < A.n_left;
    float x_left = 0, y_left = 0; int lvl = 0; uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (act) { x_left = A.lx[il]; y_left = A.ly[il]; lvl = A.loct[il]; qa = __ldg(A.ldesc + 2 * (size_t)il); qb = __ldg(A.ldesc + 2 * (size_t)il + 1); }
    const int row = (int)y_left;
    const float min_x_right = __fsub_rn(x_left, A.max_disp), max_x_right = __fsub_rn(x_left, A.min_disp);
    const bool usable = act && row >= 0 && row < A.rows0 && !(max_x_right < 0);
    unsigned best = ((unsigned)A.hamm_thr << 16) | 0xffffu;   // sentinel: distance == threshold
    for (int t0 = 0; t0 < A.n_right; t0 += 256) {
        const int n = min(256, A.n_right - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += 128) {
            const int ir = t0 + i;
            const int o = A.roct[ir];
            const float r = __fmul_rn(2.0f, A.scale[o]);
            const float yy = A.ry[ir];
            s_rx[i] = A.rx[ir]; s_oct[i] = o;
            s_lo[i] = (float)cv_floor_f(__fsub_rn(yy, r)); s_hi[i] = (float)cv_ceil_f(__fadd_rn(yy, r));
            s_desc[2 * i] = __ldg(A.rdesc + 2 * (size_t)ir); s_desc[2 * i + 1] = __ldg(A.rdesc + 2 * (size_t)ir + 1);
        }
        __syncthreads();
        if (!usable) continue;
        for (int j = 0; j < n; ++j) {
            if ((float)row < s_lo[j] || (float)row > s_hi[j]) continue;
            const int o = s_oct[j];
            if (o < lvl - 1 || o > lvl + 1) continue;
            const float xr = s_rx[j];
            if (xr < min_x_right || max_x_right < xr) continue;
            const uint4 ta = s_desc[2 * j], tb = s_desc[2 * j + 1];
            const int d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w)
                          + __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
            const unsigned key = ((unsigned)d << 16) | (unsigned)(t0 + j);
            if ((key >> 16) < (best >> 16)) best = key;   // strict '<' on the distance: first (lowest index) wins ties
        }
    }
    if (act) best_out[il] = (usable && (best >> 16) < (unsigned)A.hamm_thr) ? best : 0xffffffffu;
}

// One warp per left keypoint with a Hamming match: SAD over 11 horizontal offsets, parabola fit,
// disparity checks.  Writes x_right / depth (-1 if rejected) and the best SAD (for the median test).
__global__ void __launch_bounds__(128) k_stereo_subpixel(StereoArgs A, const unsigned* __restrict__ best_in, float* __restrict__ x_right_out,
                                                          float* __restrict__ depth_out, unsigned* __restrict__ corr_out) {
    const int il = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (il >= A.n_left) return;
    float xr_res = -1.0f, depth_res = -1.0f; unsigned corr_res = 0xffffffffu;
    const unsigned key = best_in[il];
    if (key != 0xffffffffu) {
        const int ir = (int)(key & 0xffffu);
        const int lvl = A.loct[il];
        const float x_left = A.lx[il], y_left = A.ly[il];
        const float inv_s = A.inv_scale[lvl];
        const int sxl = __float2int_rn(__fmul_rn(x_left, inv_s)), syl = __float2int_rn(__fmul_rn(y_left, inv_s));
        const int sxr = __float2int_rn(__fmul_rn(A.rx[ir], inv_s));
        constexpr int win = 5, slide = 5;
        const int W = A.pw[lvl], Hh = A.ph[lvl], S = A.pitch[lvl];
        const int ini_x = sxr - slide - win, end_x = sxr + slide + win + 1;
        const bool fits = !(ini_x < 0 || W <= end_x) && !(sxl - win < 0 || W <= sxl + win || syl - win < 0 || Hh <= syl + win);
        if (fits) {
            const uint8_t* L = A.lpyr[lvl]; const uint8_t* R = A.rpyr[lvl];
            const int cl = L[(size_t)syl * S + sxl];
            // each lane owns up to 4 of the 121 window pixels
            int lv[4], py[4], px[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = lane + 32 * k;
                py[k] = p / 11 - win; px[k] = p % 11 - win;
                lv[k] = (p < 121) ? (int)L[(size_t)(syl + py[k]) * S + sxl + px[k]] - cl : 0;
            }
            unsigned best_corr = 0xffffffffu; int best_off = 0;
            unsigned sads[2 * slide + 1];
#pragma unroll
            for (int off = -slide; off <= slide; ++off) {
                const int cr = R[(size_t)syl * S + sxr + off];
                unsigned sad = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int p = lane + 32 * k;
                    if (p < 121) {
                        const int b = (int)R[(size_t)(syl + py[k]) * S + sxr + off + px[k]] - cr;
                        sad += (unsigned)abs(lv[k] - b);
                    }
                }
                sad = __reduce_add_sync(0xffffffffu, sad);
                sads[off + slide] = sad;
                if (sad < best_corr) { best_corr = sad; best_off = off; }
            }
            if (!(best_off == -slide || best_off == slide)) {
                float c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
                for (int k = 1; k < 2 * slide; ++k)
                    if (k == best_off + slide) { c1 = (float)sads[k - 1]; c2 = (float)sads[k]; c3 = (float)sads[k + 1]; }
                const float delta = (float)((double)__fsub_rn(c1, c3) / (2.0 * ((double)__fadd_rn(c1, c3) - 2.0 * (double)c2)));
                if (!(delta < -1.0f || 1.0f < delta)) {
                    float best_x_right = __fmul_rn(A.scale[lvl], __fadd_rn((float)(sxr + best_off), delta));
                    float disp = __fsub_rn(x_left, best_x_right);
                    if (!(disp < A.min_disp || A.max_disp <= disp)) {
                        if (disp <= 0.0f) { disp = 0.01f; best_x_right = __fsub_rn(x_left, disp); }
                        depth_res = __fdiv_rn(A.focal_x_baseline, disp);
                        xr_res = best_x_right;
                        corr_res = best_corr;
                    }
                }
            }
        }
    }
    if (lane == 0) { x_right_out[il] = xr_res; depth_out[il] = depth_res; corr_out[il] = corr_res; }
}

inline int key_dist(unsigned key) { return key == 0xffffffffu ? OVS_MAX_HAMMING_DIST : (int)(key >> 16); }
inline int key_rank(unsigned key) { return key == 0xffffffffu ? -1 : (int)(key & 0xffffu); }

// data::get_cell_indices
inline bool cell_of(const ovs_grid& g, float x, float y, int* cx, int* cy) {
    *cx = (int)std::floor((double)((x - g.min_x) * g.inv_cell_width));
    *cy = (int)std::floor((double)((y - g.min_y) * g.inv_cell_height));
    return 0 <= *cx && *cx < g.num_grid_cols && 0 <= *cy && *cy < g.num_grid_rows;
}

// Runs k_window_topk for nq host queries; keys land in m->h_keys[0 .. nq*4).  `cap` (rank order, n
// entries) or nullptr.
int window_topk(ovs_frame_index* f, int nq, const float* ref_xy, const float* margin, const int* min_level, const int* max_level,
                const float* xr_q, const uint8_t* qdesc, const unsigned short* cap_rank_order, const float* fuse_inv_sigma_sq = nullptr,
                int fuse_levels = 0) {
    ovs_matcher* m = f->m;
    cudaStream_t st = m->stream;
    const size_t N = (size_t)nq;
    const size_t bytes = N * (8 + 4 + 4 + 4 + 4 + 32) + 512;
    int rc;
    if ((rc = ovs::grow_host(&m->h_stage, &m->h_stage_cap, bytes)) != OVS_OK) return rc;
    if ((rc = ovs::grow_dev(&m->d_q, &m->d_q_cap, bytes)) != OVS_OK) return rc;
    if ((rc = ovs::grow_dev(&m->d_keys, &m->d_keys_cap, (N + 1) * kTopK)) != OVS_OK) return rc;
    if ((rc = ovs::grow_host(&m->h_keys, &m->h_keys_cap, (N + 1) * kTopK)) != OVS_OK) return rc;
    // carve: desc (32 N, 16-aligned first), ref (8 N), margin, min, max, xr (4 N each)
    size_t off = 0;
    auto carve = [&](size_t b) { const size_t o = off; off += (b + 15) / 16 * 16; return o; };
    const size_t o_desc = carve(32 * N), o_ref = carve(8 * N), o_m = carve(4 * N), o_lo = carve(4 * N), o_hi = carve(4 * N), o_xr = carve(4 * N);
    memcpy(m->h_stage + o_desc, qdesc, 32 * N); memcpy(m->h_stage + o_ref, ref_xy, 8 * N); memcpy(m->h_stage + o_m, margin, 4 * N);
    memcpy(m->h_stage + o_lo, min_level, 4 * N); memcpy(m->h_stage + o_hi, max_level, 4 * N);
    if (xr_q) memcpy(m->h_stage + o_xr, xr_q, 4 * N);
    OVS_CUDA_CHECK(cudaMemcpyAsync(m->d_q, m->h_stage, off, cudaMemcpyHostToDevice, st));
    if (cap_rank_order) OVS_CUDA_CHECK(cudaMemcpyAsync(f->d_cap, cap_rank_order, 2 * (size_t)std::max(f->nranked, 1), cudaMemcpyHostToDevice, st));
    WindowFrame F;
    F.min_x = f->grid.min_x; F.min_y = f->grid.min_y; F.inv_w = f->grid.inv_cell_width; F.inv_h = f->grid.inv_cell_height;
    F.cols = f->grid.num_grid_cols; F.rows = f->grid.num_grid_rows;
    F.x = f->d_x; F.y = f->d_y; F.xr = f->has_xr ? f->d_xr : nullptr; F.oct = f->d_oct; F.desc = f->d_desc;
    F.cell_start = f->d_cell_start; F.cap = cap_rank_order ? f->d_cap : nullptr;
    WindowQueries Q;
    Q.nq = nq; Q.desc = (const uint4*)(m->d_q + o_desc); Q.ref = (const float2*)(m->d_q + o_ref); Q.margin = (const float*)(m->d_q + o_m);
    Q.min_level = (const int*)(m->d_q + o_lo); Q.max_level = (const int*)(m->d_q + o_hi); Q.xr = xr_q ? (const float*)(m->d_q + o_xr) : nullptr;
    Q.fuse_gate = fuse_inv_sigma_sq ? 1 : 0;
    for (int l = 0; l < 16; ++l) Q.inv_sigma_sq[l] = (fuse_inv_sigma_sq && l < fuse_levels) ? fuse_inv_sigma_sq[l] : 0.0f;
    OVS_CUDA_CHECK(cudaEventRecord(m->ev[0], st));
    k_window_topk<<<(nq + 3) / 4, 128, 0, st>>>(F, Q, m->d_keys);
    OVS_LAUNCH_CHECK();
    OVS_CUDA_CHECK(cudaEventRecord(m->ev[1], st));
    OVS_CUDA_CHECK(cudaMemcpyAsync(m->h_keys, m->d_keys, N * kTopK * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    OVS_CUDA_CHECK(ovs::sync_stream(st));
    float ms = 0; cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]);
    m->last_kernel_us = ms * 1000.f;
    return OVS_OK;
}

// Greedy replay helper: the unclaimed/valid entries of a query's sorted key list.
struct Resolved {
    int r = 0;               // valid entries found in the list
    int dist[kTopK], idx[kTopK];
    bool exhausted = false;  // the list ended before 4 entries: nothing exists beyond it
    int lower_bound = OVS_MAX_HAMMING_DIST;  // every unlisted candidate has distance >= this
};

template <typename Valid>
Resolved resolve(const ovs_frame_index* f, const unsigned* keys, Valid valid) {
    Resolved R;
    for (int k = 0; k < kTopK; ++k) {
        if (keys[k] == 0xffffffffu) { R.exhausted = true; break; }
        const int idx = f->rank_to_idx[key_rank(keys[k])], d = key_dist(keys[k]);
        if (valid(idx, d)) { R.dist[R.r] = d; R.idx[R.r] = idx; ++R.r; }
    }
    R.lower_bound = R.exhausted ? OVS_MAX_HAMMING_DIST : key_dist(keys[kTopK - 1]);
    return R;
}

// Re-runs one query with a distance cap per keypoint (cap[idx]: candidate valid iff d < cap[idx]).
int requery(ovs_frame_index* f, const float* ref_xy, float margin, int min_level, int max_level, const float* xr_q,
            const uint8_t* qdesc, const std::vector<unsigned short>& cap_by_idx, unsigned* keys_out) {
    std::vector<unsigned short> cap((size_t)std::max(f->nranked, 1));
    for (int r = 0; r < f->nranked; ++r) cap[r] = cap_by_idx[f->rank_to_idx[r]];
    int rc = window_topk(f, 1, ref_xy, &margin, &min_level, &max_level, xr_q, qdesc, cap.data());
    if (rc != OVS_OK) return rc;
    memcpy(keys_out, f->m->h_keys, kTopK * sizeof(unsigned));
    return OVS_OK;
}

}  // namespace

namespace {

// data::assign_keypoints_to_grid on the host mirrors already stored in f (hx, hy): counting sort by cell (cx major, cy
// minor), index order kept inside a cell.  Fills rank_to_idx / idx_to_rank / nranked and returns the cell CSR.
std::vector<int> rank_keypoints(ovs_frame_index* f) {
    const ovs_grid& grid = f->grid;
    const int n = f->n, ncells = grid.num_grid_cols * grid.num_grid_rows;
    std::vector<int> cell(n, -1), start(ncells + 1, 0);
    for (int i = 0; i < n; ++i) {
        int cx, cy;
        if (cell_of(grid, f->hx[i], f->hy[i], &cx, &cy)) { cell[i] = cx * grid.num_grid_rows + cy; start[cell[i] + 1]++; }
    }
    for (int c = 0; c < ncells; ++c) start[c + 1] += start[c];
    f->nranked = start[ncells];
    f->rank_to_idx.assign(std::max(f->nranked, 1), 0); f->idx_to_rank.assign(std::max(n, 1), -1);
    std::vector<int> pos(start.begin(), start.end() - 1);
    for (int i = 0; i < n; ++i) if (cell[i] >= 0) { const int r = pos[cell[i]]++; f->rank_to_idx[r] = i; f->idx_to_rank[i] = r; }
    return start;
}

bool alloc_index_arrays(ovs_frame_index* f, size_t R, int ncells) {
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t o_desc = 0, o_x = o_desc + up(R * 32), o_y = o_x + up(R * 4), o_xr = o_y + up(R * 4), o_rank = o_xr + up(R * 4),
                 o_cell = o_rank + up(R * 4), o_cap = o_cell + up((size_t)(ncells + 1) * 4), o_oct = o_cap + up(R * 2), need = o_oct + up(R);
    ovs_matcher* m = f->m;
    int pick = -1;
    for (int i = 0; i < (int)m->index_pool.size(); ++i)
        if (m->index_pool[i].cap >= need && (pick < 0 || m->index_pool[i].cap < m->index_pool[pick].cap)) pick = i;
    if (pick >= 0) {
        f->buf = m->index_pool[pick];
        m->index_pool.erase(m->index_pool.begin() + pick);
    } else {
        const size_t cap = need + need / 4;
        if (cudaMalloc(&f->buf.base, cap) != cudaSuccess) { f->buf = ovs_index_buf(); return false; }
        f->buf.cap = cap;
    }
    uint8_t* b = f->buf.base;
    f->d_desc = reinterpret_cast<uint4*>(b + o_desc); f->d_x = reinterpret_cast<float*>(b + o_x); f->d_y = reinterpret_cast<float*>(b + o_y);
    f->d_xr = reinterpret_cast<float*>(b + o_xr); f->d_rank = reinterpret_cast<int*>(b + o_rank); f->d_cell_start = reinterpret_cast<int*>(b + o_cell);
    f->d_cap = reinterpret_cast<unsigned short*>(b + o_cap); f->d_oct = reinterpret_cast<signed char*>(b + o_oct);
    return true;
}

// rank-ordered SoA of the index straight from the extractor's device output (ovs_keypoint AoS + descriptors)
__global__ void __launch_bounds__(128) k_index_gather(int nranked, const int* __restrict__ rank_to_idx, const ovs_keypoint* __restrict__ kps,
                                                      const uint4* __restrict__ desc, const float* __restrict__ x_right,
                                                      float* __restrict__ ox, float* __restrict__ oy, float* __restrict__ oxr,
                                                      signed char* __restrict__ ooct, uint4* __restrict__ odesc) {
    const int r = blockIdx.x * 128 + threadIdx.x;
    if (r >= nranked) return;
    const int i = rank_to_idx[r];
    const ovs_keypoint k = kps[i];
    ox[r] = k.x; oy[r] = k.y; ooct[r] = (signed char)k.octave;
    oxr[r] = x_right ? x_right[i] : -1.0f;
    odesc[2 * (size_t)r] = desc[2 * (size_t)i];
    odesc[2 * (size_t)r + 1] = desc[2 * (size_t)i + 1];
}

}  // namespace

extern "C" int ovs_frame_index_create(ovs_matcher* m, int n, const float* x, const float* y, const int32_t* octave, const float* angle,
                                      const float* x_right, const uint8_t* desc, const ovs_grid* grid, ovs_frame_index** out) {
    OVS_REQUIRE(m && grid && out && n >= 0 && (n == 0 || (x && y && octave && desc)), OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(n < 65536, OVS_ERR_UNSUPPORTED, "more than 65535 keypoints");
    OVS_REQUIRE(grid->num_grid_cols > 0 && grid->num_grid_rows > 0 && grid->num_grid_cols * grid->num_grid_rows <= 1 << 20,
                OVS_ERR_INVALID_ARG, "bad grid");
    OVS_CUDA_CHECK(cudaSetDevice(m->device));
    ovs_frame_index* f = new (std::nothrow) ovs_frame_index();
    OVS_REQUIRE(f, OVS_ERR_CUDA, "out of host memory");
    f->m = m; f->n = n; f->grid = *grid; f->has_xr = x_right != nullptr;
    f->hx.assign(x, x + n); f->hy.assign(y, y + n); f->hoct.assign(octave, octave + n);
    if (angle) f->hangle.assign(angle, angle + n); else f->hangle.assign(n, 0.f);
    if (x_right) f->hxr.assign(x_right, x_right + n); else f->hxr.assign(n, -1.0f);
    const int ncells = grid->num_grid_cols * grid->num_grid_rows;
    const std::vector<int> start = rank_keypoints(f);
    const size_t R = (size_t)std::max(f->nranked, 1);
    std::vector<float> rx(R), ry(R), rxr(R); std::vector<signed char> roct(R); std::vector<uint8_t> rdesc(R * 32);
    for (int r = 0; r < f->nranked; ++r) {
        const int i = f->rank_to_idx[r];
        rx[r] = x[i]; ry[r] = y[i]; rxr[r] = f->hxr[i]; roct[r] = (signed char)octave[i];
        memcpy(&rdesc[(size_t)r * 32], desc + (size_t)i * 32, 32);
    }
    bool ok = alloc_index_arrays(f, R, ncells);
    if (ok) {
        cudaStream_t st = m->stream;
        ok = cudaMemcpyAsync(f->d_x, rx.data(), R * 4, cudaMemcpyHostToDevice, st) == cudaSuccess
             && cudaMemcpyAsync(f->d_y, ry.data(), R * 4, cudaMemcpyHostToDevice, st) == cudaSuccess
             && cudaMemcpyAsync(f->d_xr, rxr.data(), R * 4, cudaMemcpyHostToDevice, st) == cudaSuccess
             && cudaMemcpyAsync(f->d_oct, roct.data(), R, cudaMemcpyHostToDevice, st) == cudaSuccess
             && cudaMemcpyAsync(f->d_desc, rdesc.data(), R * 32, cudaMemcpyHostToDevice, st) == cudaSuccess
             && cudaMemcpyAsync(f->d_cell_start, start.data(), (size_t)(ncells + 1) * 4, cudaMemcpyHostToDevice, st) == cudaSuccess
             && ovs::sync_stream(st) == cudaSuccess;
    }
    if (!ok) {
        ovs::set_error("frame index allocation/upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        ovs_frame_index_destroy(f);
        return OVS_ERR_CUDA;
    }
    *out = f;
    return OVS_OK;
}

// The same index from the extractor's DEVICE output (ovs_extract_device): the descriptors never leave the GPU.  Only the
// keypoint records (28 B each; the host side of the matchers needs positions, octaves and angles anyway) come to the
// host, where assign_keypoints_to_grid is a counting sort; the rank-ordered arrays are then gathered on the device.
extern "C" int ovs_frame_index_create_device(ovs_matcher* m, int n, const ovs_keypoint* d_keypts, const uint8_t* d_desc,
                                             const float* d_x_right, const ovs_grid* grid, ovs_frame_index** out) {
    OVS_REQUIRE(m && grid && out && n >= 0 && (n == 0 || (d_keypts && d_desc)), OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(n < 65536, OVS_ERR_UNSUPPORTED, "more than 65535 keypoints");
    OVS_REQUIRE(grid->num_grid_cols > 0 && grid->num_grid_rows > 0 && grid->num_grid_cols * grid->num_grid_rows <= 1 << 20,
                OVS_ERR_INVALID_ARG, "bad grid");
    OVS_REQUIRE((reinterpret_cast<uintptr_t>(d_desc) & 15) == 0, OVS_ERR_INVALID_ARG, "descriptors must be 16-byte aligned");
    OVS_CUDA_CHECK(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    const size_t N = (size_t)std::max(n, 1);
    int rc = ovs::grow_host(&m->h_stage, &m->h_stage_cap, N * (sizeof(ovs_keypoint) + 4) + 64);
    if (rc != OVS_OK) return rc;
    ovs_keypoint* hk = reinterpret_cast<ovs_keypoint*>(m->h_stage);
    float* hxr = reinterpret_cast<float*>(m->h_stage + N * sizeof(ovs_keypoint));
    if (n) OVS_CUDA_CHECK(cudaMemcpyAsync(hk, d_keypts, (size_t)n * sizeof(ovs_keypoint), cudaMemcpyDeviceToHost, st));
    if (n && d_x_right) OVS_CUDA_CHECK(cudaMemcpyAsync(hxr, d_x_right, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    OVS_CUDA_CHECK(ovs::sync_stream(st));
    ovs_frame_index* f = new (std::nothrow) ovs_frame_index();
    OVS_REQUIRE(f, OVS_ERR_CUDA, "out of host memory");
    f->m = m; f->n = n; f->grid = *grid; f->has_xr = d_x_right != nullptr;
    f->hx.resize(n); f->hy.resize(n); f->hoct.resize(n); f->hangle.resize(n); f->hxr.assign(n, -1.0f);
    for (int i = 0; i < n; ++i) {
        f->hx[i] = hk[i].x; f->hy[i] = hk[i].y; f->hoct[i] = hk[i].octave; f->hangle[i] = hk[i].angle;
        if (d_x_right) f->hxr[i] = hxr[i];
    }
    const int ncells = grid->num_grid_cols * grid->num_grid_rows;
    const std::vector<int> start = rank_keypoints(f);
    const size_t R = (size_t)std::max(f->nranked, 1);
    bool ok = alloc_index_arrays(f, R, ncells);
    int* const d_rank = f->d_rank;
    if (ok) {
        ok = cudaMemcpyAsync(d_rank, f->rank_to_idx.data(), R * 4, cudaMemcpyHostToDevice, st) == cudaSuccess
             && cudaMemcpyAsync(f->d_cell_start, start.data(), (size_t)(ncells + 1) * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
        if (ok && f->nranked) {
            k_index_gather<<<(f->nranked + 127) / 128, 128, 0, st>>>(f->nranked, d_rank, d_keypts, reinterpret_cast<const uint4*>(d_desc), d_x_right,
                                                                     f->d_x, f->d_y, f->d_xr, f->d_oct, f->d_desc);
            ovs::count_launch();
            ok = cudaGetLastError() == cudaSuccess;
        }
        ok = ok && ovs::sync_stream(st) == cudaSuccess;
    }
    if (!ok) {
        ovs::set_error("frame index allocation/gather failed: %s", cudaGetErrorString(cudaGetLastError()));
        ovs_frame_index_destroy(f);
        return OVS_ERR_CUDA;
    }
    *out = f;
    return OVS_OK;
}

extern "C" void ovs_frame_index_destroy(ovs_frame_index* f) {
    if (!f) return;
    if (f->m) cudaSetDevice(f->m->device);
    if (f->buf.base) {
        // every call on the index returned with the matcher's stream drained, so the buffer is idle: keep it for the next frame
        if (f->m && f->m->index_pool.size() < 8) f->m->index_pool.push_back(f->buf);
        else cudaFree(f->buf.base);
    }
    delete f;
}

// get_keypoints_in_cell + nearest-descriptor search for nq queries: the 4 best candidates of each
// query in the reference's visiting order.  idx_out / dist_out [nq * 4], -1 / 256 where absent.
extern "C" int ovs_match_window_topk_host(ovs_frame_index* f, int nq, const float* ref_xy, const float* margin, const int32_t* min_level,
                                          const int32_t* max_level, const float* x_right_q, const uint8_t* qdesc,
                                          int32_t* idx_out, int32_t* dist_out) {
    OVS_REQUIRE(f && nq >= 0 && (nq == 0 || (ref_xy && margin && min_level && max_level && qdesc && idx_out && dist_out)), OVS_ERR_INVALID_ARG, "bad argument");
    if (nq == 0) return OVS_OK;
    OVS_CUDA_CHECK(cudaSetDevice(f->m->device));
    int rc = window_topk(f, nq, ref_xy, margin, min_level, max_level, x_right_q, qdesc, nullptr);
    if (rc != OVS_OK) return rc;
    for (size_t i = 0; i < (size_t)nq * kTopK; ++i) {
        const unsigned k = f->m->h_keys[i];
        idx_out[i] = k == 0xffffffffu ? -1 : f->rank_to_idx[key_rank(k)];
        dist_out[i] = key_dist(k);
    }
    return OVS_OK;
}

// match::projection::match_keyframes_mutually(keyfrm_1, keyfrm_2, matched_lms_in_keyfrm_2, Sim3s, margin): every usable
// landmark of keyframe 1 (index = its keypoint in keyframe 1) looks for its nearest keypoint of keyframe 2 inside the
// window around its reprojection, levels [pred - 1, pred], distance <= HAMMING_DIST_THR_HIGH -- independently of the
// other landmarks (no first-taker rule in this matcher) -- and vice versa; a pair is kept when both directions agree.
extern "C" int ovs_projection_match_keyframes_mutually_host(ovs_frame_index* f1, ovs_frame_index* f2, const float* scale_factors,
                                                            const uint8_t* usable_1, const float* reproj_1_in_2, const int32_t* pred_level_1_in_2,
                                                            const uint8_t* lm_desc_1, const uint8_t* usable_2, const float* reproj_2_in_1,
                                                            const int32_t* pred_level_2_in_1, const uint8_t* lm_desc_2, float margin,
                                                            int32_t* matched_idx_2_of_kp_1, int* num_matches) {
    OVS_REQUIRE(f1 && f2 && scale_factors && matched_idx_2_of_kp_1 && num_matches, OVS_ERR_INVALID_ARG, "bad argument");
    const int n1 = f1->n, n2 = f2->n;
    OVS_REQUIRE((n1 == 0 || (reproj_1_in_2 && pred_level_1_in_2 && lm_desc_1)) && (n2 == 0 || (reproj_2_in_1 && pred_level_2_in_1 && lm_desc_2)),
                OVS_ERR_INVALID_ARG, "null argument");
    *num_matches = 0;
    for (int i = 0; i < n1; ++i) matched_idx_2_of_kp_1[i] = -1;
    if (n1 == 0 || n2 == 0) return OVS_OK;
    // one direction: best keypoint of `dst` for every usable landmark of the other keyframe
    auto one_way = [&](ovs_frame_index* dst, int nq, const uint8_t* usable, const float* reproj, const int32_t* lvl, const uint8_t* desc,
                       std::vector<int>& best) -> int {
        OVS_CUDA_CHECK(cudaSetDevice(dst->m->device));
        std::vector<float> mg(nq); std::vector<int32_t> lo(nq), hi(nq);
        for (int q = 0; q < nq; ++q) {
            const int l = lvl[q];
            mg[q] = margin * scale_factors[l < 0 ? 0 : l];
            lo[q] = l - 1; hi[q] = l;
            if (usable && !usable[q]) { lo[q] = 1; hi[q] = 0; }   // empty level range: no candidates
        }
        const int rc = window_topk(dst, nq, reproj, mg.data(), lo.data(), hi.data(), nullptr, desc, nullptr);
        if (rc != OVS_OK) return rc;
        best.assign(nq, -1);
        for (int q = 0; q < nq; ++q) {
            if (usable && !usable[q]) continue;
            const unsigned k = dst->m->h_keys[(size_t)q * kTopK];
            if (k != 0xffffffffu && key_dist(k) <= OVS_HAMMING_DIST_THR_HIGH) best[q] = dst->rank_to_idx[key_rank(k)];
        }
        return OVS_OK;
    };
    std::vector<int> best_2_of_1, best_1_of_2;
    int rc = one_way(f2, n1, usable_1, reproj_1_in_2, pred_level_1_in_2, lm_desc_1, best_2_of_1);
    if (rc != OVS_OK) return rc;
    rc = one_way(f1, n2, usable_2, reproj_2_in_1, pred_level_2_in_1, lm_desc_2, best_1_of_2);
    if (rc != OVS_OK) return rc;
    for (int i1 = 0; i1 < n1; ++i1) {
        const int i2 = best_2_of_1[i1];
        if (i2 >= 0 && best_1_of_2[i2] == i1) { matched_idx_2_of_kp_1[i1] = i2; ++*num_matches; }
    }
    return OVS_OK;
}

// match::projection::match_frame_and_landmarks(frm, local_landmarks, margin)
extern "C" int ovs_projection_match_frame_and_landmarks_host(ovs_frame_index* f, const float* scale_factors, int num_scale_levels, int nlm, const uint8_t* lm_usable,
                                                             const float* reproj_xy, const float* x_right_in_tracking,
                                                             const int32_t* pred_scale_level, const uint8_t* lm_desc,
                                                             const uint8_t* kp_has_observed_lm, float margin, float lowe_ratio,
                                                             int32_t* matched_lm_of_kp, int* num_matches) {
    OVS_REQUIRE(f && scale_factors && num_matches && matched_lm_of_kp && nlm >= 0, OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(nlm == 0 || (reproj_xy && pred_scale_level && lm_desc), OVS_ERR_INVALID_ARG, "null argument");
    OVS_CUDA_CHECK(cudaSetDevice(f->m->device));
    const int n = f->n;
    for (int i = 0; i < n; ++i) matched_lm_of_kp[i] = -1;
    *num_matches = 0;
    if (nlm == 0 || n == 0) return OVS_OK;
    // queries = usable landmarks
    OVS_REQUIRE(num_scale_levels >= 1, OVS_ERR_INVALID_ARG, "bad scale table");
    std::vector<int> qlm; qlm.reserve(nlm);
    for (int l = 0; l < nlm; ++l) {
        if (lm_usable && !lm_usable[l]) continue;
        // the reference indexes scale_factors_.at(pred_scale_level): out of range throws there, is an error here
        OVS_REQUIRE(pred_scale_level[l] >= 0 && pred_scale_level[l] < num_scale_levels, OVS_ERR_INVALID_ARG,
                    "predicted scale level %d of landmark %d outside the scale table (%d levels)", pred_scale_level[l], l, num_scale_levels);
        qlm.push_back(l);
    }
    const int nq = (int)qlm.size();
    if (nq == 0) return OVS_OK;
    std::vector<float> ref(2 * (size_t)nq), mg(nq), xr(nq); std::vector<int> lo(nq), hi(nq); std::vector<uint8_t> qd(32 * (size_t)nq);
    for (int q = 0; q < nq; ++q) {
        const int l = qlm[q], lvl = pred_scale_level[l];
        ref[2 * q] = reproj_xy[2 * l]; ref[2 * q + 1] = reproj_xy[2 * l + 1];
        mg[q] = margin * scale_factors[lvl]; lo[q] = lvl - 1; hi[q] = lvl;
        xr[q] = x_right_in_tracking ? x_right_in_tracking[l] : -1.0f;
        memcpy(&qd[32 * (size_t)q], lm_desc + 32 * (size_t)l, 32);
    }
    std::vector<unsigned short> cap(std::max(n, 1), 0xffff);   // 0 = keypoint unavailable
    if (kp_has_observed_lm) for (int i = 0; i < n; ++i) if (kp_has_observed_lm[i]) cap[i] = 0;
    int rc;
    {
        std::vector<unsigned short> cap_rank((size_t)std::max(f->nranked, 1));
        for (int r = 0; r < f->nranked; ++r) cap_rank[r] = cap[f->rank_to_idx[r]];
        rc = window_topk(f, nq, ref.data(), mg.data(), lo.data(), hi.data(), f->has_xr ? xr.data() : nullptr, qd.data(), cap_rank.data());
        if (rc != OVS_OK) return rc;
    }
    std::vector<unsigned> keys(f->m->h_keys, f->m->h_keys + (size_t)nq * kTopK);
    int nm = 0;
    for (int q = 0; q < nq; ++q) {
        unsigned kq[kTopK];
        memcpy(kq, &keys[(size_t)q * kTopK], sizeof(kq));
        for (int attempt = 0; attempt < 2; ++attempt) {
            const Resolved R = resolve(f, kq, [&](int idx, int) { return cap[idx] != 0; });
            bool decided = true, accept = false;
            int best_idx = -1;
            if (R.r >= 2 || R.exhausted || attempt == 1) {
                if (R.r >= 1 && R.dist[0] <= OVS_HAMMING_DIST_THR_HIGH) {
                    best_idx = R.idx[0];
                    const int second = R.r >= 2 ? R.dist[1] : OVS_MAX_HAMMING_DIST;
                    const int best_level = f->hoct[R.idx[0]], second_level = R.r >= 2 ? f->hoct[R.idx[1]] : -1;
                    accept = !(best_level == second_level && (float)R.dist[0] > lowe_ratio * (float)second);
                }
            } else if (R.r == 1) {
                if (R.dist[0] <= OVS_HAMMING_DIST_THR_HIGH) {
                    best_idx = R.idx[0];
                    if ((float)R.dist[0] > lowe_ratio * (float)R.lower_bound) decided = false;   // the ratio test needs the true second best
                    else accept = true;
                }
            } else if (R.lower_bound <= OVS_HAMMING_DIST_THR_HIGH) decided = false;
            if (!decided) {
                rc = requery(f, &ref[2 * q], mg[q], lo[q], hi[q], f->has_xr ? &xr[q] : nullptr, &qd[32 * (size_t)q], cap, kq);
                if (rc != OVS_OK) return rc;
                continue;
            }
            if (accept) { matched_lm_of_kp[best_idx] = qlm[q]; cap[best_idx] = 0; ++nm; }
            break;
        }
    }
    *num_matches = nm;
    return OVS_OK;
}

// match::fuse (match/fuse.cc; ORB-SLAM2 ORBmatcher::Fuse): the matching core of fuse::replace_duplication /
// detect_duplication.  Every usable landmark (reprojected into the keyframe by the caller, who also does the depth-range and
// viewing-angle tests and predict_scale_level) searches the window margin * scale_factors[level] over the levels
// [level - 1, level]; candidates whose reprojection error exceeds the chi-square bound of their own octave are skipped;
// the nearest descriptor wins (first in visiting order on ties), accepted at <= HAMMING_DIST_THR_LOW.  There is no
// first-taker rule in this matcher, so the whole search is one kernel and no replay.  best_idx_of_lm[q] = keypoint or -1.
extern "C" int ovs_fuse_best_keypoints_host(ovs_frame_index* f, int nq, const uint8_t* usable, const float* reproj_xy, const float* reproj_x_right,
                                            const int32_t* pred_level, const uint8_t* lm_desc, const float* scale_factors,
                                            const float* inv_level_sigma_sq, int num_scale_levels, float margin,
                                            int32_t* best_idx_of_lm, int* num_matches) {
    OVS_REQUIRE(f && num_matches && nq >= 0 && (nq == 0 || (reproj_xy && pred_level && lm_desc && best_idx_of_lm)), OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(scale_factors && inv_level_sigma_sq && num_scale_levels >= 1 && num_scale_levels <= 16, OVS_ERR_INVALID_ARG, "bad scale tables");
    *num_matches = 0;
    for (int q = 0; q < nq; ++q) best_idx_of_lm[q] = -1;
    if (nq == 0 || f->n == 0) return OVS_OK;
    OVS_CUDA_CHECK(cudaSetDevice(f->m->device));
    for (int i = 0; i < f->n; ++i)
        OVS_REQUIRE(f->hoct[i] >= 0 && f->hoct[i] < num_scale_levels, OVS_ERR_INVALID_ARG, "octave %d of keypoint %d outside the scale tables", f->hoct[i], i);
    std::vector<int> ql; ql.reserve(nq);
    for (int q = 0; q < nq; ++q) {
        if (usable && !usable[q]) continue;
        OVS_REQUIRE(pred_level[q] < num_scale_levels, OVS_ERR_INVALID_ARG, "predicted level %d of landmark %d outside the scale tables", pred_level[q], q);
        ql.push_back(q);
    }
    const int nu = (int)ql.size();
    if (nu == 0) return OVS_OK;
    std::vector<float> ref(2 * (size_t)nu), mg(nu), xr(nu); std::vector<int> lo(nu), hi(nu); std::vector<uint8_t> qd(32 * (size_t)nu);
    for (int k = 0; k < nu; ++k) {
        const int q = ql[k], l = pred_level[q];
        ref[2 * k] = reproj_xy[2 * q]; ref[2 * k + 1] = reproj_xy[2 * q + 1];
        mg[k] = margin * scale_factors[l < 0 ? 0 : l]; lo[k] = l - 1; hi[k] = l;
        xr[k] = reproj_x_right ? reproj_x_right[q] : -1.0f;
        memcpy(&qd[32 * (size_t)k], lm_desc + 32 * (size_t)q, 32);
    }
    const int rc = window_topk(f, nu, ref.data(), mg.data(), lo.data(), hi.data(), reproj_x_right ? xr.data() : nullptr, qd.data(), nullptr,
                               inv_level_sigma_sq, num_scale_levels);
    if (rc != OVS_OK) return rc;
    int num = 0;
    for (int k = 0; k < nu; ++k) {
        const unsigned key = f->m->h_keys[(size_t)k * kTopK];
        if (key != 0xffffffffu && key_dist(key) <= OVS_HAMMING_DIST_THR_LOW) { best_idx_of_lm[ql[k]] = f->rank_to_idx[key_rank(key)]; ++num; }
    }
    *num_matches = num;
    return OVS_OK;
}

namespace {
// match::angle_checker<int>(30, 3)::get_invalid_matches over (delta_angle, tag) pairs
void angle_checker_invalid(const std::vector<float>& deltas, std::vector<uint8_t>& invalid) {
    const int H = 30, keepn = 3;
    const float inv_len = 1.0f / H;
    std::vector<int> bin(deltas.size()), count(H, 0), order(H);
    for (size_t i = 0; i < deltas.size(); ++i) {
        float d = deltas[i];
        if (d < 0.0) d += 360.0;
        if (360.0 <= d) d -= 360.0;
        bin[i] = (int)((unsigned)lrintf(d * inv_len) % (unsigned)H);
        count[bin[i]]++;
    }
    for (int b = 0; b < H; ++b) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return count[a] > count[b]; });
    std::vector<uint8_t> keep(H, 0);
    const int top = count[order[0]];
    for (int r = 0; r < keepn; ++r) {
        if (r > 0 && (float)count[order[r]] < 0.1f * (float)top) break;
        keep[order[r]] = 1;
    }
    invalid.resize(deltas.size());
    for (size_t i = 0; i < deltas.size(); ++i) invalid[i] = !keep[bin[i]];
}
}  // namespace

// The loop shared by projection::match_current_and_last_frames, match_frame_and_keyframe,
// match_by_Sim3_transform and each direction of match_keyframes_mutually (match/projection.cc): for
// every usable query (a landmark reprojected into this frame by the caller) search the window
// [min_level, max_level] x margin for the nearest descriptor among the keypoints that are still
// available, accept it when the distance is <= hamm_dist_thr, mark the keypoint taken; finally drop
// the matches that disagree with the dominant rotation (angle_checker) when check_orientation.
extern "C" int ovs_projection_match_best_host(ovs_frame_index* f, int nq, const uint8_t* usable, const float* ref_xy, const float* ref_x_right,
                                              const float* margin, const int32_t* min_level, const int32_t* max_level, const float* q_angle,
                                              const uint8_t* q_desc, const uint8_t* kp_unavailable, unsigned hamm_dist_thr, int check_orientation,
                                              int32_t* matched_query_of_kp, int* num_matches) {
    OVS_REQUIRE(f && num_matches && matched_query_of_kp && nq >= 0, OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(nq == 0 || (ref_xy && margin && min_level && max_level && q_desc && (!check_orientation || q_angle)), OVS_ERR_INVALID_ARG, "null argument");
    OVS_REQUIRE(hamm_dist_thr <= OVS_MAX_HAMMING_DIST, OVS_ERR_INVALID_ARG, "hamm_dist_thr out of range");
    OVS_CUDA_CHECK(cudaSetDevice(f->m->device));
    const int n = f->n;
    for (int i = 0; i < n; ++i) matched_query_of_kp[i] = -1;
    *num_matches = 0;
    std::vector<int> ql; ql.reserve(nq);
    for (int i = 0; i < nq; ++i) if (!usable || usable[i]) ql.push_back(i);
    const int nu = (int)ql.size();
    if (nu == 0 || n == 0) return OVS_OK;
    std::vector<float> ref(2 * (size_t)nu), mg(nu), xr(nu); std::vector<int> lo(nu), hi(nu); std::vector<uint8_t> qd(32 * (size_t)nu);
    for (int q = 0; q < nu; ++q) {
        const int i = ql[q];
        ref[2 * q] = ref_xy[2 * i]; ref[2 * q + 1] = ref_xy[2 * i + 1];
        mg[q] = margin[i]; lo[q] = min_level[i]; hi[q] = max_level[i];
        xr[q] = ref_x_right ? ref_x_right[i] : -1.0f;
        memcpy(&qd[32 * (size_t)q], q_desc + 32 * (size_t)i, 32);
    }
    const bool use_xr = f->has_xr && ref_x_right != nullptr;
    std::vector<unsigned short> cap(std::max(n, 1), 0xffff);
    if (kp_unavailable) for (int i = 0; i < n; ++i) if (kp_unavailable[i]) cap[i] = 0;
    int rc;
    {
        std::vector<unsigned short> cap_rank((size_t)std::max(f->nranked, 1));
        for (int r = 0; r < f->nranked; ++r) cap_rank[r] = cap[f->rank_to_idx[r]];
        // without reprojected x_right the x_right test of the reference is not part of this matcher
        const bool saved = f->has_xr;
        f->has_xr = use_xr;
        rc = window_topk(f, nu, ref.data(), mg.data(), lo.data(), hi.data(), use_xr ? xr.data() : nullptr, qd.data(), cap_rank.data());
        f->has_xr = saved;
        if (rc != OVS_OK) return rc;
    }
    std::vector<unsigned> keys(f->m->h_keys, f->m->h_keys + (size_t)nu * kTopK);
    int nm = 0;
    std::vector<float> deltas; std::vector<int> delta_kp;
    for (int q = 0; q < nu; ++q) {
        unsigned kq[kTopK];
        memcpy(kq, &keys[(size_t)q * kTopK], sizeof(kq));
        for (int attempt = 0; attempt < 2; ++attempt) {
            const Resolved R = resolve(f, kq, [&](int idx, int) { return cap[idx] != 0; });
            if (R.r == 0 && !R.exhausted && attempt == 0 && R.lower_bound <= (int)hamm_dist_thr) {
                const bool saved = f->has_xr;
                f->has_xr = use_xr;
                rc = requery(f, &ref[2 * q], mg[q], lo[q], hi[q], use_xr ? &xr[q] : nullptr, &qd[32 * (size_t)q], cap, kq);
                f->has_xr = saved;
                if (rc != OVS_OK) return rc;
                continue;
            }
            if (R.r >= 1 && !((int)hamm_dist_thr < R.dist[0])) {
                const int best_idx = R.idx[0];
                matched_query_of_kp[best_idx] = ql[q]; cap[best_idx] = 0; ++nm;
                if (check_orientation) { deltas.push_back(q_angle[ql[q]] - f->hangle[best_idx]); delta_kp.push_back(best_idx); }
            }
            break;
        }
    }
    if (check_orientation && !deltas.empty()) {
        std::vector<uint8_t> invalid;
        angle_checker_invalid(deltas, invalid);
        for (size_t k = 0; k < deltas.size(); ++k) if (invalid[k]) { matched_query_of_kp[delta_kp[k]] = -1; --nm; }
    }
    *num_matches = nm;
    return OVS_OK;
}

// match::projection::match_current_and_last_frames(curr_frm, last_frm, margin)
extern "C" int ovs_projection_match_current_and_last_host(ovs_frame_index* curr, const float* scale_factors, int num_scale_levels, int n_last,
                                                          const uint8_t* last_usable, const float* reproj_xy, const float* reproj_x_right,
                                                          const int32_t* last_scale_level, const float* last_angle, const uint8_t* lm_desc,
                                                          const uint8_t* kp_has_observed_lm, float margin, int assume_forward, int assume_backward,
                                                          int check_orientation, int32_t* matched_last_of_kp, int* num_matches) {
    OVS_REQUIRE(curr && scale_factors && num_matches && matched_last_of_kp && n_last >= 0, OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(n_last == 0 || (last_usable && reproj_xy && last_scale_level && lm_desc), OVS_ERR_INVALID_ARG, "null argument");
    std::vector<float> mg(std::max(n_last, 1)); std::vector<int32_t> lo(std::max(n_last, 1)), hi(std::max(n_last, 1));
    for (int i = 0; i < n_last; ++i) {
        const int lvl = last_scale_level[i];
        OVS_REQUIRE((last_usable && !last_usable[i]) || (lvl >= 0 && lvl < num_scale_levels), OVS_ERR_INVALID_ARG,
                    "scale level %d of last-frame keypoint %d outside the scale table (%d levels)", lvl, i, num_scale_levels);
        mg[i] = (lvl >= 0 && lvl < num_scale_levels) ? margin * scale_factors[lvl] : 0.0f;
        if (assume_forward) { lo[i] = lvl; hi[i] = num_scale_levels - 1; }
        else if (assume_backward) { lo[i] = 0; hi[i] = lvl; }
        else { lo[i] = lvl - 1; hi[i] = lvl + 1; }
    }
    // the reference applies the x_right test whenever the current keypoint has one; a NULL reproj_x_right means -1
    std::vector<float> xr;
    if (!reproj_x_right && curr->has_xr) { xr.assign(std::max(n_last, 1), -1.0f); reproj_x_right = xr.data(); }
    return ovs_projection_match_best_host(curr, n_last, last_usable, reproj_xy, reproj_x_right, mg.data(), lo.data(), hi.data(), last_angle, lm_desc,
                                          kp_has_observed_lm, OVS_HAMMING_DIST_THR_HIGH, check_orientation, matched_last_of_kp, num_matches);
}

// match::area::match_in_consistent_area(frm_1, frm_2, prev_matched_pts, matched_indices_2_in_frm_1, margin)
extern "C" int ovs_area_match_in_consistent_area_host(ovs_frame_index* f2, int n1, const int32_t* octave_1, const float* angle_1, const uint8_t* desc_1,
                                                      float* prev_matched_xy, int32_t* matched_idx_2_in_1, int margin, float lowe_ratio,
                                                      int check_orientation, int* num_matches) {
    OVS_REQUIRE(f2 && num_matches && n1 >= 0 && (n1 == 0 || (octave_1 && desc_1 && prev_matched_xy && matched_idx_2_in_1)), OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(!check_orientation || angle_1 || n1 == 0, OVS_ERR_INVALID_ARG, "angles required for the orientation check");
    OVS_CUDA_CHECK(cudaSetDevice(f2->m->device));
    ovs_frame_index* f = f2;
    const int n2 = f->n;
    for (int i = 0; i < n1; ++i) matched_idx_2_in_1[i] = -1;
    *num_matches = 0;
    std::vector<int> q1; q1.reserve(n1);
    for (int i = 0; i < n1; ++i) if (!(0 < octave_1[i])) q1.push_back(i);   // level-0 keypoints only
    const int nq = (int)q1.size();
    if (nq == 0 || n2 == 0) return OVS_OK;
    std::vector<float> ref(2 * (size_t)nq), mg(nq, (float)margin); std::vector<int> lo(nq), hi(nq); std::vector<uint8_t> qd(32 * (size_t)nq);
    for (int q = 0; q < nq; ++q) {
        const int i = q1[q];
        ref[2 * q] = prev_matched_xy[2 * i]; ref[2 * q + 1] = prev_matched_xy[2 * i + 1];
        lo[q] = octave_1[i]; hi[q] = octave_1[i];
        memcpy(&qd[32 * (size_t)q], desc_1 + 32 * (size_t)i, 32);
    }
    int rc = window_topk(f, nq, ref.data(), mg.data(), lo.data(), hi.data(), nullptr, qd.data(), nullptr);
    if (rc != OVS_OK) return rc;
    std::vector<unsigned> keys(f->m->h_keys, f->m->h_keys + (size_t)nq * kTopK);
    // matched_dists_in_frm_2 doubles as the per-keypoint distance cap (candidate valid iff d < cap)
    std::vector<unsigned short> cap(std::max(n2, 1), (unsigned short)OVS_MAX_HAMMING_DIST);
    std::vector<int> matched_idx_1_in_2(std::max(n2, 1), -1);
    std::vector<float> deltas; std::vector<int> delta_idx;
    int nm = 0;
    for (int q = 0; q < nq; ++q) {
        const int idx_1 = q1[q];
        unsigned kq[kTopK];
        memcpy(kq, &keys[(size_t)q * kTopK], sizeof(kq));
        for (int attempt = 0; attempt < 2; ++attempt) {
            const Resolved R = resolve(f, kq, [&](int idx, int d) { return d < (int)cap[idx]; });
            bool decided = true, accept = false;
            if (R.r >= 2 || R.exhausted || attempt == 1) {
                if (R.r >= 1 && !(OVS_HAMMING_DIST_THR_LOW < R.dist[0])) {
                    const int second = R.r >= 2 ? R.dist[1] : OVS_MAX_HAMMING_DIST;
                    accept = !((float)second * lowe_ratio < (float)R.dist[0]);
                }
            } else if (R.r == 1) {
                if (!(OVS_HAMMING_DIST_THR_LOW < R.dist[0])) {
                    if ((float)R.lower_bound * lowe_ratio < (float)R.dist[0]) decided = false;
                    else accept = true;
                }
            } else if (R.lower_bound <= OVS_HAMMING_DIST_THR_LOW) decided = false;
            if (!decided) {
                rc = requery(f, &ref[2 * q], mg[q], lo[q], hi[q], nullptr, &qd[32 * (size_t)q], cap, kq);
                if (rc != OVS_OK) return rc;
                continue;
            }
            if (accept) {
                const int best_idx_2 = R.idx[0];
                const int prev_idx_1 = matched_idx_1_in_2[best_idx_2];
                if (0 <= prev_idx_1) { matched_idx_2_in_1[prev_idx_1] = -1; --nm; }
                matched_idx_2_in_1[idx_1] = best_idx_2;
                matched_idx_1_in_2[best_idx_2] = idx_1;
                cap[best_idx_2] = (unsigned short)R.dist[0];
                ++nm;
                if (check_orientation) { deltas.push_back(angle_1[idx_1] - f->hangle[best_idx_2]); delta_idx.push_back(idx_1); }
            }
            break;
        }
    }
    if (check_orientation && !deltas.empty()) {
        std::vector<uint8_t> invalid;
        angle_checker_invalid(deltas, invalid);
        for (size_t k = 0; k < deltas.size(); ++k)
            if (invalid[k] && 0 <= matched_idx_2_in_1[delta_idx[k]]) { matched_idx_2_in_1[delta_idx[k]] = -1; --nm; }
    }
    for (int i = 0; i < n1; ++i)
        if (0 <= matched_idx_2_in_1[i]) { prev_matched_xy[2 * i] = f->hx[matched_idx_2_in_1[i]]; prev_matched_xy[2 * i + 1] = f->hy[matched_idx_2_in_1[i]]; }
    *num_matches = nm;
    return OVS_OK;
}

// match::stereo(left_pyr, right_pyr, ...).compute(stereo_x_right, depths); the pyramids are the
// device-resident image_pyramid_ of the two extractors (after their extract() of this stereo pair).
extern "C" int ovs_stereo_compute_host(ovs_matcher* m, const ovs_extractor* left, const ovs_extractor* right,
                                       int n_left, const float* lx, const float* ly, const int32_t* loct, const uint8_t* ldesc,
                                       int n_right, const float* rx, const float* ry, const int32_t* roct, const uint8_t* rdesc,
                                       float focal_x_baseline, float true_baseline, float* stereo_x_right, float* depths, int* num_matched) {
    OVS_REQUIRE(m && left && right && n_left >= 0 && n_right >= 0, OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(n_left == 0 || (lx && ly && loct && ldesc && stereo_x_right && depths), OVS_ERR_INVALID_ARG, "null argument");
    OVS_REQUIRE(n_right == 0 || (rx && ry && roct && rdesc), OVS_ERR_INVALID_ARG, "null argument");
    OVS_REQUIRE(n_right < 65535, OVS_ERR_UNSUPPORTED, "more than 65534 right keypoints");
    OVS_REQUIRE(true_baseline > 0, OVS_ERR_INVALID_ARG, "true_baseline must be positive");
    for (int i = 0; i < n_left; ++i) { stereo_x_right[i] = -1.0f; depths[i] = -1.0f; }
    if (num_matched) *num_matched = 0;
    if (n_left == 0 || n_right == 0) return OVS_OK;
    OVS_CUDA_CHECK(cudaSetDevice(m->device));
    StereoArgs A;
    memset(&A, 0, sizeof(A));
    float sf[16], isf[16];
    int rc = ovs_extractor_scale_factors(left, sf, isf, nullptr, nullptr);
    if (rc != OVS_OK) return rc;
    int L = 0;
    for (; L < 16; ++L) {
        const uint8_t *pl, *pr; size_t pitl, pitr; int wl, hl, wr, hr;
        if (ovs_extractor_pyramid_level(left, L, &pl, &pitl, &wl, &hl) != OVS_OK) break;
        if (ovs_extractor_pyramid_level(right, L, &pr, &pitr, &wr, &hr) != OVS_OK) break;
        OVS_REQUIRE(wl == wr && hl == hr && pitl == pitr, OVS_ERR_INVALID_ARG, "left/right pyramids differ at level %d", L);
        A.lpyr[L] = pl; A.rpyr[L] = pr; A.pw[L] = wl; A.ph[L] = hl; A.pitch[L] = (int)pitl; A.scale[L] = sf[L]; A.inv_scale[L] = isf[L];
    }
    OVS_REQUIRE(L > 0, OVS_ERR_INVALID_ARG, "extractors hold no pyramid (call extract first)");
    for (int i = 0; i < n_left; ++i) OVS_REQUIRE(loct[i] >= 0 && loct[i] < L, OVS_ERR_INVALID_ARG, "left octave out of range");
    for (int i = 0; i < n_right; ++i) OVS_REQUIRE(roct[i] >= 0 && roct[i] < L, OVS_ERR_INVALID_ARG, "right octave out of range");
    const size_t NL = (size_t)n_left, NR = (size_t)n_right;
    size_t off = 0;
    auto carve = [&](size_t b) { const size_t o = off; off += (b + 15) / 16 * 16; return o; };
    const size_t o_ld = carve(32 * NL), o_rd = carve(32 * NR), o_lx = carve(4 * NL), o_ly = carve(4 * NL), o_lo = carve(4 * NL),
                 o_rx = carve(4 * NR), o_ry = carve(4 * NR), o_ro = carve(4 * NR);
    const size_t in_bytes = off;
    const size_t o_best = carve(4 * NL), o_xr = carve(4 * NL), o_dp = carve(4 * NL), o_co = carve(4 * NL);
    if ((rc = ovs::grow_host(&m->h_stage, &m->h_stage_cap, off)) != OVS_OK) return rc;
    if ((rc = ovs::grow_dev(&m->d_t, &m->d_t_cap, off)) != OVS_OK) return rc;
    uint8_t* hs = m->h_stage; uint8_t* ds = m->d_t;
    memcpy(hs + o_ld, ldesc, 32 * NL); memcpy(hs + o_rd, rdesc, 32 * NR);
    memcpy(hs + o_lx, lx, 4 * NL); memcpy(hs + o_ly, ly, 4 * NL); memcpy(hs + o_lo, loct, 4 * NL);
    memcpy(hs + o_rx, rx, 4 * NR); memcpy(hs + o_ry, ry, 4 * NR); memcpy(hs + o_ro, roct, 4 * NR);
    cudaStream_t st = m->stream;
    OVS_CUDA_CHECK(cudaMemcpyAsync(ds, hs, in_bytes, cudaMemcpyHostToDevice, st));
    A.n_left = n_left; A.n_right = n_right;
    A.ldesc = (const uint4*)(ds + o_ld); A.rdesc = (const uint4*)(ds + o_rd);
    A.lx = (const float*)(ds + o_lx); A.ly = (const float*)(ds + o_ly); A.loct = (const int*)(ds + o_lo);
    A.rx = (const float*)(ds + o_rx); A.ry = (const float*)(ds + o_ry); A.roct = (const int*)(ds + o_ro);
    A.min_disp = 0.0f; A.max_disp = focal_x_baseline / true_baseline;
    A.hamm_thr = (OVS_HAMMING_DIST_THR_HIGH + OVS_HAMMING_DIST_THR_LOW) / 2;
    A.rows0 = A.ph[0]; A.focal_x_baseline = focal_x_baseline;
    OVS_CUDA_CHECK(cudaEventRecord(m->ev[0], st));
    k_stereo_match<<<(n_left + 127) / 128, 128, 0, st>>>(A, (unsigned*)(ds + o_best));
    OVS_LAUNCH_CHECK();
    k_stereo_subpixel<<<(n_left + 3) / 4, 128, 0, st>>>(A, (const unsigned*)(ds + o_best), (float*)(ds + o_xr), (float*)(ds + o_dp), (unsigned*)(ds + o_co));
    OVS_LAUNCH_CHECK();
    OVS_CUDA_CHECK(cudaEventRecord(m->ev[1], st));
    OVS_CUDA_CHECK(cudaMemcpyAsync(hs + o_xr, ds + o_xr, (o_co + 4 * NL) - o_xr, cudaMemcpyDeviceToHost, st));
    OVS_CUDA_CHECK(ovs::sync_stream(st));
    float ms = 0; cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]);
    m->last_kernel_us = ms * 1000.f;
    const float* hxr = (const float*)(hs + o_xr); const float* hdp = (const float*)(hs + o_dp); const unsigned* hco = (const unsigned*)(hs + o_co);
    // median test on the SAD of the accepted matches (sorted (correlation, idx_left) pairs)
    std::vector<std::pair<unsigned, int>> corr;
    for (int i = 0; i < n_left; ++i) {
        if (hco[i] == 0xffffffffu) continue;
        stereo_x_right[i] = hxr[i]; depths[i] = hdp[i];
        corr.emplace_back(hco[i], i);
    }
    if (!corr.empty()) {
        std::sort(corr.begin(), corr.end());
        const float median = (float)corr[corr.size() / 2].first;
        const float thr = 2.0f * median;
        for (int k = (int)corr.size() - 1; k >= 0; --k) {
            if ((float)corr[k].first < thr) break;
            stereo_x_right[corr[k].second] = -1; depths[corr[k].second] = -1;
        }
    }
    if (num_matched) *num_matched = (int)corr.size();
    return OVS_OK;
}

// ====================================================================================================================
// match::robust::match_for_triangulation(keyfrm_1, keyfrm_2, E_12, matched_idx_pairs) (match/robust.cc): BoW-node-guided
// candidate pairs + epipole and epipolar-plane tests.  The BoW feature vectors are inputs (bow_node_k[i] = vocabulary
// node of keypoint i); keyframe-2 keypoints are laid out node by node on the device, each eligible keyframe-1 keypoint is
// one query (one warp) over its node's segment.  The kernel returns the 8 best admissible candidates per query in the
// order the sequential loop prefers them (smallest distance; among equal distances the LAST one); the host replays the
// "a keyframe-2 keypoint is given to its first taker" rule and the orientation histogram.
namespace {

constexpr int kTriK = 8;

struct TriArgs {
    int nq;
    const uint4* qdesc;        // [nq][2]
    const double* qbearing;    // [nq][3]
    const float* qscale;       // [nq] scale factor of the keypoint's octave
    const unsigned char* qstereo;
    const int2* qseg;          // [nq] candidate rank range [begin, end)
    const uint4* tdesc;        // rank order (node major, index minor)
    const double* tbearing;
    const unsigned char* tstereo;
    double E[9];
    double epipole[3];
    // re-query of one keypoint whose candidate list was exhausted by earlier takers: queries [q0, nq) only, keyframe-2
    // keypoints with taken[rank] != 0 are skipped, results go to slot q + out_shift
    int q0, out_shift;
    const unsigned char* taken;   // [R] or null
};

__device__ __forceinline__ bool epipolar_inlier(const double* E, const double b1x, const double b1y, const double b1z, const double b2x,
                                                const double b2y, const double b2z, const float scale) {
    const double ex = E[0] * b2x + E[1] * b2y + E[2] * b2z;
    const double ey = E[3] * b2x + E[4] * b2y + E[5] * b2z;
    const double ez = E[6] * b2x + E[7] * b2y + E[8] * b2z;
    const double norm = sqrt(ex * ex + ey * ey + ez * ez);
    const double cos_residual = (ex * b1x + ey * b1y + ez * b1z) / norm;
    const double residual_rad = 3.14159265358979323846 / 2.0 - fabs(acos(cos_residual));
    const double thr = 0.2 * 3.14159265358979323846 / 180.0;
    return residual_rad < thr * (double)scale;
}

// one warp per query; keys = distance << 16 | (0xffff - rank): ascending order = the sequential loop's preference
__global__ void __launch_bounds__(128) k_triangulation_topk(TriArgs A, unsigned* __restrict__ keys_out) {
    const int q = A.q0 + blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= A.nq) return;
    const uint4 qa = A.qdesc[2 * (size_t)q], qb = A.qdesc[2 * (size_t)q + 1];
    const double b1x = A.qbearing[3 * (size_t)q], b1y = A.qbearing[3 * (size_t)q + 1], b1z = A.qbearing[3 * (size_t)q + 2];
    const float scale = A.qscale[q];
    const bool stereo_1 = A.qstereo[q] != 0;
    const int2 seg = A.qseg[q];
    unsigned extra = 0xffffffffu;   // lane r < 8 carries entry r of the running top-8 from one chunk of candidates to the next
    for (int c0 = seg.x; c0 < seg.y; c0 += 32 * kTriK) {
        // each lane evaluates up to kTriK candidates of this chunk
        unsigned mine[kTriK];
#pragma unroll
        for (int k = 0; k < kTriK; ++k) {
            mine[k] = 0xffffffffu;
            const int c = c0 + k * 32 + lane;
            if (c < seg.y && !(A.taken && A.taken[c])) {
                const uint4 ta = A.tdesc[2 * (size_t)c], tb = A.tdesc[2 * (size_t)c + 1];
                const int d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w)
                              + __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
                if (d <= OVS_HAMMING_DIST_THR_LOW) {
                    const double b2x = A.tbearing[3 * (size_t)c], b2y = A.tbearing[3 * (size_t)c + 1], b2z = A.tbearing[3 * (size_t)c + 2];
                    bool ok = true;
                    if (!stereo_1 && !A.tstereo[c]) {
                        const double cos_dist = A.epipole[0] * b2x + A.epipole[1] * b2y + A.epipole[2] * b2z;
                        if (0.998 < cos_dist) ok = false;
                    }
                    if (ok && epipolar_inlier(A.E, b1x, b1y, b1z, b2x, b2y, b2z, scale)) mine[k] = ((unsigned)d << 16) | (0xffffu - (unsigned)c);
                }
            }
        }
        // the 8 smallest keys of {this chunk} U {running top-8}; keys are unique (they carry the rank)
        unsigned next_extra = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < kTriK; ++r) {
            unsigned lmin = extra;
#pragma unroll
            for (int k = 0; k < kTriK; ++k) lmin = min(lmin, mine[k]);
            const unsigned wmin = __reduce_min_sync(0xffffffffu, lmin);
            if (wmin != 0xffffffffu) {
                if (extra == wmin) extra = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < kTriK; ++k) if (mine[k] == wmin) mine[k] = 0xffffffffu;
            }
            if (lane == r) next_extra = wmin;
        }
        extra = next_extra;
    }
    if (lane < kTriK) keys_out[(size_t)(q + A.out_shift) * kTriK + lane] = extra;
}

}  // namespace

extern "C" int ovs_robust_match_for_triangulation_host(ovs_matcher* m, int n1, const uint8_t* desc_1, const double* bearing_1, const int32_t* octave_1,
                                                       const float* angle_1, const uint8_t* has_lm_1, const uint8_t* is_stereo_1,
                                                       const int32_t* bow_node_1, int n2, const uint8_t* desc_2, const double* bearing_2,
                                                       const float* angle_2, const uint8_t* has_lm_2, const uint8_t* is_stereo_2,
                                                       const int32_t* bow_node_2, const double* E_12, const double* epipole_in_2,
                                                       const float* scale_factors_1, int num_scale_levels, int check_orientation,
                                                       int32_t* matched_idx_2_of_1, int* num_matches) {
    OVS_REQUIRE(m && E_12 && epipole_in_2 && scale_factors_1 && matched_idx_2_of_1 && num_matches && n1 >= 0 && n2 >= 0, OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(n1 == 0 || (desc_1 && bearing_1 && octave_1 && angle_1 && has_lm_1 && bow_node_1), OVS_ERR_INVALID_ARG, "null keyframe-1 array");
    OVS_REQUIRE(n2 == 0 || (desc_2 && bearing_2 && angle_2 && has_lm_2 && bow_node_2), OVS_ERR_INVALID_ARG, "null keyframe-2 array");
    OVS_REQUIRE(n2 < 65536, OVS_ERR_UNSUPPORTED, "more than 65535 keypoints");
    *num_matches = 0;
    for (int i = 0; i < n1; ++i) matched_idx_2_of_1[i] = -1;
    if (n1 == 0 || n2 == 0) return OVS_OK;
    for (int i = 0; i < n1; ++i)
        OVS_REQUIRE(octave_1[i] >= 0 && octave_1[i] < num_scale_levels, OVS_ERR_INVALID_ARG, "octave %d of keypoint %d outside the scale table", octave_1[i], i);
    OVS_CUDA_CHECK(cudaSetDevice(m->device));
    // keyframe 2: eligible keypoints (no landmark, has a node) node by node, index order inside a node
    std::vector<int> rank2;
    for (int i = 0; i < n2; ++i) if (!has_lm_2[i] && bow_node_2[i] >= 0) rank2.push_back(i);
    std::stable_sort(rank2.begin(), rank2.end(), [&](int a, int b) { return bow_node_2[a] < bow_node_2[b]; });
    // keyframe 1: queries in the order the reference visits them (node ascending, index ascending)
    std::vector<int> q1;
    for (int i = 0; i < n1; ++i) if (!has_lm_1[i] && bow_node_1[i] >= 0) q1.push_back(i);
    std::stable_sort(q1.begin(), q1.end(), [&](int a, int b) { return bow_node_1[a] < bow_node_1[b]; });
    const int R = (int)rank2.size(), Q = (int)q1.size();
    if (R == 0 || Q == 0) return OVS_OK;
    std::vector<int2> seg(Q);
    {
        size_t lo = 0;
        for (int k = 0; k < Q; ++k) {
            const int node = bow_node_1[q1[k]];
            while (lo < rank2.size() && bow_node_2[rank2[lo]] < node) ++lo;
            size_t hi = lo;
            while (hi < rank2.size() && bow_node_2[rank2[hi]] == node) ++hi;
            seg[k] = make_int2((int)lo, (int)hi);
        }
    }
    // staging layout (host pinned / device), widest alignment first:
    // [qdesc 32Q][tdesc 32R][qbearing 24Q][tbearing 24R][qseg 8Q][qscale 4Q][qstereo Q][tstereo R][taken R]
    const size_t o_qd = 0, o_td = o_qd + 32 * (size_t)Q, o_qb = o_td + 32 * (size_t)R, o_tb = o_qb + 24 * (size_t)Q, o_sg = o_tb + 24 * (size_t)R,
                 o_qs = o_sg + 8 * (size_t)Q, o_q8 = o_qs + 4 * (size_t)Q, o_t8 = o_q8 + (size_t)Q, in_total = ((o_t8 + (size_t)R + 15) / 16) * 16,
                 o_tk = in_total, total = ((o_tk + (size_t)R + 15) / 16) * 16;
    int rc;
    if ((rc = ovs::grow_host(&m->h_stage, &m->h_stage_cap, total)) != OVS_OK) return rc;
    if ((rc = ovs::grow_dev(&m->d_q, &m->d_q_cap, total)) != OVS_OK) return rc;
    if ((rc = ovs::grow_dev(&m->d_keys, &m->d_keys_cap, (size_t)(Q + 1) * kTriK)) != OVS_OK) return rc;
    if ((rc = ovs::grow_host(&m->h_keys, &m->h_keys_cap, (size_t)(Q + 1) * kTriK)) != OVS_OK) return rc;
    uint8_t* hs = m->h_stage;
    for (int k = 0; k < Q; ++k) {
        const int i = q1[k];
        memcpy(hs + o_qd + 32 * (size_t)k, desc_1 + 32 * (size_t)i, 32);
        memcpy(hs + o_qb + 24 * (size_t)k, bearing_1 + 3 * (size_t)i, 24);
        reinterpret_cast<float*>(hs + o_qs)[k] = scale_factors_1[octave_1[i]];
        reinterpret_cast<int2*>(hs + o_sg)[k] = seg[k];
        hs[o_q8 + k] = is_stereo_1 ? is_stereo_1[i] : 0;
    }
    for (int r = 0; r < R; ++r) {
        const int i = rank2[r];
        memcpy(hs + o_td + 32 * (size_t)r, desc_2 + 32 * (size_t)i, 32);
        memcpy(hs + o_tb + 24 * (size_t)r, bearing_2 + 3 * (size_t)i, 24);
        hs[o_t8 + r] = is_stereo_2 ? is_stereo_2[i] : 0;
    }
    cudaStream_t st = m->stream;
    uint8_t* ds = m->d_q;
    OVS_CUDA_CHECK(cudaMemcpyAsync(ds, hs, in_total, cudaMemcpyHostToDevice, st));
    TriArgs A{};
    A.nq = Q; A.qdesc = reinterpret_cast<const uint4*>(ds + o_qd); A.tdesc = reinterpret_cast<const uint4*>(ds + o_td);
    A.qbearing = reinterpret_cast<const double*>(ds + o_qb); A.tbearing = reinterpret_cast<const double*>(ds + o_tb);
    A.qscale = reinterpret_cast<const float*>(ds + o_qs); A.qseg = reinterpret_cast<const int2*>(ds + o_sg);
    A.qstereo = ds + o_q8; A.tstereo = ds + o_t8;
    for (int k = 0; k < 9; ++k) A.E[k] = E_12[k];
    for (int k = 0; k < 3; ++k) A.epipole[k] = epipole_in_2[k];
    OVS_CUDA_CHECK(cudaEventRecord(m->ev[0], st));
    k_triangulation_topk<<<(Q + 3) / 4, 128, 0, st>>>(A, m->d_keys);
    OVS_LAUNCH_CHECK();
    OVS_CUDA_CHECK(cudaEventRecord(m->ev[1], st));
    OVS_CUDA_CHECK(cudaMemcpyAsync(m->h_keys, m->d_keys, (size_t)Q * kTriK * 4, cudaMemcpyDeviceToHost, st));
    OVS_CUDA_CHECK(ovs::sync_stream(st));
    float ms = 0; cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]);
    m->last_kernel_us = ms * 1000.f;
    // sequential replay: a keyframe-2 keypoint goes to its first taker (the flags live in the pinned staging area so that
    // a re-query can upload the slice of a node)
    uint8_t* const taken = hs + o_tk;
    memset(taken, 0, (size_t)R);
    std::vector<float> deltas; std::vector<int> delta_idx;
    int num = 0;
    for (int k = 0; k < Q; ++k) {
        const int i1 = q1[k];
        const unsigned* keys = m->h_keys + (size_t)k * kTriK;
        int pick = -1, seen = 0;
        for (int j = 0; j < kTriK && keys[j] != 0xffffffffu; ++j) {
            ++seen;
            const int r = 0xffff - (int)(keys[j] & 0xffffu);
            if (!taken[r]) { pick = r; break; }
        }
        if (pick < 0 && seen == kTriK) {
            // all 8 listed candidates were taken and the list may be truncated: ask the GPU again for this keypoint alone,
            // with the keyframe-2 keypoints of its node that are already claimed excluded (rare: needs 8 earlier keypoints
            // of the same node to have claimed them).  The replay is sequential, so it waits for the answer.
            ++m->num_requeries;
            const size_t len = (size_t)(seg[k].y - seg[k].x);
            OVS_CUDA_CHECK(cudaMemcpyAsync(ds + o_tk + seg[k].x, taken + seg[k].x, len, cudaMemcpyHostToDevice, st));
            TriArgs B = A;
            B.q0 = k; B.nq = k + 1; B.out_shift = Q - k; B.taken = ds + o_tk;
            k_triangulation_topk<<<1, 128, 0, st>>>(B, m->d_keys);
            OVS_LAUNCH_CHECK();
            OVS_CUDA_CHECK(cudaMemcpyAsync(m->h_keys + (size_t)Q * kTriK, m->d_keys + (size_t)Q * kTriK, kTriK * 4, cudaMemcpyDeviceToHost, st));
            OVS_CUDA_CHECK(ovs::sync_stream(st));
            const unsigned key = m->h_keys[(size_t)Q * kTriK];
            if (key != 0xffffffffu) pick = 0xffff - (int)(key & 0xffffu);
        }
        if (pick < 0) continue;
        taken[pick] = 1;
        matched_idx_2_of_1[i1] = rank2[pick];
        ++num;
        if (check_orientation) { deltas.push_back(angle_1[i1] - angle_2[rank2[pick]]); delta_idx.push_back(i1); }
    }
    if (check_orientation && !deltas.empty()) {
        std::vector<uint8_t> invalid;
        angle_checker_invalid(deltas, invalid);
        for (size_t k = 0; k < deltas.size(); ++k) if (invalid[k]) { matched_idx_2_of_1[delta_idx[k]] = -1; --num; }
    }
    *num_matches = num;
    return OVS_OK;
}

// ============================================================================ match::bow_tree
// match::bow_tree::{match_frame_and_keyframe, match_keyframes} (match/bow_tree.cc): the BoW feature vectors are inputs
// (per-keypoint vocabulary node ids, < 0 = none).  Nodes ascending, keypoints of a node in index order -- the lock-step walk
// over the two std::map<NodeId, std::vector<unsigned>> -- every query keypoint takes its nearest candidate of the same
// node that is still free, subject to HAMMING_DIST_THR_LOW and the ratio test against the second nearest free candidate.
// k_node_topk gives, per query, the 8 best (distance, visiting order) candidates of its node; the host replays the
// sequential first-taker rule on those lists and re-queries the GPU (claimed candidates excluded) only when a list cannot
// decide: the same scheme as robust::brute_force_match.
namespace {

constexpr int kNodeK = 8;

struct NodeArgs {
    int nq, q0, out_shift;
    const uint4* qdesc;            // [nq][2], visiting order
    const int2* qseg;              // [nq] candidate rank range [begin, end)
    const uint4* tdesc;            // rank order (node major, index minor)
    const unsigned char* taken;    // [R] or null
};

// one warp per query; keys = distance << 16 | rank: ascending order = the sequential loop's preference (strict '<': first wins)
__global__ void __launch_bounds__(128) k_node_topk(NodeArgs A, unsigned* __restrict__ keys_out) {
    const int q = A.q0 + blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= A.nq) return;
    const uint4 qa = A.qdesc[2 * (size_t)q], qb = A.qdesc[2 * (size_t)q + 1];
    const int2 seg = A.qseg[q];
    unsigned extra = 0xffffffffu;   // lane r < 8 carries entry r of the running top-8 from one chunk of candidates to the next
    for (int c0 = seg.x; c0 < seg.y; c0 += 32 * kNodeK) {
        unsigned mine[kNodeK];
#pragma unroll
        for (int k = 0; k < kNodeK; ++k) {
            mine[k] = 0xffffffffu;
            const int c = c0 + k * 32 + lane;
            if (c < seg.y && !(A.taken && A.taken[c])) {
                const uint4 ta = A.tdesc[2 * (size_t)c], tb = A.tdesc[2 * (size_t)c + 1];
                const int d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w)
                              + __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
                mine[k] = ((unsigned)d << 16) | (unsigned)c;
            }
        }
        unsigned next_extra = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < kNodeK; ++r) {
            unsigned lmin = extra;
#pragma unroll
            for (int k = 0; k < kNodeK; ++k) lmin = min(lmin, mine[k]);
            const unsigned wmin = __reduce_min_sync(0xffffffffu, lmin);
            if (wmin != 0xffffffffu) {
                if (extra == wmin) extra = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < kNodeK; ++k) if (mine[k] == wmin) mine[k] = 0xffffffffu;
            }
            if (lane == r) next_extra = wmin;
        }
        extra = next_extra;
    }
    if (lane < kNodeK) keys_out[(size_t)(q + A.out_shift) * kNodeK + lane] = extra;
}

// queries A (valid_a[i] != 0, node >= 0), candidates B (valid_b null or != 0, node >= 0): match_b_of_a[i] = index in B or -1
int bow_core(ovs_matcher* m, int na, const uint8_t* desc_a, const uint8_t* valid_a, const int32_t* node_a,
             int nb, const uint8_t* desc_b, const uint8_t* valid_b, const int32_t* node_b, float lowe_ratio,
             std::vector<int>& match_b_of_a, std::vector<int>& visit_order) {
    match_b_of_a.assign(std::max(na, 1), -1);
    visit_order.clear();
    if (na == 0 || nb == 0) return OVS_OK;
    OVS_CUDA_CHECK(cudaSetDevice(m->device));
    std::vector<int> rankb, qa;
    for (int i = 0; i < nb; ++i) if ((!valid_b || valid_b[i]) && node_b[i] >= 0) rankb.push_back(i);
    std::stable_sort(rankb.begin(), rankb.end(), [&](int x, int y) { return node_b[x] < node_b[y]; });
    for (int i = 0; i < na; ++i) if ((!valid_a || valid_a[i]) && node_a[i] >= 0) qa.push_back(i);
    std::stable_sort(qa.begin(), qa.end(), [&](int x, int y) { return node_a[x] < node_a[y]; });
    const int R = (int)rankb.size(), Q = (int)qa.size();
    if (R == 0 || Q == 0) return OVS_OK;
    OVS_REQUIRE(R < 65536, OVS_ERR_UNSUPPORTED, "more than 65535 candidate keypoints");
    std::vector<int2> seg(Q);
    {
        size_t lo = 0;
        for (int k = 0; k < Q; ++k) {
            const int node = node_a[qa[k]];
            while (lo < rankb.size() && node_b[rankb[lo]] < node) ++lo;
            size_t hi = lo;
            while (hi < rankb.size() && node_b[rankb[hi]] == node) ++hi;
            seg[k] = make_int2((int)lo, (int)hi);
        }
    }
    // staging (host pinned / device): [qdesc 32Q][tdesc 32R][qseg 8Q] | [taken R]
    const size_t o_qd = 0, o_td = o_qd + 32 * (size_t)Q, o_sg = o_td + 32 * (size_t)R, in_total = ((o_sg + 8 * (size_t)Q + 15) / 16) * 16,
                 o_tk = in_total, total = ((o_tk + (size_t)R + 15) / 16) * 16;
    int rc;
    if ((rc = ovs::grow_host(&m->h_stage, &m->h_stage_cap, total)) != OVS_OK) return rc;
    if ((rc = ovs::grow_dev(&m->d_q, &m->d_q_cap, total)) != OVS_OK) return rc;
    if ((rc = ovs::grow_dev(&m->d_keys, &m->d_keys_cap, (size_t)(Q + 1) * kNodeK)) != OVS_OK) return rc;
    if ((rc = ovs::grow_host(&m->h_keys, &m->h_keys_cap, (size_t)(Q + 1) * kNodeK)) != OVS_OK) return rc;
    uint8_t* hs = m->h_stage;
    for (int k = 0; k < Q; ++k) {
        memcpy(hs + o_qd + 32 * (size_t)k, desc_a + 32 * (size_t)qa[k], 32);
        reinterpret_cast<int2*>(hs + o_sg)[k] = seg[k];
    }
    for (int r = 0; r < R; ++r) memcpy(hs + o_td + 32 * (size_t)r, desc_b + 32 * (size_t)rankb[r], 32);
    cudaStream_t st = m->stream;
    uint8_t* ds = m->d_q;
    OVS_CUDA_CHECK(cudaMemcpyAsync(ds, hs, in_total, cudaMemcpyHostToDevice, st));
    NodeArgs A{};
    A.nq = Q; A.qdesc = reinterpret_cast<const uint4*>(ds + o_qd); A.tdesc = reinterpret_cast<const uint4*>(ds + o_td);
    A.qseg = reinterpret_cast<const int2*>(ds + o_sg);
    OVS_CUDA_CHECK(cudaEventRecord(m->ev[0], st));
    k_node_topk<<<(Q + 3) / 4, 128, 0, st>>>(A, m->d_keys);
    OVS_LAUNCH_CHECK();
    OVS_CUDA_CHECK(cudaEventRecord(m->ev[1], st));
    OVS_CUDA_CHECK(cudaMemcpyAsync(m->h_keys, m->d_keys, (size_t)Q * kNodeK * 4, cudaMemcpyDeviceToHost, st));
    OVS_CUDA_CHECK(ovs::sync_stream(st));
    float ms = 0; cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]);
    m->last_kernel_us = ms * 1000.f;
    // a candidate farther than d_star can neither be an acceptable best nor make the ratio test fail (see brute_force_match)
    int d_star = OVS_HAMMING_DIST_THR_LOW + 1;
    while (d_star < OVS_MAX_HAMMING_DIST && lowe_ratio * (float)(unsigned)d_star < (float)OVS_HAMMING_DIST_THR_LOW) ++d_star;
    ++d_star;
    uint8_t* const taken = hs + o_tk;
    memset(taken, 0, (size_t)R);
    for (int k = 0; k < Q; ++k) {
        unsigned keys[kNodeK];
        memcpy(keys, m->h_keys + (size_t)k * kNodeK, sizeof(keys));
        for (int attempt = 0; attempt < 2; ++attempt) {
            int rem[kNodeK], r = 0;
            bool exhausted = false;
            for (int j = 0; j < kNodeK; ++j) {
                if (keys[j] == 0xffffffffu) { exhausted = true; break; }
                if (!taken[keys[j] & 0xffffu]) rem[r++] = j;
            }
            const int lb = exhausted ? OVS_MAX_HAMMING_DIST : (int)(keys[kNodeK - 1] >> 16);   // unlisted candidates are >= this
            const bool complete = exhausted || lb >= d_star || attempt == 1;
            int best = OVS_MAX_HAMMING_DIST, best_r = -1, second = OVS_MAX_HAMMING_DIST;
            bool decided = true;
            if (r >= 2 || complete) {
                if (r >= 1) { best = (int)(keys[rem[0]] >> 16); best_r = (int)(keys[rem[0]] & 0xffffu); }
                if (r >= 2) second = (int)(keys[rem[1]] >> 16);
            } else if (r == 1) {
                best = (int)(keys[rem[0]] >> 16); best_r = (int)(keys[rem[0]] & 0xffffu);
                if (best <= OVS_HAMMING_DIST_THR_LOW && lowe_ratio * (float)(unsigned)lb < (float)best) decided = false;   // needs the true second best
                second = lb;
            } else if (lb <= OVS_HAMMING_DIST_THR_LOW) {
                decided = false;
            }
            if (!decided) {
                ++m->num_requeries;
                const size_t len = (size_t)(seg[k].y - seg[k].x);
                OVS_CUDA_CHECK(cudaMemcpyAsync(ds + o_tk + seg[k].x, taken + seg[k].x, len, cudaMemcpyHostToDevice, st));
                NodeArgs B = A;
                B.q0 = k; B.nq = k + 1; B.out_shift = Q - k; B.taken = ds + o_tk;
                k_node_topk<<<1, 128, 0, st>>>(B, m->d_keys);
                OVS_LAUNCH_CHECK();
                OVS_CUDA_CHECK(cudaMemcpyAsync(m->h_keys + (size_t)Q * kNodeK, m->d_keys + (size_t)Q * kNodeK, kNodeK * 4, cudaMemcpyDeviceToHost, st));
                OVS_CUDA_CHECK(ovs::sync_stream(st));
                memcpy(keys, m->h_keys + (size_t)Q * kNodeK, sizeof(keys));
                continue;
            }
            if (OVS_HAMMING_DIST_THR_LOW < best) break;
            if (lowe_ratio * (float)(unsigned)second < (float)best) break;
            taken[best_r] = 1;
            match_b_of_a[qa[k]] = rankb[best_r];
            visit_order.push_back(qa[k]);
            break;
        }
    }
    return OVS_OK;
}

}  // namespace

// bow_tree::match_frame_and_keyframe(keyfrm, frm, matched_lms_in_frm): matched_keyfrm_idx_of_frm[i] = the keyframe keypoint
// whose landmark frame keypoint i receives, or -1.  lm_valid_kf[k]: keyfrm landmark k non-null and not will_be_erased().
extern "C" int ovs_bow_tree_match_frame_and_keyframe_host(ovs_matcher* m, int n_kf, const uint8_t* desc_kf, const float* angle_kf, const uint8_t* lm_valid_kf,
                                                          const int32_t* bow_node_kf, int n_frm, const uint8_t* desc_frm, const float* angle_frm,
                                                          const int32_t* bow_node_frm, float lowe_ratio, int check_orientation,
                                                          int32_t* matched_keyfrm_idx_of_frm, int* num_matches) {
    OVS_REQUIRE(m && num_matches && n_kf >= 0 && n_frm >= 0 && (n_frm == 0 || matched_keyfrm_idx_of_frm), OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(n_kf == 0 || (desc_kf && angle_kf && lm_valid_kf && bow_node_kf), OVS_ERR_INVALID_ARG, "null keyframe array");
    OVS_REQUIRE(n_frm == 0 || (desc_frm && angle_frm && bow_node_frm), OVS_ERR_INVALID_ARG, "null frame array");
    *num_matches = 0;
    for (int i = 0; i < n_frm; ++i) matched_keyfrm_idx_of_frm[i] = -1;
    std::vector<int> match, order;
    const int rc = bow_core(m, n_kf, desc_kf, lm_valid_kf, bow_node_kf, n_frm, desc_frm, nullptr, bow_node_frm, lowe_ratio, match, order);
    if (rc != OVS_OK) return rc;
    int num = 0;
    std::vector<float> deltas; std::vector<int> delta_idx;
    for (int k : order) {
        const int f = match[k];
        matched_keyfrm_idx_of_frm[f] = k; ++num;
        if (check_orientation) { deltas.push_back(angle_kf[k] - angle_frm[f]); delta_idx.push_back(f); }
    }
    if (check_orientation && !deltas.empty()) {
        std::vector<uint8_t> invalid;
        angle_checker_invalid(deltas, invalid);
        for (size_t k = 0; k < deltas.size(); ++k) if (invalid[k]) { matched_keyfrm_idx_of_frm[delta_idx[k]] = -1; --num; }
    }
    *num_matches = num;
    return OVS_OK;
}

// bow_tree::match_keyframes(keyfrm_1, keyfrm_2, matched_lms_in_keyfrm_1): both keypoints need a valid landmark, a keyframe-2
// keypoint is matched at most once.  matched_idx_2_of_1[i1] = keypoint of keyframe 2 or -1.
extern "C" int ovs_bow_tree_match_keyframes_host(ovs_matcher* m, int n1, const uint8_t* desc_1, const float* angle_1, const uint8_t* lm_valid_1,
                                                 const int32_t* bow_node_1, int n2, const uint8_t* desc_2, const float* angle_2, const uint8_t* lm_valid_2,
                                                 const int32_t* bow_node_2, float lowe_ratio, int check_orientation,
                                                 int32_t* matched_idx_2_of_1, int* num_matches) {
    OVS_REQUIRE(m && num_matches && n1 >= 0 && n2 >= 0 && (n1 == 0 || matched_idx_2_of_1), OVS_ERR_INVALID_ARG, "bad argument");
    OVS_REQUIRE(n1 == 0 || (desc_1 && angle_1 && lm_valid_1 && bow_node_1), OVS_ERR_INVALID_ARG, "null keyframe-1 array");
    OVS_REQUIRE(n2 == 0 || (desc_2 && angle_2 && lm_valid_2 && bow_node_2), OVS_ERR_INVALID_ARG, "null keyframe-2 array");
    *num_matches = 0;
    for (int i = 0; i < n1; ++i) matched_idx_2_of_1[i] = -1;
    std::vector<int> match, order;
    const int rc = bow_core(m, n1, desc_1, lm_valid_1, bow_node_1, n2, desc_2, lm_valid_2, bow_node_2, lowe_ratio, match, order);
    if (rc != OVS_OK) return rc;
    int num = 0;
    std::vector<float> deltas; std::vector<int> delta_idx;
    for (int i1 : order) {
        matched_idx_2_of_1[i1] = match[i1]; ++num;
        if (check_orientation) { deltas.push_back(angle_1[i1] - angle_2[match[i1]]); delta_idx.push_back(i1); }
    }
    if (check_orientation && !deltas.empty()) {
        std::vector<uint8_t> invalid;
        angle_checker_invalid(deltas, invalid);
        for (size_t k = 0; k < deltas.size(); ++k) if (invalid[k]) { matched_idx_2_of_1[delta_idx[k]] = -1; --num; }
    }
    *num_matches = num;
    return OVS_OK;
}
