// keypoint_tree.cpp -- see keypoint_tree.h.
#include "keypoint_tree.h"

#include <algorithm>
#include <cmath>

namespace ovs {
namespace {

struct List {
    int head = -1, tail = -1, size = 0;
};

inline void push_back(std::vector<TreeNode>& N, List& L, int id) {
    N[id].prev = L.tail; N[id].next = -1;
    if (L.tail >= 0) N[L.tail].next = id; else L.head = id;
    L.tail = id; ++L.size;
}
inline void push_front(std::vector<TreeNode>& N, List& L, int id) {
    N[id].next = L.head; N[id].prev = -1;
    if (L.head >= 0) N[L.head].prev = id; else L.tail = id;
    L.head = id; ++L.size;
}
inline int erase(std::vector<TreeNode>& N, List& L, int id) {
    const int nx = N[id].next, pv = N[id].prev;
    if (pv >= 0) N[pv].next = nx; else L.head = nx;
    if (nx >= 0) N[nx].prev = pv; else L.tail = pv;
    --L.size;
    return nx;
}

// split centre of a node box, as orb_extractor_node::divide_node computes it
inline void centre_of(int bx, int by, int ex, int ey, int* cx, int* cy) {
    *cx = bx + (int)std::ceil((ex - bx) / 2.0);
    *cy = by + (int)std::ceil((ey - by) / 2.0);
}

// orb_extractor_node::divide_node + orb_extractor::assign_child_nodes: split node `id` into up
// to four children (stable partition of its index range into the other permutation buffer),
// push the non-empty ones to the list front in child order 0..3, and record those with more
// than one keypoint in `pool`.  The node's own quadrant counts were taken while its parent
// scattered it (one sweep per tree level instead of two); the scatter goes through four running
// pointers -- candidates arrive in row-major runs, so the quadrant branches predict well and no
// store-to-load chain through a position array is left.
inline void divide_and_assign(const uint32_t* cand, TreeScratch& s, List& L, int id, int& serial, std::vector<int>& pool) {
    const TreeNode nd = s.nodes[id];  // copy: nodes may reallocate below
    int cx, cy;
    centre_of(nd.bx, nd.by, nd.ex, nd.ey, &cx, &cy);
    const int* src = s.perm[nd.buf].data() + nd.begin;
    int* dst = s.perm[nd.buf ^ 1].data() + nd.begin;
    const int* cnt = nd.ccnt;
    const int off[4] = {0, cnt[0], cnt[0] + cnt[1], cnt[0] + cnt[1] + cnt[2]};
    const int bxs[4] = {nd.bx, cx, nd.bx, cx}, bys[4] = {nd.by, nd.by, cy, cy};
    const int exs[4] = {cx, nd.ex, cx, nd.ex}, eys[4] = {cy, cy, nd.ey, nd.ey};
    int ccx[4], ccy[4], gcnt[4][4] = {};
    for (int k = 0; k < 4; ++k) centre_of(bxs[k], bys[k], exs[k], eys[k], &ccx[k], &ccy[k]);
    int* p0 = dst + off[0]; int* p1 = dst + off[1]; int* p2 = dst + off[2]; int* p3 = dst + off[3];
    for (int i = 0; i < nd.count; ++i) {
        const int v = src[i];
        const uint32_t c = cand[v];
        const int x = cand_x(c), y = cand_y(c);
        if (cy <= y) {
            if (cx <= x) { *p3++ = v; ++gcnt[3][(ccx[3] <= x ? 1 : 0) + (ccy[3] <= y ? 2 : 0)]; }
            else         { *p2++ = v; ++gcnt[2][(ccx[2] <= x ? 1 : 0) + (ccy[2] <= y ? 2 : 0)]; }
        } else {
            if (cx <= x) { *p1++ = v; ++gcnt[1][(ccx[1] <= x ? 1 : 0) + (ccy[1] <= y ? 2 : 0)]; }
            else         { *p0++ = v; ++gcnt[0][(ccx[0] <= x ? 1 : 0) + (ccy[0] <= y ? 2 : 0)]; }
        }
    }
    for (int k = 0; k < 4; ++k) {
        const int my_serial = serial++;
        if (cnt[k] == 0) continue;
        TreeNode ch;
        ch.bx = bxs[k]; ch.by = bys[k]; ch.ex = exs[k]; ch.ey = eys[k];
        ch.begin = nd.begin + off[k]; ch.count = cnt[k];
        ch.prev = ch.next = -1;
        ch.serial = my_serial;
        ch.buf = nd.buf ^ 1;
        ch.leaf = (cnt[k] == 1);
        for (int g = 0; g < 4; ++g) ch.ccnt[g] = gcnt[k][g];
        const int cid = (int)s.nodes.size();
        s.nodes.push_back(ch);
        push_front(s.nodes, L, cid);
        if (cnt[k] > 1) pool.push_back(cid);
    }
}

}  // namespace

int distribute_keypoints_via_tree(const uint32_t* cand, int n, int min_x, int max_x, int min_y, int max_y,
                                  unsigned num_keypts, int* out, TreeScratch& s) {
    if (n <= 0) return 0;
    s.perm[0].resize(n); s.perm[1].resize(n);
    s.nodes.clear(); s.nodes.reserve(4 * (size_t)n + 64);
    s.pool.clear(); s.prev_pool.clear();
    List L;
    int serial = 0;

    // initialize_nodes
    const double ratio = (double)(max_x - min_x) / (max_y - min_y);
    double delta_x, delta_y;
    unsigned gx, gy;
    if (ratio > 1) {
        gx = (unsigned)std::round(ratio); gy = 1;
        delta_x = (double)(max_x - min_x) / gx; delta_y = max_y - min_y;
    } else {
        gx = 1; gy = (unsigned)std::round(1 / ratio);
        delta_x = max_x - min_x; delta_y = (double)(max_y - min_y) / gy;
    }
    const unsigned nini = gx * gy;
    // bucket candidates by initial node, keeping input order (counting sort)
    std::vector<int>& key = s.prev_pool;  // reuse as scratch
    key.resize(n);
    std::vector<int> cnt(nini + 1, 0);
    // node of a candidate = (unsigned)((float)x / delta_x) + (unsigned)((float)y / delta_y) * gx: the two divisions are
    // tabulated per coordinate value (a few thousand divisions per call instead of two per candidate)
    const int xs = std::max(max_x - min_x, 0) + 1, ys = std::max(max_y - min_y, 0) + 1;
    std::vector<unsigned> tabx(xs + 1), taby(ys + 1);
    for (int x = 0; x <= xs; ++x) tabx[x] = (unsigned)((float)x / delta_x);
    for (int y = 0; y <= ys; ++y) taby[y] = (unsigned)((float)y / delta_y) * gx;
    for (int i = 0; i < n; ++i) {
        const int x = cand_x(cand[i]), y = cand_y(cand[i]);
        const unsigned ix = x <= xs ? tabx[x] : (unsigned)((float)x / delta_x);
        const unsigned iy = y <= ys ? taby[y] : (unsigned)((float)y / delta_y) * gx;
        unsigned k = ix + iy;
        if (k >= nini) k = nini - 1;
        key[i] = (int)k;
        ++cnt[k + 1];
    }
    for (unsigned k = 0; k < nini; ++k) cnt[k + 1] += cnt[k];
    // boxes and split centres of the initial nodes, then one sweep that scatters the candidates and counts their quadrants
    std::vector<int> ibx(nini), iby(nini), iex(nini), iey(nini), icx(nini), icy(nini), iq(4 * (size_t)nini, 0);
    for (unsigned i = 0; i < nini; ++i) {
        const unsigned ix = i % gx, iy = i / gx;
        ibx[i] = (int)(delta_x * ix); iby[i] = (int)(delta_y * iy);
        iex[i] = (int)(delta_x * (ix + 1)); iey[i] = (int)(delta_y * (iy + 1));
        centre_of(ibx[i], iby[i], iex[i], iey[i], &icx[i], &icy[i]);
    }
    {
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        for (int i = 0; i < n; ++i) {
            const int k = key[i];
            s.perm[0][pos[k]++] = i;
            ++iq[4 * (size_t)k + (icx[k] <= cand_x(cand[i]) ? 1 : 0) + (icy[k] <= cand_y(cand[i]) ? 2 : 0)];
        }
    }
    for (unsigned i = 0; i < nini; ++i) {
        const int my_serial = serial++;
        const int c = cnt[i + 1] - cnt[i];
        if (c == 0) continue;  // empty initial nodes are erased before any split
        TreeNode nd;
        nd.bx = ibx[i]; nd.by = iby[i]; nd.ex = iex[i]; nd.ey = iey[i];
        nd.begin = cnt[i]; nd.count = c; nd.prev = nd.next = -1;
        nd.serial = my_serial; nd.buf = 0; nd.leaf = (c == 1);
        for (int g = 0; g < 4; ++g) nd.ccnt[g] = iq[4 * (size_t)i + g];
        s.nodes.push_back(nd);
        push_back(s.nodes, L, (int)s.nodes.size() - 1);
    }
    key.clear();

    bool is_filled = false;
    for (;;) {
        const int prev_size = L.size;
        s.pool.clear();
        for (int it = L.head; it >= 0;) {
            if (s.nodes[it].leaf) { it = s.nodes[it].next; continue; }
            divide_and_assign(cand, s, L, it, serial, s.pool);
            it = erase(s.nodes, L, it);
        }
        if ((long)num_keypts <= (long)L.size || L.size == prev_size) { is_filled = true; break; }
        if ((long)num_keypts < (long)L.size + 3L * (long)s.pool.size()) { is_filled = false; break; }
    }
    while (!is_filled) {
        const int prev_size = L.size;
        s.prev_pool.swap(s.pool);
        s.pool.clear();
        // std::sort(rbegin, rend) on pair<count, node*>: descending by (count, address)
        std::sort(s.prev_pool.begin(), s.prev_pool.end(), [&](int a, int b) {
            const TreeNode& A = s.nodes[a]; const TreeNode& B = s.nodes[b];
            if (A.count != B.count) return A.count > B.count;
            return A.serial > B.serial;
        });
        for (size_t i = 0; i < s.prev_pool.size(); ++i) {
            const int id = s.prev_pool[i];
            divide_and_assign(cand, s, L, id, serial, s.pool);
            erase(s.nodes, L, id);
            if ((long)num_keypts <= (long)L.size) { is_filled = true; break; }
        }
        if (is_filled || (long)num_keypts <= (long)L.size || L.size == prev_size) { is_filled = true; break; }
    }

    // find_keypoints_with_max_response: first strict maximum in each node, in list order
    int nout = 0;
    for (int it = L.head; it >= 0; it = s.nodes[it].next) {
        const TreeNode& nd = s.nodes[it];
        const int* p = s.perm[nd.buf].data() + nd.begin;
        int best = p[0];
        for (int k = 1; k < nd.count; ++k)
            if (cand_score(cand[p[k]]) > cand_score(cand[best])) best = p[k];
        out[nout++] = best;
    }
    return nout;
}

}  // namespace ovs
