// tests/hostcheck/tree_levelsync.cpp -- round-2 design prototype, NOT product code.
//
// orb_extractor::distribute_keypoints_via_tree restated as a sequence of ARRAY passes (counting, prefix sums, stable
// partitions, one sort per finishing round) instead of std::list surgery, i.e. in the shape a device-side version needs:
// every loop below is either "for each node / keypoint independently" or a scan / sort / compaction.  The test suite
// checks it against the list-based host implementation (keypoint_tree.cpp) and against the oracle on
// random candidate sets; porting it to CUDA (one CTA per pyramid level) removes the 0.45 ms of host work and the two
// synchronisations per frame that the extractor has today.
//
// How the list order is reproduced without a list.  A pass over the list visits the nodes in order; every node with more
// than one keypoint is split, its non-empty children are pushed to the FRONT (child 0 first, so child 3 ends up
// frontmost) and the node is erased; children are not visited in the same pass.  Hence after a pass over list A:
//     new list = [children of the LAST split node (3,2,1,0)] ... [children of the FIRST split node (3,2,1,0)] ++ [unsplit nodes of A, in order]
// and the serial numbers (creation order, the stand-in for the heap-address tie-break) advance by 4 per split node in
// visiting order.  The finishing rounds split the pool in descending (count, serial) order until the target count is
// reached: the number of nodes each split adds is known before splitting (non-empty children - 1), so the cut-off index
// is a prefix-sum search; the split nodes are erased wherever they sit (a compaction) and their children go to the front
// in the same reversed order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace {

inline int cx_of(uint32_t c) { return (int)(c & 0xfffu); }
inline int cy_of(uint32_t c) { return (int)((c >> 12) & 0xfffu); }
inline int cs_of(uint32_t c) { return (int)(c >> 24); }

struct Node {
    int bx, by, ex, ey;
    int begin, count;   // range in perm[buf]
    int serial;
    int buf;
};

struct Split {          // the result of dividing one node: up to four children, in child order
    Node child[4];
    int nonempty;       // how many children have keypoints
};

// divide_node + assign_child_nodes for one node: a stable 4-way partition of its range into the other buffer
Split divide(const uint32_t* cand, const Node& nd, std::vector<int> (&perm)[2], int first_serial) {
    Split s{};
    const int half_x = (int)std::ceil((nd.ex - nd.bx) / 2.0), half_y = (int)std::ceil((nd.ey - nd.by) / 2.0);
    const int cx = nd.bx + half_x, cy = nd.by + half_y;
    const int* src = perm[nd.buf].data() + nd.begin;
    int* dst = perm[nd.buf ^ 1].data() + nd.begin;
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < nd.count; ++i) { const uint32_t c = cand[src[i]]; ++cnt[(cx <= cx_of(c) ? 1 : 0) + (cy <= cy_of(c) ? 2 : 0)]; }
    const int off[4] = {0, cnt[0], cnt[0] + cnt[1], cnt[0] + cnt[1] + cnt[2]};
    int pos[4] = {off[0], off[1], off[2], off[3]};
    for (int i = 0; i < nd.count; ++i) { const uint32_t c = cand[src[i]]; dst[pos[(cx <= cx_of(c) ? 1 : 0) + (cy <= cy_of(c) ? 2 : 0)]++] = src[i]; }
    const int bxs[4] = {nd.bx, cx, nd.bx, cx}, bys[4] = {nd.by, nd.by, cy, cy}, exs[4] = {cx, nd.ex, cx, nd.ex}, eys[4] = {cy, cy, nd.ey, nd.ey};
    for (int k = 0; k < 4; ++k) {
        s.child[k] = Node{bxs[k], bys[k], exs[k], eys[k], nd.begin + off[k], cnt[k], first_serial + k, nd.buf ^ 1};
        s.nonempty += cnt[k] > 0;
    }
    return s;
}

}  // namespace

extern "C" int hc_distribute_levelsync(const uint32_t* cand, int n, int min_x, int max_x, int min_y, int max_y, unsigned num_keypts, int* out) {
    if (n <= 0) return 0;
    std::vector<int> perm[2];
    perm[0].resize(n); perm[1].resize(n);
    int serial = 0;
    // ---- initialize_nodes: bucket by initial node (counting sort keeps the input order)
    const double ratio = (double)(max_x - min_x) / (max_y - min_y);
    double delta_x, delta_y; unsigned gx, gy;
    if (ratio > 1) { gx = (unsigned)std::round(ratio); gy = 1; delta_x = (double)(max_x - min_x) / gx; delta_y = max_y - min_y; }
    else { gx = 1; gy = (unsigned)std::round(1 / ratio); delta_x = max_x - min_x; delta_y = (double)(max_y - min_y) / gy; }
    const unsigned nini = gx * gy;
    std::vector<int> key(n), cnt(nini + 1, 0);
    for (int i = 0; i < n; ++i) {
        unsigned k = (unsigned)((float)cx_of(cand[i]) / delta_x) + (unsigned)((float)cy_of(cand[i]) / delta_y) * gx;
        if (k >= nini) k = nini - 1;
        key[i] = (int)k; ++cnt[k + 1];
    }
    std::partial_sum(cnt.begin(), cnt.end(), cnt.begin());
    { std::vector<int> pos(cnt.begin(), cnt.end() - 1); for (int i = 0; i < n; ++i) perm[0][pos[key[i]]++] = i; }
    std::vector<Node> list;   // the node list, in list order
    for (unsigned i = 0; i < nini; ++i) {
        const int my_serial = serial++;
        const int c = cnt[i + 1] - cnt[i];
        if (c == 0) continue;
        const unsigned ix = i % gx, iy = i / gx;
        list.push_back(Node{(int)(delta_x * ix), (int)(delta_y * iy), (int)(delta_x * (ix + 1)), (int)(delta_y * (iy + 1)), cnt[i], c, my_serial, 0});
    }

    // appends the children of the given splits (already in visiting order) to the FRONT of `rest`, reversed as push_front does
    auto rebuild = [](const std::vector<Split>& splits, const std::vector<Node>& rest, std::vector<Node>& pool_out) {
        std::vector<Node> nl;
        for (int i = (int)splits.size() - 1; i >= 0; --i)
            for (int k = 3; k >= 0; --k)
                if (splits[i].child[k].count > 0) nl.push_back(splits[i].child[k]);
        nl.insert(nl.end(), rest.begin(), rest.end());
        for (const Split& s : splits)                       // the pool keeps CREATION order
            for (int k = 0; k < 4; ++k)
                if (s.child[k].count > 1) pool_out.push_back(s.child[k]);
        return nl;
    };

    std::vector<Node> pool;
    bool is_filled = false;
    for (;;) {                                              // ---- whole passes
        const size_t prev_size = list.size();
        std::vector<int> multi;                             // positions of the nodes to split, in list order (a compaction)
        for (size_t i = 0; i < list.size(); ++i) if (list[i].count > 1) multi.push_back((int)i);
        std::vector<Split> splits(multi.size());
        for (size_t j = 0; j < multi.size(); ++j) splits[j] = divide(cand, list[multi[j]], perm, serial + 4 * (int)j);   // independent
        serial += 4 * (int)multi.size();
        std::vector<Node> rest;
        for (const Node& nd : list) if (nd.count == 1) rest.push_back(nd);
        pool.clear();
        list = rebuild(splits, rest, pool);
        if ((long)num_keypts <= (long)list.size() || list.size() == prev_size) { is_filled = true; break; }
        if ((long)num_keypts < (long)list.size() + 3L * (long)pool.size()) break;
    }
    while (!is_filled) {                                    // ---- finishing rounds
        const size_t prev_size = list.size();
        std::vector<Node> prev_pool;
        prev_pool.swap(pool);
        std::sort(prev_pool.begin(), prev_pool.end(), [](const Node& a, const Node& b) { return a.count != b.count ? a.count > b.count : a.serial > b.serial; });
        // split everything (independent), then find how many splits the sequential loop would have performed
        std::vector<Split> splits(prev_pool.size());
        for (size_t j = 0; j < prev_pool.size(); ++j) splits[j] = divide(cand, prev_pool[j], perm, serial + 4 * (int)j);
        size_t used = prev_pool.size();
        long size = (long)list.size();
        for (size_t j = 0; j < prev_pool.size(); ++j) {     // a prefix sum + first-index search
            size += splits[j].nonempty - 1;
            if ((long)num_keypts <= size) { used = j + 1; is_filled = true; break; }
        }
        splits.resize(used);
        serial += 4 * (int)used;
        // erase the split nodes wherever they are (they are identified by their serial), keep the others in order
        std::vector<int> gone(used);
        for (size_t j = 0; j < used; ++j) gone[j] = prev_pool[j].serial;
        std::sort(gone.begin(), gone.end());
        std::vector<Node> rest;
        for (const Node& nd : list) if (!std::binary_search(gone.begin(), gone.end(), nd.serial)) rest.push_back(nd);
        pool.clear();
        list = rebuild(splits, rest, pool);
        if (is_filled || (long)num_keypts <= (long)list.size() || list.size() == prev_size) break;
    }
    // ---- find_keypoints_with_max_response: first strict maximum per node, in list order (independent per node)
    int nout = 0;
    for (const Node& nd : list) {
        const int* p = perm[nd.buf].data() + nd.begin;
        int best = p[0];
        for (int k = 1; k < nd.count; ++k) if (cs_of(cand[p[k]]) > cs_of(cand[best])) best = p[k];
        out[nout++] = best;
    }
    return nout;
}
