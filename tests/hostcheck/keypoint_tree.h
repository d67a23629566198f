// keypoint_tree.h -- TEST INFRASTRUCTURE: list-based host restatement of orb_extractor::distribute_keypoints_via_tree
// (the product runs k_tree_distribute on the GPU; this was the product's host code in round 1 and is kept as a checker)
// (feature/orb_extractor.cc, feature/orb_extractor_node.cc; names as in SURVEY.md 8a).
//
// The reference keeps a std::list of nodes, each owning a std::vector<cv::KeyPoint>, and
// splits nodes until the requested number of leaves is reached, then keeps the best-response
// keypoint per leaf.  The procedure is sequential and order dependent (SURVEY.md section 7,
// "hard parts"), and runs on the compacted FAST candidates (tens of thousands of points, not
// pixels), so it stays on the host; the GPU produces the candidate list and consumes the
// selection.  Layout here is index-based: one permutation array partitioned in place (stable
// 4-way counting split), nodes in a flat vector linked by indices.
#pragma once
#include <stdint.h>
#include <vector>

namespace ovs {

// FAST candidate as produced by the cell NMS kernel: x | y << 12 | score << 24, x and y
// relative to the 19 px level border (the reference's keypts_to_distribute coordinates).
inline int cand_x(uint32_t c) { return (int)(c & 0xfffu); }
inline int cand_y(uint32_t c) { return (int)((c >> 12) & 0xfffu); }
inline int cand_score(uint32_t c) { return (int)(c >> 24); }

struct TreeNode {
    int bx, by, ex, ey;  // pt_begin_, pt_end_
    int begin, count;    // range in the permutation buffer `buf`
    int prev, next;      // list links (-1 = none)
    int serial;          // creation order (stands in for the heap address tie-break)
    uint8_t buf;         // which of the two permutation buffers holds the range
    bool leaf;
    int ccnt[4];         // keypoints per quadrant of THIS node, counted while its parent scattered them (one sweep per level)
};

struct TreeScratch {
    std::vector<int> perm[2];
    std::vector<TreeNode> nodes;
    std::vector<int> pool, prev_pool;
};

// Writes the indices (into cand[]) of the selected keypoints, in the reference's output order,
// to out[] (capacity >= num_keypts + 3).  Returns the number selected.
int distribute_keypoints_via_tree(const uint32_t* cand, int n, int min_x, int max_x, int min_y, int max_y,
                                  unsigned num_keypts, int* out, TreeScratch& s);

}  // namespace ovs
