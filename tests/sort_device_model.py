"""numpy model of the graph preparation's stable radix sort (k_sort_hist / k_sort_tile_prefix / k_sort_scatter in
openvslam_b200/csrc/optimize.cu): the same tiles, warps, chunks and counters, with match.any restated as "lanes of the chunk that
hold the same digit".  tests/test_sort_device_model.py pins the formulation against numpy's stable argsort on the CPU; the
kernels themselves are pinned on the GPU (test_graph_preparation_sort_is_a_stable_sort)."""
import numpy as np

BITS = 11
BINS = 1 << BITS
TILE = 2048
WARPS = 8


def one_pass(keys, vals, shift):
    n = len(keys)
    ntiles = (n + TILE - 1) // TILE
    dig = ((keys >> np.uint32(shift)) & np.uint32(BINS - 1)).astype(np.int64)
    hist = np.zeros((ntiles, BINS), np.int64)                       # k_sort_hist
    for t in range(ntiles):
        hist[t] = np.bincount(dig[t * TILE:(t + 1) * TILE], minlength=BINS)
    bin_total = hist.sum(0)                                         # k_sort_tile_prefix
    tile_prefix = np.cumsum(hist, 0) - hist
    bin_base = np.cumsum(bin_total) - bin_total                     # k_sort_scatter: scan of the totals
    ko = np.empty_like(keys); vo = np.empty_like(vals)
    written = np.zeros(n, bool)
    for t in range(ntiles):
        s_base = bin_base + tile_prefix[t]
        cnt = np.zeros((WARPS, BINS), np.int64)
        for w in range(WARPS):                                      # per-warp digit counts
            lo = t * TILE + w * 256
            d = dig[lo:min(lo + 256, n)]
            if len(d):
                cnt[w] = np.bincount(d, minlength=BINS)
        cnt = np.cumsum(cnt, 0) - cnt                               # exclusive prefix over the warps
        for w in range(WARPS):
            for c in range(8):                                      # chunks of 32 consecutive entries, in order
                base = t * TILE + w * 256 + c * 32
                lanes = np.arange(base, min(base + 32, n))
                if len(lanes) == 0:
                    continue
                d = dig[lanes]
                for k, idx in enumerate(lanes):
                    rank = int(np.sum(d[:k] == d[k]))               # popc(match.any mask & lanes below)
                    pos = s_base[d[k]] + cnt[w][d[k]] + rank
                    assert not written[pos]
                    written[pos] = True
                    ko[pos] = keys[idx]; vo[pos] = vals[idx]
                np.add.at(cnt[w], d, 1)                             # the leader lane of every digit adds the chunk's count
    assert written.all()
    return ko, vo


def sort_pairs(keys, vals, end_bit):
    shift = 0
    while shift < end_bit:
        keys, vals = one_pass(keys, vals, shift)
        shift += BITS
    return keys, vals
