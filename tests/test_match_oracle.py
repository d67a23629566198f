"""The matcher oracle (oracle/match_oracle.c) against independent numpy restatements on small cases -- the oracle is what
the GPU parity tests trust, so its own logic is cross-checked here without a GPU: Hamming distance, the brute-force
matcher's greedy rule, get_keypoints_in_cell, the mutual projection matcher and the triangulation matcher."""
import numpy as np
import pytest

from openvslam_b200 import synth


def _ham(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def test_hamming_and_bruteforce(oracle):
    rng = np.random.default_rng(0)
    d1 = rng.integers(0, 256, (60, 32), dtype=np.uint8); d2 = rng.integers(0, 256, (70, 32), dtype=np.uint8)
    d2[:30] = d1[:30]; d2[:30, 3] ^= rng.integers(0, 8, 30).astype(np.uint8)          # near duplicates
    D = np.array([[_ham(a, b) for b in d2] for a in d1])
    for i in (0, 7, 59):
        for j in (0, 11, 69):
            assert oracle.hamming(d1[i], d2[j]) == D[i, j]
    # robust::brute_force_match: for each keyframe keypoint (idx_2 ascending) best / second over the still free frame
    # keypoints, thresholds, ratio test, greedy claim
    valid = (rng.random(70) < 0.8).astype(np.uint8)
    pairs = oracle.robust_brute_force_match(d1, d2, valid, 0.8)
    free = np.ones(60, bool); ref = []
    for j in range(70):
        if not valid[j]:
            continue
        cand = [(D[i, j], i) for i in range(60) if free[i]]
        best = min(cand)                                    # lowest index wins ties, like the strict `<` scan
        rest = [c for c in cand if c[1] != best[1]]
        second = min(rest)[0] if rest else 256
        if best[0] > 50 or 0.8 * second < best[0]:
            continue
        ref.append((best[1], j)); free[best[1]] = False
    assert [tuple(p) for p in pairs] == ref and len(ref) >= 20


def test_get_keypoints_in_cell_is_a_box_query(oracle):
    rng = np.random.default_rng(1)
    n = 800
    x = rng.uniform(0, 752, n).astype(np.float32); y = rng.uniform(0, 480, n).astype(np.float32)
    octv = rng.integers(0, 8, n).astype(np.int32)
    f = oracle.MatchFrame(x, y, octv, np.zeros(n, np.float32), None, rng.integers(0, 256, (n, 32), dtype=np.uint8), oracle.om_grid(0, 752, 0, 480))
    for _ in range(40):
        rx, ry, m = rng.uniform(-20, 770), rng.uniform(-20, 500), rng.uniform(3, 60)
        lo = int(rng.integers(0, 6)); hi = lo + int(rng.integers(0, 3))
        got = oracle.get_keypoints_in_cell(f, rx, ry, m, lo, hi)
        assert np.array_equal(got, oracle.get_keypoints_in_cell_literal(f, rx, ry, m, lo, hi))   # same candidates, same visiting order
        inside = (np.abs(x - np.float32(rx)) < np.float32(m)) & (np.abs(y - np.float32(ry)) < np.float32(m)) & (octv >= lo) & (octv <= hi)
        assert sorted(got.tolist()) == np.flatnonzero(inside).tolist()


def test_cell_list_cache_follows_the_frame(oracle):
    """The oracle caches keypt_indices_in_cells_ per frame (as data::frame does): a different frame at the same addresses, or
    the same arrays with other coordinates, must rebuild it."""
    rng = np.random.default_rng(11)
    n = 500
    x = rng.uniform(0, 752, n).astype(np.float32); y = rng.uniform(0, 480, n).astype(np.float32)
    octv = rng.integers(0, 8, n).astype(np.int32); d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    g = oracle.om_grid(0, 752, 0, 480)
    f = oracle.MatchFrame(x, y, octv, np.zeros(n, np.float32), None, d, g)
    for trial in range(6):
        for _ in range(10):
            rx, ry, m = rng.uniform(0, 752), rng.uniform(0, 480), rng.uniform(5, 50)
            assert np.array_equal(oracle.get_keypoints_in_cell(f, rx, ry, m, -1, -1), oracle.get_keypoints_in_cell_literal(f, rx, ry, m, -1, -1))
        f.x[:] = rng.uniform(0, 752, n).astype(np.float32)      # in place: same pointers, new coordinates
        if trial == 3:
            f = oracle.MatchFrame(f.x[::-1].copy(), y, octv, np.zeros(n, np.float32), None, d, oracle.om_grid(0, 752, 0, 480, 32, 24))


def test_mutual_projection_matcher_against_numpy(oracle):
    rng = np.random.default_rng(2)
    n = 300
    x = rng.uniform(20, 730, n).astype(np.float32); y = rng.uniform(20, 460, n).astype(np.float32)
    octv = rng.integers(0, 4, n).astype(np.int32); d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    x2 = (x + 2)[perm]; y2 = (y - 1)[perm]; o2 = octv[perm]; d2 = d[perm].copy(); d2[:, 5] ^= rng.integers(0, 16, n).astype(np.uint8)
    g = oracle.om_grid(0, 752, 0, 480)
    f1 = oracle.MatchFrame(x, y, octv, np.zeros(n, np.float32), None, d, g)
    f2 = oracle.MatchFrame(x2, y2, o2, np.zeros(n, np.float32), None, d2, g)
    sf = oracle.scale_factors(1.2, 8)
    r12 = np.stack([x + 2, y - 1], 1).astype(np.float32); r21 = np.stack([x2 - 2, y2 + 1], 1).astype(np.float32)
    u1 = (rng.random(n) < 0.9).astype(np.uint8); u2 = (rng.random(n) < 0.9).astype(np.uint8)
    num, m = oracle.projection_match_keyframes_mutually(f1, f2, sf, u1, r12, octv, d, u2, r21, o2, d2, 6.0)

    def best(fx, fy, fo, fd, qxy, ql, qd, usable):
        out = np.full(len(ql), -1)
        for q in range(len(ql)):
            if not usable[q]:
                continue
            mg = np.float32(6.0) * sf[ql[q]]
            cand = np.flatnonzero((np.abs(fx - qxy[q, 0]) < mg) & (np.abs(fy - qxy[q, 1]) < mg) & (fo >= ql[q] - 1) & (fo <= ql[q]))
            if len(cand) == 0:
                continue
            dist = np.array([_ham(qd[q], fd[c]) for c in cand])
            if dist.min() <= 100 and (dist == dist.min()).sum() == 1:        # unique minimum: independent of the visiting order
                out[q] = cand[dist.argmin()]
            elif dist.min() <= 100:
                out[q] = -2                                                   # tie: order dependent, skip in the comparison
        return out
    b21 = best(x2, y2, o2, d2, r12, octv, d, u1); b12 = best(x, y, octv, d, r21, o2, d2, u2)
    checked = 0
    for i in range(n):
        if b21[i] == -2 or (b21[i] >= 0 and b12[b21[i]] == -2):
            continue
        want = b21[i] if (b21[i] >= 0 and b12[b21[i]] == i) else -1
        assert m[i] == want
        checked += 1
    assert checked > 0.9 * n and num == (m >= 0).sum() and num > 150


def test_triangulation_matcher_invariants(oracle):
    p = synth.triangulation_problem(400, 5)
    sf = oracle.scale_factors(1.2, 8)
    keys = ("desc_1", "bearing_1", "octave_1", "angle_1", "has_lm_1", "is_stereo_1", "bow_node_1",
            "desc_2", "bearing_2", "angle_2", "has_lm_2", "is_stereo_2", "bow_node_2", "E_12", "epipole_in_2")
    num, m = oracle.robust_match_for_triangulation(*[p[k] for k in keys], sf, False)
    hit = np.flatnonzero(m >= 0)
    assert num == len(hit) and len(np.unique(m[hit])) == len(hit)           # a keyframe-2 keypoint is given away once
    for i in hit:
        j = m[i]
        assert p["bow_node_1"][i] == p["bow_node_2"][j] and not p["has_lm_1"][i] and not p["has_lm_2"][j]
        assert _ham(p["desc_1"][i], p["desc_2"][j]) <= 50
        ep = p["E_12"] @ p["bearing_2"][j]
        res = np.pi / 2 - abs(np.arccos(ep @ p["bearing_1"][i] / np.linalg.norm(ep)))
        assert res < 0.2 * np.pi / 180 * sf[p["octave_1"][i]]
    # with the orientation check on, only matches are removed, never added or changed
    num2, m2 = oracle.robust_match_for_triangulation(*[p[k] for k in keys], sf, True)
    assert num2 <= num and ((m2 == m) | (m2 == -1)).all()
    assert (m[hit] == p["truth_idx_2_of_1"][hit]).mean() > 0.95


def test_bow_tree_matchers_against_numpy(oracle):
    """match::bow_tree (oracle only so far): node-guided nearest / second-nearest search with the ratio test
    and the first-taker rule, against a plain Python restatement."""
    p = synth.triangulation_problem(300, 9, n_nodes=12)
    d1, d2 = p["desc_1"], p["desc_2"]; n1, n2 = len(d1), len(d2)
    v1 = p["has_lm_1"] ^ 1; v2 = p["has_lm_2"] ^ 1                 # "has a valid landmark" flags for this matcher
    node1, node2 = p["bow_node_1"], p["bow_node_2"]
    D = (np.unpackbits(d1[:, None, :] ^ d2[None, :, :], axis=2).sum(2)).astype(np.int64)

    def ref(frame_mode, ratio):
        out = np.full(n2 if frame_mode else n1, -1)
        taken = np.zeros(n2, bool)
        for node in range(max(node1.max(), node2.max()) + 1):
            for i in np.flatnonzero(node1 == node):
                if not v1[i]:
                    continue
                best = second = 256; bj = -1
                for j in np.flatnonzero(node2 == node):
                    if taken[j] or (not frame_mode and not v2[j]):
                        continue
                    d = D[i, j]
                    if d < best:
                        second, best, bj = best, d, j
                    elif d < second:
                        second = d
                if best > 50 or np.float32(ratio) * np.float32(second) < np.float32(best):
                    continue
                taken[bj] = True
                if frame_mode:
                    out[bj] = i
                else:
                    out[i] = bj
        return out
    for ratio in (0.6, 0.75, 0.9):
        n, m = oracle.bow_tree_match_frame_and_keyframe(d1, p["angle_1"], v1, node1, d2, p["angle_2"], node2, ratio, False)
        want = ref(True, ratio)
        assert np.array_equal(m, want) and n == (want >= 0).sum()
        n, m = oracle.bow_tree_match_keyframes(d1, p["angle_1"], v1, node1, d2, p["angle_2"], v2, node2, ratio, False)
        want = ref(False, ratio)
        assert np.array_equal(m, want) and n == (want >= 0).sum() and n > 50
    # orientation check only removes matches
    n0, m0 = oracle.bow_tree_match_keyframes(d1, p["angle_1"], v1, node1, d2, p["angle_2"], v2, node2, 0.75, False)
    n1_, m1 = oracle.bow_tree_match_keyframes(d1, p["angle_1"], v1, node1, d2, p["angle_2"], v2, node2, 0.75, True)
    assert n1_ <= n0 and ((m1 == m0) | (m1 == -1)).all()


def test_fuse_matching_core_against_numpy(oracle):
    """match::fuse (oracle only so far): window + level range + chi-square gate + nearest descriptor."""
    rng = np.random.default_rng(4)
    n = 500
    x = rng.uniform(20, 730, n).astype(np.float32); y = rng.uniform(20, 460, n).astype(np.float32)
    octv = rng.integers(0, 5, n).astype(np.int32); d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    f = oracle.MatchFrame(x, y, octv, np.zeros(n, np.float32), None, d, oracle.om_grid(0, 752, 0, 480))
    sf = oracle.scale_factors(1.2, 8); inv_sig = (1.0 / (sf * sf)).astype(np.float32)
    nq = 400
    sel = rng.integers(0, n, nq)
    reproj = np.stack([x[sel] + rng.normal(0, 1.5, nq), y[sel] + rng.normal(0, 1.5, nq)], 1).astype(np.float32)
    lvl = np.clip(octv[sel] + rng.integers(0, 2, nq), 0, 7).astype(np.int32)
    qd = d[sel].copy(); qd[:, 7] ^= rng.integers(0, 64, nq).astype(np.uint8)
    usable = (rng.random(nq) < 0.9).astype(np.uint8)
    num, m = oracle.fuse_best_keypoints(f, reproj, None, lvl, qd, sf, inv_sig, 3.0, usable)
    checked = 0
    for q in range(nq):
        if not usable[q]:
            assert m[q] == -1
            continue
        mg = np.float32(3.0) * sf[lvl[q]]
        ex = reproj[q, 0] - x; ey = reproj[q, 1] - y
        inside = (np.abs(ex) < mg) & (np.abs(ey) < mg) & (octv >= lvl[q] - 1) & (octv <= lvl[q])
        gate = (ex * ex + ey * ey).astype(np.float32) * inv_sig[octv] <= np.float32(5.99)
        cand = np.flatnonzero(inside & gate)
        if len(cand) == 0:
            assert m[q] == -1
            continue
        dist = np.array([_ham(qd[q], d[c]) for c in cand])
        if (dist == dist.min()).sum() > 1:
            continue                                          # tie: visiting-order dependent
        want = cand[dist.argmin()] if dist.min() <= 50 else -1
        assert m[q] == want
        checked += 1
    assert checked > 300 and num == (m >= 0).sum() and num > 200
