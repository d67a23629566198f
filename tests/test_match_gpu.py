"""GPU parity of the Hamming brute-force matcher (match::robust::brute_force_match) against the
oracle: distances, ranks (lowest index wins ties), second-best, and the greedy uniqueness rule."""
import numpy as np
import pytest

from openvslam_b200 import synth

pytestmark = pytest.mark.gpu


def _rand_desc(rng, n):
    return rng.integers(0, 256, (n, 32), dtype=np.uint8)


def _noisy_copy(rng, d, flips):
    out = d.copy()
    for i in range(len(out)):
        bits = rng.choice(256, flips, replace=False)
        for b in bits:
            out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


@pytest.mark.parametrize("n1,n2", [(1, 1), (3, 2), (5, 700), (1000, 1000), (2000, 2000), (4000, 4000), (4007, 3991), (257, 65535)])
def test_topk_ranks_bit_exact(oracle, n1, n2):
    from openvslam_b200 import match
    rng = np.random.default_rng(n1 * 7 + n2)
    q = _rand_desc(rng, n1); t = _rand_desc(rng, n2)
    m = min(n1, n2, 300)
    t[rng.choice(n2, m, replace=False)] = _noisy_copy(rng, q[rng.choice(n1, m, replace=False)], 20)
    mt = match.robust()
    bi, bd, sd = mt.brute_force_nearest(q, t)
    obi, obd, osd = oracle.bruteforce(q, t)
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and np.array_equal(sd, osd)
    keys = mt.brute_force_topk(q, t)
    assert (np.diff(keys.astype(np.int64), axis=1) > 0).all() or n2 < 8
    # checksum-of-distances property at full size: sum of best distances equals the oracle's
    assert int(bd.sum()) == int(obd.sum())
    mt.close()


def test_topk_ties_prefer_lowest_index(oracle):
    from openvslam_b200 import match
    rng = np.random.default_rng(5)
    q = _rand_desc(rng, 64)
    t = np.concatenate([q, q, q])[rng.permutation(192)]  # every query has 3 exact duplicates
    mt = match.robust()
    bi, bd, sd = mt.brute_force_nearest(q, t)
    obi, obd, osd = oracle.bruteforce(q, t)
    assert np.array_equal(bi, obi) and (bd == 0).all() and (sd == 0).all() and np.array_equal(sd, osd)
    mt.close()


def test_empty_and_ragged(oracle):
    from openvslam_b200 import match
    rng = np.random.default_rng(6)
    mt = match.robust()
    q = _rand_desc(rng, 10)
    bi, bd, sd = mt.brute_force_nearest(q, np.zeros((0, 32), np.uint8))
    assert (bi == -1).all() and (bd == 256).all() and (sd == 256).all()
    assert len(mt.brute_force_match(np.zeros((0, 32), np.uint8), q)) == 0
    assert len(mt.brute_force_match(q, np.zeros((0, 32), np.uint8))) == 0
    bi, bd, sd = mt.brute_force_nearest(q, q[:1])
    assert (bi == 0).all() and (sd == 256).all()
    mt.close()


@pytest.mark.parametrize("n,flips,seed", [(500, 10, 1), (2000, 25, 2), (4000, 30, 3)])
def test_robust_brute_force_match_equals_oracle(oracle, n, flips, seed):
    from openvslam_b200 import match
    rng = np.random.default_rng(seed)
    frm = _rand_desc(rng, n)
    kf = _noisy_copy(rng, frm[rng.permutation(n)], flips)
    kf[::7] = _rand_desc(rng, len(kf[::7]))         # unmatched keyframe keypoints
    valid = (rng.random(n) < 0.8).astype(np.uint8)   # keyframe keypoints with a valid landmark
    mt = match.robust(lowe_ratio=0.6)
    got = mt.brute_force_match(frm, kf, valid)
    ref = oracle.robust_brute_force_match(frm, kf, valid, 0.6)
    assert np.array_equal(got, ref) and len(got) > n // 3
    mt.close()


def test_robust_greedy_uniqueness_requery(oracle):
    """Many keyframe descriptors competing for the same frame descriptors: the reference removes
    a matched frame keypoint from later scans, so later best/second-best change.  Exercises the
    GPU re-query path (top-4 list exhausted by claimed indices)."""
    from openvslam_b200 import match
    rng = np.random.default_rng(9)
    base = _rand_desc(rng, 40)
    frm = np.concatenate([_noisy_copy(rng, base, 3) for _ in range(8)])   # 8 near-duplicates of each
    kf = np.concatenate([_noisy_copy(rng, base, 2) for _ in range(12)])   # 12 competitors for each
    for ratio in (0.6, 0.9, 1.0):
        mt = match.robust(lowe_ratio=ratio)
        got = mt.brute_force_match(frm, kf)
        ref = oracle.robust_brute_force_match(frm, kf, None, ratio)
        assert np.array_equal(got, ref), ratio
        mt.close()


def test_match_on_real_descriptors(oracle):
    """Extract two shifted frames, brute-force match them, compare with the oracle end to end."""
    from openvslam_b200 import feature, match
    a = synth.frame(752, 480, seed=70)
    b = synth.shifted(a, 3, 2)
    ext = feature.orb_extractor(feature.orb_params(max_num_keypts=1000))
    ka, da = ext.extract(a)
    kb, db = ext.extract(b)
    mt = match.robust(lowe_ratio=0.75)
    got = mt.brute_force_match(da, db)
    ref = oracle.robust_brute_force_match(da, db, None, 0.75)
    assert np.array_equal(got, ref)
    # ground truth: matched keypoints are displaced by the known shift (mostly)
    dx = kb["x"][got[:, 1]] - ka["x"][got[:, 0]]
    assert np.mean(np.abs(dx - 3) < 2.5) > 0.8
    ext.close(); mt.close()


@pytest.mark.parametrize("n,seed,check", [(1500, 1, True), (600, 2, False), (3000, 3, True)])
def test_match_for_triangulation(oracle, n, seed, check):
    """robust::match_for_triangulation: BoW-node-guided candidates, epipole + epipolar-plane tests, first-taker rule,
    last-of-equal-distances tie rule, orientation histogram -- bit-exact against the oracle."""
    from openvslam_b200 import match
    p = synth.triangulation_problem(n, seed)
    sf = oracle.scale_factors(1.2, 8)
    keys = ("desc_1", "bearing_1", "octave_1", "angle_1", "has_lm_1", "is_stereo_1", "bow_node_1",
            "desc_2", "bearing_2", "angle_2", "has_lm_2", "is_stereo_2", "bow_node_2", "E_12", "epipole_in_2")
    mt = match.robust(check_orientation=check)
    num, m = mt.match_for_triangulation(*[p[k] for k in keys], sf)
    onum, om = oracle.robust_match_for_triangulation(*[p[k] for k in keys], sf, check)
    assert num == onum and np.array_equal(m, om)
    good = (m >= 0) & (m == p["truth_idx_2_of_1"])
    assert num > 0.25 * n and good.sum() > 0.9 * num          # mostly the true correspondences
    mt.close()


def test_match_for_triangulation_tiny_nodes_exhaust_the_candidate_lists(oracle):
    """Two nodes only: segments of hundreds of keypoints, many equal descriptors -> the 8-entry lists run out and the
    GPU re-query (same kernel, claimed keypoints excluded) has to reproduce the sequential loop."""
    from openvslam_b200 import match
    p = synth.triangulation_problem(500, 7, n_nodes=2)
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (3, 32), dtype=np.uint8)
    p["desc_1"] = base[rng.integers(0, 3, len(p["desc_1"]))]
    p["desc_2"] = base[rng.integers(0, 3, len(p["desc_2"]))]
    # every pair passes the geometric tests: with E = e_z e_z' the residual is pi/2 - acos(b1z * sign(b2z)), negative when
    # the two z components have opposite signs (the one-sided test of check_epipolar_constraint), and the epipole is behind
    p["E_12"] = np.zeros((3, 3)); p["E_12"][2, 2] = 1.0
    for k, z in (("bearing_1", -0.5), ("bearing_2", 0.5)):
        b = p[k].copy(); b[:, 2] = z * np.linalg.norm(b[:, :2], axis=1); b /= np.linalg.norm(b, axis=1, keepdims=True); p[k] = b
    p["epipole_in_2"] = np.array([0.0, 0.0, -1.0])
    sf = oracle.scale_factors(1.2, 8)
    keys = ("desc_1", "bearing_1", "octave_1", "angle_1", "has_lm_1", "is_stereo_1", "bow_node_1",
            "desc_2", "bearing_2", "angle_2", "has_lm_2", "is_stereo_2", "bow_node_2", "E_12", "epipole_in_2")
    mt = match.robust(check_orientation=False)
    before = mt.num_requeries()
    num, m = mt.match_for_triangulation(*[p[k] for k in keys], sf)
    onum, om = oracle.robust_match_for_triangulation(*[p[k] for k in keys], sf, False)
    assert num == onum and np.array_equal(m, om) and num > 100
    assert mt.num_requeries() > before
    mt.close()


def test_brute_force_match_on_device_resident_descriptors(oracle):
    """ovs_robust_brute_force_match_device: descriptors stay in HBM, the greedy replay is the same code as the host entry."""
    import torch
    from openvslam_b200 import match
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (3000, 32), dtype=np.uint8)
    d1 = base.copy(); d2 = base[rng.permutation(3000)[:2500]].copy()
    flip = rng.integers(0, 256, d2.shape, dtype=np.uint8) & rng.integers(0, 256, d2.shape, dtype=np.uint8) & rng.integers(0, 256, d2.shape, dtype=np.uint8) & rng.integers(0, 256, d2.shape, dtype=np.uint8)
    d2 ^= flip
    valid = (rng.random(len(d2)) < 0.9).astype(np.uint8)
    dev = torch.device("cuda", 0)
    t1 = torch.from_numpy(d1).to(dev); t2 = torch.from_numpy(d2).to(dev)
    mt = match.robust(lowe_ratio=0.75)
    got = mt.brute_force_match_device(t1.data_ptr(), len(d1), t2.data_ptr(), len(d2), valid)
    host = mt.brute_force_match(d1, d2, valid)
    ref = oracle.robust_brute_force_match(d1, d2, valid, 0.75)
    assert np.array_equal(got, ref) and np.array_equal(host, ref) and len(ref) > 1000
    mt.close()


def _bow_problem(n, seed, n_nodes, dup=False):
    p = synth.triangulation_problem(n, seed, n_nodes=n_nodes)
    rng = np.random.default_rng(seed + 100)
    if dup:   # many equal descriptors: ties everywhere, lists run out
        base = rng.integers(0, 256, (4, 32), dtype=np.uint8)
        p["desc_1"] = base[rng.integers(0, 4, len(p["desc_1"]))]
        p["desc_2"] = base[rng.integers(0, 4, len(p["desc_2"]))]
        p["desc_2"][:, 3] ^= (rng.random(len(p["desc_2"])) < 0.5).astype(np.uint8)
    v1 = (rng.random(len(p["desc_1"])) < 0.8).astype(np.uint8); v2 = (rng.random(len(p["desc_2"])) < 0.8).astype(np.uint8)
    p["bow_node_1"][rng.random(len(v1)) < 0.03] = -1; p["bow_node_2"][rng.random(len(v2)) < 0.03] = -1
    return p, v1, v2


@pytest.mark.parametrize("n,seed,nodes,dup,orient", [(1500, 1, 40, False, True), (3000, 2, 60, False, False), (3000, 3, 3, False, True), (600, 4, 2, True, False)])
def test_bow_tree_matchers_match_the_oracle(oracle, n, seed, nodes, dup, orient):
    """match::bow_tree::{match_keyframes, match_frame_and_keyframe}: node-guided nearest / second nearest with the first-taker
    rule, ratio test and orientation histogram; tiny vocabularies and duplicated descriptors force GPU re-queries."""
    from openvslam_b200 import match
    p, v1, v2 = _bow_problem(n, seed, nodes, dup)
    mt = match.bow_tree(lowe_ratio=0.75 if dup else 0.6, check_orientation=orient)
    before = mt.num_requeries()
    num, m = mt.match_keyframes(p["desc_1"], p["angle_1"], v1, p["bow_node_1"], p["desc_2"], p["angle_2"], v2, p["bow_node_2"])
    onum, om = oracle.bow_tree_match_keyframes(p["desc_1"], p["angle_1"], v1, p["bow_node_1"], p["desc_2"], p["angle_2"], v2, p["bow_node_2"],
                                               mt.lowe_ratio_, orient)
    assert num == onum and np.array_equal(m, om)
    num2, m2 = mt.match_frame_and_keyframe(p["desc_1"], p["angle_1"], v1, p["bow_node_1"], p["desc_2"], p["angle_2"], p["bow_node_2"])
    onum2, om2 = oracle.bow_tree_match_frame_and_keyframe(p["desc_1"], p["angle_1"], v1, p["bow_node_1"], p["desc_2"], p["angle_2"], p["bow_node_2"],
                                                          mt.lowe_ratio_, orient)
    assert num2 == onum2 and np.array_equal(m2, om2)
    if not dup:
        assert num > n // 10
    if dup:
        assert mt.num_requeries() > before
    mt.close()


def test_bow_tree_empty_inputs():
    from openvslam_b200 import match
    mt = match.bow_tree()
    z8 = np.zeros((0, 32), np.uint8); zf = np.zeros(0, np.float32); zi = np.zeros(0, np.int32); zu = np.zeros(0, np.uint8)
    d = np.zeros((5, 32), np.uint8); a = np.zeros(5, np.float32); nd = np.zeros(5, np.int32); v = np.ones(5, np.uint8)
    assert mt.match_keyframes(z8, zf, zu, zi, d, a, v, nd)[0] == 0
    n, m = mt.match_keyframes(d, a, v, nd, z8, zf, zu, zi)
    assert n == 0 and (m == -1).all()
    n, m = mt.match_frame_and_keyframe(d, a, v, -np.ones(5, np.int32), d, a, nd)     # no keyframe keypoint has a node
    assert n == 0 and (m == -1).all()
    mt.close()
