#!/usr/bin/env python
"""Stage-by-stage GPU-vs-oracle diagnostics (never stops at the first mismatch); writes
gpurun_out/check.json.  Development aid: the authoritative checks are tests/ -m gpu."""
import json, os, sys, time, traceback
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openvslam_b200 import feature, match, synth, _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = {}


def first_diff(a, b):
    d = np.argwhere(a != b)
    return [int(v) for v in d[0]] if len(d) else None


def check_extract(w, h, n, seed):
    key = "extract_%dx%d_%d" % (w, h, n)
    r = OUT[key] = {}
    img = synth.frame(w, h, seed=seed)
    ext = feature.orb_extractor(feature.orb_params(max_num_keypts=n))
    P = O.params(n)
    t = time.time(); kps, desc = ext.extract(img); r["first_call_ms"] = (time.time() - t) * 1e3
    ts = []
    for _ in range(5):
        t = time.time(); kps, desc = ext.extract(img); ts.append((time.time() - t) * 1e3)
    r["call_ms"] = ts; r["timings_us"] = ext.last_timings_us()
    t = time.time(); okps, odesc, dbg = O.extract(img, P); r["oracle_ms"] = (time.time() - t) * 1e3
    r["n_gpu"], r["n_oracle"], r["dbg"] = len(kps), len(okps), dbg
    levels = O.build_pyramid(img, P); sf = O.scale_factors(1.2, 8)
    r["levels"] = []
    for l in range(8):
        e = {}
        g = ext.image_pyramid(l)
        e["pyr_bad"] = int((g != levels[l]).sum()); e["pyr_first"] = first_diff(g, levels[l])
        ref = O.fast_score_map(levels[l]); ref[ref < 7] = 0
        s = ext.debug_score_map(l)
        e["score_bad"] = int((s != ref).sum()); e["score_first"] = first_diff(s, ref); e["score_nonzero"] = int((ref > 0).sum())
        if e["score_first"]:
            y, x = e["score_first"]; e["score_vals"] = [int(s[y, x]), int(ref[y, x])]
        c = O.level_candidates(P, levels[l], float(sf[l]))
        cref = np.stack([c["x"], c["y"], c["score"]], 1).reshape(-1, 3)
        got = ext.debug_candidates(l)
        e["cand_n"] = [len(got), len(cref)]
        m = min(len(got), len(cref))
        e["cand_first_bad"] = first_diff(got[:m], cref[:m])
        if e["cand_first_bad"]:
            i = e["cand_first_bad"][0]; e["cand_vals"] = [got[i].tolist(), cref[i].tolist()]
        r["levels"].append(e)
    if len(kps) == len(okps):
        for f in ("x", "y", "size", "angle", "response", "octave"):
            bad = np.flatnonzero(kps[f] != okps[f])
            r["kp_bad_" + f] = int(len(bad))
            if len(bad):
                i = int(bad[0]); r["kp_first_" + f] = [i, float(kps[f][i]), float(okps[f][i]), int(okps["lx"][i]), int(okps["ly"][i]), int(okps["octave"][i])]
        badd = np.flatnonzero((desc != odesc).any(1))
        r["desc_bad"] = int(len(badd))
        if len(badd):
            i = int(badd[0]); r["desc_first"] = [i, int(np.unpackbits(desc[i] ^ odesc[i]).sum()), float(okps["angle"][i]), int(okps["octave"][i])]
    ext.close()
    return kps, desc


def check_match(n, seed):
    key = "match_%d" % n
    r = OUT[key] = {}
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    t = q[rng.permutation(n)].copy()
    t[:, :3] ^= rng.integers(0, 256, (n, 3), dtype=np.uint8)
    mt = match.robust(lowe_ratio=0.6)
    bi, bd, sd = mt.brute_force_nearest(q, t)
    ts = []
    for _ in range(5):
        t0 = time.time(); mt.brute_force_nearest(q, t); ts.append((time.time() - t0) * 1e3)
    r["call_ms"] = ts; r["kernel_us"] = mt.last_kernel_us()
    t0 = time.time(); obi, obd, osd = O.bruteforce(q, t); r["oracle_ms"] = (time.time() - t0) * 1e3
    r["bad_idx"] = int((bi != obi).sum()); r["bad_dist"] = int((bd != obd).sum()); r["bad_second"] = int((sd != osd).sum())
    got = mt.brute_force_match(q, t); ref = O.robust_brute_force_match(q, t, None, 0.6)
    r["robust_n"] = [len(got), len(ref)]; r["robust_equal"] = bool(np.array_equal(got, ref))
    mt.close()


def check_pose(n, model, stereo, seed):
    from openvslam_b200 import optimize
    key = "pose_%s_%d" % (model, n)
    r = OUT[key] = {}
    p = synth.pose_problem(n, model=model, seed=seed, stereo=stereo)
    xr = p["obs_xr"] if stereo else None
    po = optimize.pose_optimizer()
    ninl, pose, flags, st = po.optimize(optimize.camera(**p["cam"]), not stereo, p["pts_w"], p["obs_xy"], xr, p["inv_sigma_sq"], p["poses"][0])
    ts = []
    for _ in range(3):
        t0 = time.time(); po.optimize(optimize.camera(**p["cam"]), not stereo, p["pts_w"], p["obs_xy"], xr, p["inv_sigma_sq"], p["poses"][0]); ts.append((time.time() - t0) * 1e3)
    t0 = time.time()
    on, opose, oflags, ost = O.pose_optimize(O.camera(**p["cam"]), not stereo, p["pts_w"], p["obs_xy"], xr, p["inv_sigma_sq"], p["poses"][0])
    r["oracle_ms"] = (time.time() - t0) * 1e3
    r.update(call_ms=ts, ninl=[ninl, on], flags_diff=int((flags != oflags).sum()), pose_maxdiff=float(np.abs(pose - opose).max()),
             stats=st, ostats=ost, pose_err_gt=float(np.abs(pose - p["poses_gt"][0]).max()))
    po.close()


def check_ba(kf, kx, nl, model, stereo, seed):
    from openvslam_b200 import optimize
    key = "ba_%s_%d_%d" % (model, kf, nl)
    r = OUT[key] = {}
    p = synth.ba_problem(kf, kx, nl, model=model, seed=seed, stereo=stereo)
    xr = p["obs_xr"] if stereo else None
    ba = optimize.local_bundle_adjuster()
    args = (p["poses"], p["fixed"], p["points"], p["obs_kf"], p["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"])
    t0 = time.time(); poses, points, outl, st = ba.optimize(optimize.camera(**p["cam"]), not stereo, *args); r["first_call_ms"] = (time.time() - t0) * 1e3
    ts = []
    for _ in range(2):
        t0 = time.time(); ba.optimize(optimize.camera(**p["cam"]), not stereo, *args); ts.append((time.time() - t0) * 1e3)
    t0 = time.time(); oposes, opoints, ooutl, ost = O.local_ba(O.camera(**p["cam"]), not stereo, *args); r["oracle_ms"] = (time.time() - t0) * 1e3
    chi = lambda P_, X_, m: synth.reprojection_chi2(p["cam"], P_, X_, p["obs_kf"], p["obs_lm"], p["obs_xy"], xr, p["inv_sigma_sq"], m)
    c, oc, c0 = chi(poses, points, ~outl), chi(oposes, opoints, ~ooutl), chi(p["poses"], p["points"], None)
    r.update(M=len(p["obs_kf"]), call_ms=ts, stats=st, ostats=ost, outl_diff=int((outl != ooutl).sum()), n_outl=[int(outl.sum()), int(ooutl.sum())],
             chi=[c, oc, c0], chi_rel=abs(c - oc) / oc, pose_maxdiff=float(np.abs(poses - oposes).max()), point_maxdiff=float(np.abs(points - opoints).max()))
    ba.close()


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if "--ba-only" not in sys.argv:
        main_extract_match()
    for args in [(400, "perspective", True, 1), (2000, "perspective", True, 2), (4000, "equirectangular", False, 3)]:
        try:
            check_pose(*args)
        except Exception:
            OUT["pose_%s_%d_error" % (args[1], args[0])] = traceback.format_exc()
    for args in [(6, 2, 300, "equirectangular", False, 6), (10, 4, 1500, "perspective", True, 2), (50, 10, 20000, "equirectangular", False, 4)]:
        try:
            check_ba(*args)
        except Exception:
            OUT["ba_%s_%d_error" % (args[3], args[0])] = traceback.format_exc()
    finish()


def main_extract_match():
    for args in [(640, 480, 1000, 3), (1920, 960, 4000, 4), (333, 131, 300, 5)]:
        try:
            check_extract(*args)
        except Exception:
            OUT["extract_%dx%d_%d_error" % args[:3]] = traceback.format_exc()
    for n in (1000, 4000):
        try:
            check_match(n, n)
        except Exception:
            OUT["match_%d_error" % n] = traceback.format_exc()


def finish():
    OUT["launches"] = _lib.launch_count()
    with open(os.path.join(ROOT, "gpurun_out", "check.json"), "w") as f:
        json.dump(OUT, f, indent=1, default=str)
    # compact summary on stdout
    for k, v in OUT.items():
        if isinstance(v, dict) and "levels" in v:
            print(k, "n", v["n_gpu"], v["n_oracle"], "pyr_bad", [e["pyr_bad"] for e in v["levels"]], "score_bad", [e["score_bad"] for e in v["levels"]],
                  "cand", [e["cand_n"] for e in v["levels"]], "cand_bad", [e["cand_first_bad"] for e in v["levels"]],
                  {kk: vv for kk, vv in v.items() if kk.startswith(("kp_bad", "desc_bad"))}, "timings", v["timings_us"], "call_ms", v["call_ms"], "oracle_ms", v["oracle_ms"])
        else:
            print(k, v)


if __name__ == "__main__":
    main()
