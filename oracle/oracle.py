"""ctypes front for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs.  The product package (openvslam_b200/) never imports this module."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
MAX_LEVELS = 16


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h", ".inc"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


class FastPt(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("score", C.c_int)]


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int), ("lx", C.c_int), ("ly", C.c_int)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("lx", "<i4"), ("ly", "<i4")])
FASTPT_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("score", "<i4")])


class Params(C.Structure):
    _fields_ = [("max_num_keypts", C.c_uint32), ("scale_factor", C.c_float), ("num_levels", C.c_uint32),
                ("ini_fast_thr", C.c_uint32), ("min_fast_thr", C.c_uint32)]


class Debug(C.Structure):
    _fields_ = [("level_w", C.c_int * MAX_LEVELS), ("level_h", C.c_int * MAX_LEVELS),
                ("num_candidates", C.c_int * MAX_LEVELS), ("num_selected", C.c_int * MAX_LEVELS)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oo_fast_atan2.restype = C.c_float
        _lib.oo_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.oo_ic_angle.restype = C.c_float
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def params(max_num_keypts=2000, scale_factor=1.2, num_levels=8, ini_fast_thr=20, min_fast_thr=7):
    return Params(max_num_keypts, scale_factor, num_levels, ini_fast_thr, min_fast_thr)


def scale_factors(scale_factor, num_levels):
    out = np.zeros(num_levels, np.float32)
    lib().oo_scale_factors(C.c_float(scale_factor), num_levels, out.ctypes.data_as(C.c_void_p))
    return out


def level_sizes(w, h, scale_factor, num_levels):
    sf = scale_factors(scale_factor, num_levels)
    out = [(w, h)]
    for l in range(1, num_levels):
        lw, lh = C.c_int(), C.c_int()
        lib().oo_level_size(w, h, C.c_float(float(sf[l])), C.byref(lw), C.byref(lh))
        out.append((lw.value, lh.value))
    return out


def keypts_per_level(max_num_keypts, scale_factor, num_levels):
    out = np.zeros(num_levels, np.uint32)
    lib().oo_keypts_per_level(max_num_keypts, C.c_float(scale_factor), num_levels, out.ctypes.data_as(C.c_void_p))
    return out


def resize_linear(src, dw, dh):
    src, p = _u8(src)
    dst = np.empty((dh, dw), np.uint8)
    rc = lib().oo_resize_linear_u8(p, src.shape[1], src.shape[0], src.strides[0], dst.ctypes.data_as(C.c_void_p), dw, dh, dw)
    assert rc == 0
    return dst


def fast_detect(img, threshold, nonmax=True):
    img, p = _u8(img)
    h, w = img.shape
    out = np.zeros(w * h, FASTPT_DTYPE)
    n = lib().oo_fast_detect(p, w, h, img.strides[0], int(threshold), int(nonmax), out.ctypes.data_as(C.c_void_p), out.size)
    return out[:n].copy()


def fast_score_map(img):
    img, p = _u8(img)
    h, w = img.shape
    out = np.zeros((h, w), np.uint8)
    lib().oo_fast_score_map(p, w, h, img.strides[0], out.ctypes.data_as(C.c_void_p), w)
    return out


def distribute_via_tree(cand, min_x, max_x, min_y, max_y, num_keypts):
    cand = np.ascontiguousarray(cand, dtype=FASTPT_DTYPE)
    out = np.zeros(len(cand) + 1, np.int32)
    n = lib().oo_distribute_via_tree(cand.ctypes.data_as(C.c_void_p), len(cand), min_x, max_x, min_y, max_y,
                                     C.c_uint(num_keypts), out.ctypes.data_as(C.c_void_p))
    return out[:n].copy()


def fast_atan2(y, x):
    return lib().oo_fast_atan2(C.c_float(y), C.c_float(x))


def ic_angle(img, x, y):
    img, p = _u8(img)
    m01, m10 = C.c_int(), C.c_int()
    a = lib().oo_ic_angle(p, img.strides[0], int(x), int(y), C.byref(m01), C.byref(m10))
    return a, m01.value, m10.value


def gaussian7(img):
    img, p = _u8(img)
    h, w = img.shape
    out = np.empty((h, w), np.uint8)
    lib().oo_gaussian7(p, w, h, img.strides[0], out.ctypes.data_as(C.c_void_p), w)
    return out


def sincosf(a):
    s, c = C.c_float(), C.c_float()
    lib().oo_sincosf(C.c_float(a), C.byref(s), C.byref(c))
    return s.value, c.value


def orb_descriptor(blurred, x, y, angle_deg):
    blurred, p = _u8(blurred)
    d = np.zeros(32, np.uint8)
    lib().oo_orb_descriptor(p, blurred.strides[0], int(x), int(y), C.c_float(angle_deg), d.ctypes.data_as(C.c_void_p))
    return d


def rect_mask(cols, rows, rects):
    r = np.ascontiguousarray(rects, np.float32).reshape(-1, 4)
    m = np.empty((rows, cols), np.uint8)
    lib().oo_rect_mask(cols, rows, r.ctypes.data_as(C.c_void_p), len(r), m.ctypes.data_as(C.c_void_p))
    return m


def build_pyramid(img, P):
    img, p = _u8(img)
    h, w = img.shape
    sizes = level_sizes(w, h, P.scale_factor, P.num_levels)
    levels = [np.empty((lh, lw), np.uint8) for lw, lh in sizes]
    arr = (C.c_void_p * len(levels))(*[l.ctypes.data_as(C.c_void_p) for l in levels])
    lib().oo_build_pyramid(C.byref(P), p, w, h, img.strides[0], arr)
    return levels


def extract(img, P, mask=None, with_desc=True, max_out=None):
    """Returns (keypoints[KP_DTYPE], descriptors[N,32] u8, debug dict)."""
    img, p = _u8(img)
    h, w = img.shape
    # the tree's first pass splits every initial node (round(aspect ratio) of them): a very wide image returns far more than the budget
    aspect = max((w - 38) / max(h - 38, 1), (h - 38) / max(w - 38, 1), 1.0)
    max_out = max_out or (int(P.max_num_keypts) * 2 + 64 + int(P.num_levels) * 4 * (int(round(aspect)) + 1))
    kps = np.zeros(max_out, KP_DTYPE)
    desc = np.zeros((max_out, 32), np.uint8)
    dbg = Debug()
    if mask is not None:
        mask, mp = _u8(mask)
        assert mask.shape == img.shape
        ms = mask.strides[0]
    else:
        mp, ms = None, 0
    n = lib().oo_extract(C.byref(P), p, w, h, img.strides[0], mp, ms, kps.ctypes.data_as(C.c_void_p),
                         desc.ctypes.data_as(C.c_void_p) if with_desc else None, max_out, C.byref(dbg))
    assert 0 <= n <= max_out, n
    L = P.num_levels
    d = dict(level_w=list(dbg.level_w)[:L], level_h=list(dbg.level_h)[:L],
             num_candidates=list(dbg.num_candidates)[:L], num_selected=list(dbg.num_selected)[:L])
    return kps[:n].copy(), desc[:n].copy(), d


# ---------------------------------------------------------------------------- matchers
def hamming(a, b):
    a, pa = _u8(a); b, pb = _u8(b)
    return int(lib().om_hamming(pa, pb))


def bruteforce(desc1, desc2):
    d1, p1 = _u8(desc1); d2, p2 = _u8(desc2)
    n1, n2 = len(d1), len(d2)
    bi = np.zeros(n1, np.int32); bd = np.zeros(n1, np.int32); sd = np.zeros(n1, np.int32)
    lib().om_bruteforce(p1, n1, p2, n2, bi.ctypes.data_as(C.c_void_p), bd.ctypes.data_as(C.c_void_p), sd.ctypes.data_as(C.c_void_p))
    return bi, bd, sd


def robust_brute_force_match(desc_frm, desc_keyfrm, lm_valid_2=None, lowe_ratio=0.6):
    d1, p1 = _u8(desc_frm); d2, p2 = _u8(desc_keyfrm)
    n1, n2 = len(d1), len(d2)
    vp = None
    if lm_valid_2 is not None:
        lm_valid_2, vp = _u8(lm_valid_2)
    pairs = np.zeros((max(min(n1, n2), 1), 2), np.int32)
    n = lib().om_robust_brute_force_match(p1, n1, p2, n2, vp, C.c_float(lowe_ratio), pairs.ctypes.data_as(C.c_void_p))
    return pairs[:n].copy()


def level_candidates(P, level_img, scale=1.0, mask=None):
    """Candidates (FASTPT_DTYPE, coordinates relative to the 19 px border) of one pyramid level in the
    reference's visiting order.  `mask`, if given, is the level-0 mask."""
    level_img, p = _u8(level_img)
    lh, lw = level_img.shape
    if mask is not None:
        mask, mp = _u8(mask); mh, mw = mask.shape; ms = mask.strides[0]
    else:
        mp, mh, mw, ms = None, 0, 0, 0
    out = C.c_void_p()
    lib().oo_level_candidates.restype = C.c_int
    n = lib().oo_level_candidates(C.byref(P), p, lw, lh, level_img.strides[0], C.c_float(scale), mp, mw, mh, ms, C.byref(out))
    res = np.zeros(n, FASTPT_DTYPE)
    if n:
        C.memmove(res.ctypes.data, out.value, n * FASTPT_DTYPE.itemsize)
    lib().oo_free(out)
    return res


# ------------------------------------------------------------------ pose optimiser / local BA
CAM_PERSPECTIVE, CAM_EQUIRECTANGULAR = 0, 1
MAX_ROUNDS = 8


class Camera(C.Structure):
    _fields_ = [("model", C.c_int), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("focal_x_baseline", C.c_double), ("cols", C.c_double), ("rows", C.c_double)]


class BaStats(C.Structure):
    _fields_ = [("num_rounds", C.c_int), ("num_iterations", C.c_int), ("num_trials", C.c_int),
                ("round_iterations", C.c_int * MAX_ROUNDS), ("lambda_init", C.c_double * MAX_ROUNDS),
                ("last_lambda", C.c_double), ("last_chi2", C.c_double), ("final_chi2", C.c_double)]


def camera(model="perspective", fx=0, fy=0, cx=0, cy=0, focal_x_baseline=0, cols=0, rows=0):
    return Camera(CAM_EQUIRECTANGULAR if model == "equirectangular" else CAM_PERSPECTIVE, fx, fy, cx, cy, focal_x_baseline, cols, rows)


def _p(a, dt):
    a = np.ascontiguousarray(a, dt)
    return a, a.ctypes.data_as(C.c_void_p)


def pose_oplus(pose, u):
    pose, pp = _p(pose, np.float64); u, pu = _p(u, np.float64)
    out = np.zeros(12)
    lib().ob_pose_oplus(pp, pu, out.ctypes.data_as(C.c_void_p))
    return out


def edge_eval(cam, pose, pw, obs, stereo):
    pose, pp = _p(pose, np.float64); pw, ppw = _p(pw, np.float64); obs, po = _p(obs, np.float64)
    e = np.zeros(3); Jp = np.zeros(18); Jl = np.zeros(9); pc = np.zeros(3)
    dim = lib().ob_edge_eval(C.byref(cam), pp, ppw, po, int(stereo), e.ctypes.data_as(C.c_void_p), Jp.ctypes.data_as(C.c_void_p),
                             Jl.ctypes.data_as(C.c_void_p), pc.ctypes.data_as(C.c_void_p))
    return e[:dim].copy(), Jp[:6 * dim].reshape(dim, 6).copy(), Jl[:3 * dim].reshape(dim, 3).copy(), pc


def _stats(st):
    return dict(num_rounds=st.num_rounds, num_iterations=st.num_iterations, num_trials=st.num_trials,
                round_iterations=list(st.round_iterations)[:st.num_rounds], lambda_init=list(st.lambda_init)[:st.num_rounds],
                last_lambda=st.last_lambda, last_chi2=st.last_chi2, final_chi2=st.final_chi2)


def pose_optimize(cam, setup_is_mono, pts_w, obs_xy, obs_xr, inv_sigma_sq, pose_cw, num_trials=4, num_each_iter=10):
    pts_w, pp = _p(pts_w, np.float64); obs_xy, po = _p(obs_xy, np.float32); inv_sigma_sq, pi = _p(inv_sigma_sq, np.float32)
    n = len(inv_sigma_sq)
    if obs_xr is not None:
        obs_xr, px = _p(obs_xr, np.float32)
    else:
        px = None
    pose = np.array(pose_cw, np.float64).reshape(12).copy()
    flags = np.zeros(n, np.uint8)
    st = BaStats()
    lib().ob_pose_optimize.restype = C.c_int
    ninl = lib().ob_pose_optimize(C.byref(cam), int(setup_is_mono), n, pp, po, px, pi, pose.ctypes.data_as(C.c_void_p),
                                  flags.ctypes.data_as(C.c_void_p), num_trials, num_each_iter, C.byref(st))
    return ninl, pose, flags.astype(bool), _stats(st)


def local_ba(cam, setup_is_mono, poses, fixed, points, obs_kf, obs_lm, obs_xy, obs_xr, inv_sigma_sq,
             num_first_iter=5, num_second_iter=10, force_stop=None):
    poses = np.array(poses, np.float64).reshape(-1, 12).copy(); points = np.array(points, np.float64).reshape(-1, 3).copy()
    fixed, pf = _p(fixed, np.uint8); obs_kf, pk = _p(obs_kf, np.int32); obs_lm, pl = _p(obs_lm, np.int32)
    obs_xy, po = _p(obs_xy, np.float32); inv_sigma_sq, pi = _p(inv_sigma_sq, np.float32)
    if obs_xr is not None:
        obs_xr, px = _p(obs_xr, np.float32)
    else:
        px = None
    M = len(obs_kf)
    out = np.zeros(M, np.uint8)
    st = BaStats()
    fs = None
    if force_stop is not None:
        fs = C.c_int(int(force_stop))
    lib().ob_local_ba(C.byref(cam), int(setup_is_mono), len(poses), poses.ctypes.data_as(C.c_void_p), pf, len(points),
                      points.ctypes.data_as(C.c_void_p), M, pk, pl, po, px, pi, num_first_iter, num_second_iter,
                      C.byref(fs) if fs is not None else None, out.ctypes.data_as(C.c_void_p), C.byref(st))
    return poses, points, out.astype(bool), _stats(st)


def global_ba(cam, setup_is_mono, poses, fixed, points, obs_kf, obs_lm, obs_xy, obs_xr, inv_sigma_sq, num_iter=10, use_huber_kernel=True,
              force_stop=None):
    """optimize::global_bundle_adjuster::optimize (8f rank 4; oracle only so far): one Levenberg round over the whole graph."""
    poses = np.array(poses, np.float64).reshape(-1, 12).copy(); points = np.array(points, np.float64).reshape(-1, 3).copy()
    fixed, pf = _p(fixed, np.uint8); obs_kf, pk = _p(obs_kf, np.int32); obs_lm, pl = _p(obs_lm, np.int32)
    obs_xy, po = _p(obs_xy, np.float32); inv_sigma_sq, pi = _p(inv_sigma_sq, np.float32)
    px = None
    if obs_xr is not None:
        obs_xr, px = _p(obs_xr, np.float32)
    st = BaStats()
    fs = C.c_int(int(force_stop)) if force_stop is not None else None
    lib().ob_global_ba(C.byref(cam), int(setup_is_mono), len(poses), poses.ctypes.data_as(C.c_void_p), pf, len(points),
                       points.ctypes.data_as(C.c_void_p), len(obs_kf), pk, pl, po, px, pi, int(num_iter), int(bool(use_huber_kernel)),
                       C.byref(fs) if fs is not None else None, C.byref(st))
    return poses, points, _stats(st)


# ------------------------------------------------------------------ windowed matchers / stereo
class OmGrid(C.Structure):
    _fields_ = [("min_x", C.c_float), ("min_y", C.c_float), ("inv_cell_width", C.c_float), ("inv_cell_height", C.c_float),
                ("num_grid_cols", C.c_int), ("num_grid_rows", C.c_int)]


class OmFrame(C.Structure):
    _fields_ = [("n", C.c_int), ("x", C.c_void_p), ("y", C.c_void_p), ("octave", C.c_void_p), ("angle", C.c_void_p),
                ("x_right", C.c_void_p), ("desc", C.c_void_p), ("grid", OmGrid)]


def om_grid(min_x, max_x, min_y, max_y, cols=64, rows=48):
    return OmGrid(min_x, min_y, np.float32(float(cols) / (max_x - min_x)), np.float32(float(rows) / (max_y - min_y)), cols, rows)


class MatchFrame:
    def __init__(self, x, y, octave, angle, x_right, desc, grid):
        self.x = np.ascontiguousarray(x, np.float32); self.y = np.ascontiguousarray(y, np.float32)
        self.octave = np.ascontiguousarray(octave, np.int32); self.angle = np.ascontiguousarray(angle, np.float32)
        self.x_right = None if x_right is None else np.ascontiguousarray(x_right, np.float32)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.n = len(self.x)
        self.c = OmFrame(self.n, self.x.ctypes.data, self.y.ctypes.data, self.octave.ctypes.data, self.angle.ctypes.data,
                         None if self.x_right is None else self.x_right.ctypes.data, self.desc.ctypes.data, grid)


def get_keypoints_in_cell(frm, ref_x, ref_y, margin, min_level, max_level):
    out = np.zeros(frm.n + 1, np.int32)
    n = lib().om_get_keypoints_in_cell(C.byref(frm.c), C.c_float(ref_x), C.c_float(ref_y), C.c_float(margin), int(min_level), int(max_level),
                                       out.ctypes.data_as(C.c_void_p))
    return out[:n].copy()


def get_keypoints_in_cell_literal(frm, ref_x, ref_y, margin, min_level, max_level):
    """The same query with the cell lists rebuilt on every call (checker of the cached version)."""
    out = np.zeros(frm.n + 1, np.int32)
    n = lib().om_get_keypoints_in_cell_literal(C.byref(frm.c), C.c_float(ref_x), C.c_float(ref_y), C.c_float(margin), int(min_level), int(max_level),
                                               out.ctypes.data_as(C.c_void_p))
    return out[:n].copy()


def projection_match_frame_and_landmarks(frm, scale_factors, reproj_xy, x_right_in_tracking, pred_scale_level, lm_desc, lm_usable=None,
                                         kp_has_observed_lm=None, margin=5.0, lowe_ratio=0.6):
    sf, psf = _p(scale_factors, np.float32); rp, prp = _p(reproj_xy, np.float32); lv, plv = _p(pred_scale_level, np.int32)
    d, pd = _p(lm_desc, np.uint8)
    pxr = pu = pk = None
    if x_right_in_tracking is not None:
        x_right_in_tracking, pxr = _p(x_right_in_tracking, np.float32)
    if lm_usable is not None:
        lm_usable, pu = _p(lm_usable, np.uint8)
    if kp_has_observed_lm is not None:
        kp_has_observed_lm, pk = _p(kp_has_observed_lm, np.uint8)
    out = np.full(max(frm.n, 1), -1, np.int32)
    n = lib().om_projection_match_frame_and_landmarks(C.byref(frm.c), psf, len(lv), pu, prp, pxr, plv, pd, pk, C.c_float(margin), C.c_float(lowe_ratio),
                                                      out.ctypes.data_as(C.c_void_p))
    return n, out[:frm.n]


def projection_match_current_and_last(curr, scale_factors, num_scale_levels, last_usable, reproj_xy, reproj_x_right, last_scale_level, last_angle,
                                      lm_desc, kp_has_observed_lm=None, margin=20.0, assume_forward=False, assume_backward=False, check_orientation=True):
    sf, psf = _p(scale_factors, np.float32); rp, prp = _p(reproj_xy, np.float32); lv, plv = _p(last_scale_level, np.int32)
    d, pd = _p(lm_desc, np.uint8); la, pla = _p(last_angle, np.float32); lu, plu = _p(last_usable, np.uint8)
    pxr = pk = None
    if reproj_x_right is not None:
        reproj_x_right, pxr = _p(reproj_x_right, np.float32)
    if kp_has_observed_lm is not None:
        kp_has_observed_lm, pk = _p(kp_has_observed_lm, np.uint8)
    out = np.full(max(curr.n, 1), -1, np.int32)
    n = lib().om_projection_match_current_and_last(C.byref(curr.c), psf, int(num_scale_levels), len(lv), plu, prp, pxr, plv, pla, pd, pk, C.c_float(margin),
                                                   int(assume_forward), int(assume_backward), int(check_orientation), out.ctypes.data_as(C.c_void_p))
    return n, out[:curr.n]


def projection_match_best(frm, ref_xy, ref_x_right, margin, min_level, max_level, q_angle, q_desc, usable=None, kp_unavailable=None,
                          hamm_dist_thr=100, check_orientation=True):
    rp, prp = _p(ref_xy, np.float32); mg, pmg = _p(margin, np.float32); lo, plo = _p(min_level, np.int32); hi, phi = _p(max_level, np.int32)
    qa, pqa = _p(q_angle, np.float32); d, pd = _p(q_desc, np.uint8)
    pxr = pu = pk = None
    if ref_x_right is not None:
        ref_x_right, pxr = _p(ref_x_right, np.float32)
    if usable is not None:
        usable, pu = _p(usable, np.uint8)
    if kp_unavailable is not None:
        kp_unavailable, pk = _p(kp_unavailable, np.uint8)
    out = np.full(max(frm.n, 1), -1, np.int32)
    n = lib().om_projection_match_best(C.byref(frm.c), len(mg), pu, prp, pxr, pmg, plo, phi, pqa, pd, pk, C.c_uint(int(hamm_dist_thr)),
                                       int(check_orientation), out.ctypes.data_as(C.c_void_p))
    return n, out[:frm.n]


def projection_match_keyframes_mutually(f1, f2, scale_factors, usable_1, reproj_1_in_2, pred_level_1_in_2, lm_desc_1,
                                        usable_2, reproj_2_in_1, pred_level_2_in_1, lm_desc_2, margin):
    sf, psf = _p(scale_factors, np.float32)
    u1, pu1 = _p(usable_1, np.uint8); r12, p12 = _p(reproj_1_in_2, np.float32); l12, pl12 = _p(pred_level_1_in_2, np.int32); d1, pd1 = _p(lm_desc_1, np.uint8)
    u2, pu2 = _p(usable_2, np.uint8); r21, p21 = _p(reproj_2_in_1, np.float32); l21, pl21 = _p(pred_level_2_in_1, np.int32); d2, pd2 = _p(lm_desc_2, np.uint8)
    out = np.full(max(f1.n, 1), -1, np.int32)
    n = lib().om_projection_match_keyframes_mutually(C.byref(f1.c), C.byref(f2.c), psf, pu1, p12, pl12, pd1, pu2, p21, pl21, pd2, C.c_float(margin),
                                                     out.ctypes.data_as(C.c_void_p))
    return n, out[:f1.n]


def area_match_in_consistent_area(f1, f2, prev_matched_pts, margin=100, lowe_ratio=0.9, check_orientation=True):
    prev = np.ascontiguousarray(prev_matched_pts, np.float32).copy()
    out = np.full(max(f1.n, 1), -1, np.int32)
    n = lib().om_area_match_in_consistent_area(C.byref(f1.c), C.byref(f2.c), prev.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), int(margin),
                                               C.c_float(lowe_ratio), int(check_orientation))
    return n, out[:f1.n], prev


def angle_checker_invalid(delta_angles, histogram_length=30, num_bins_to_retain=3):
    d, pd = _p(delta_angles, np.float32)
    out = np.zeros(max(len(d), 1), np.uint8)
    lib().om_angle_checker_invalid(pd, len(d), histogram_length, num_bins_to_retain, out.ctypes.data_as(C.c_void_p))
    return out[:len(d)].astype(bool)


def stereo_compute(left_pyr, right_pyr, scale_factors, kps_left, desc_left, kps_right, desc_right, focal_x_baseline, true_baseline):
    L = len(left_pyr)
    lp = [np.ascontiguousarray(a, np.uint8) for a in left_pyr]; rp = [np.ascontiguousarray(a, np.uint8) for a in right_pyr]
    larr = (C.c_void_p * L)(*[a.ctypes.data for a in lp]); rarr = (C.c_void_p * L)(*[a.ctypes.data for a in rp])
    pw = np.array([a.shape[1] for a in lp], np.int32); ph = np.array([a.shape[0] for a in lp], np.int32); ps = np.array([a.strides[0] for a in lp], np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32); isf = (np.float32(1.0) / sf).astype(np.float32)
    lx, plx = _p(kps_left["x"], np.float32); ly, ply = _p(kps_left["y"], np.float32); lo, plo = _p(kps_left["octave"], np.int32); ld, pld = _p(desc_left, np.uint8)
    rx, prx = _p(kps_right["x"], np.float32); ry, pry = _p(kps_right["y"], np.float32); ro, pro = _p(kps_right["octave"], np.int32); rd, prd = _p(desc_right, np.uint8)
    nl, nr = len(lx), len(rx)
    xr = np.full(max(nl, 1), -1, np.float32); dp = np.full(max(nl, 1), -1, np.float32)
    n = lib().om_stereo_compute(larr, rarr, pw.ctypes.data_as(C.c_void_p), ph.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p), L,
                                sf.ctypes.data_as(C.c_void_p), isf.ctypes.data_as(C.c_void_p), nl, plx, ply, plo, pld, nr, prx, pry, pro, prd,
                                C.c_float(focal_x_baseline), C.c_float(true_baseline), xr.ctypes.data_as(C.c_void_p), dp.ctypes.data_as(C.c_void_p))
    return xr[:nl], dp[:nl], n


def color_to_gray(img, rgb_order=False):
    """util::convert_to_grayscale: H x W x {3,4} u8 -> H x W u8 (BGR(A) order unless rgb_order)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, c = img.shape
    out = np.zeros((h, w), np.uint8)
    lib().oo_color_to_gray(img.ctypes.data_as(C.c_void_p), w, h, w * c, c, int(bool(rgb_order)), out.ctypes.data_as(C.c_void_p), w)
    return out


def robust_match_for_triangulation(desc_1, bearing_1, octave_1, angle_1, has_lm_1, is_stereo_1, bow_node_1,
                                   desc_2, bearing_2, angle_2, has_lm_2, is_stereo_2, bow_node_2, E_12, epipole_in_2, scale_factors_1,
                                   check_orientation=True):
    d1, pd1 = _p(desc_1, np.uint8); b1, pb1 = _p(bearing_1, np.float64); o1, po1 = _p(octave_1, np.int32); a1, pa1 = _p(angle_1, np.float32)
    l1, pl1 = _p(has_lm_1, np.uint8); s1, ps1 = _p(is_stereo_1, np.uint8); n1, pn1 = _p(bow_node_1, np.int32)
    d2, pd2 = _p(desc_2, np.uint8); b2, pb2 = _p(bearing_2, np.float64); a2, pa2 = _p(angle_2, np.float32)
    l2, pl2 = _p(has_lm_2, np.uint8); s2, ps2 = _p(is_stereo_2, np.uint8); n2, pn2 = _p(bow_node_2, np.int32)
    E, pE = _p(E_12, np.float64); ep, pep = _p(epipole_in_2, np.float64); sf, psf = _p(scale_factors_1, np.float32)
    out = np.full(max(len(o1), 1), -1, np.int32)
    n = lib().om_robust_match_for_triangulation(len(o1), pd1, pb1, po1, pa1, pl1, ps1, pn1, len(a2), pd2, pb2, pa2, pl2, ps2, pn2, pE, pep, psf,
                                                int(check_orientation), out.ctypes.data_as(C.c_void_p))
    return n, out[:len(o1)]


# ------------------------------------------------------------------ camera-model steps around the extractor (8f rank 3)
def undistort_points(xy, fx, fy, cx, cy, dist, iters=20):
    """camera::perspective::undistort_keypoints: cv::undistortPoints(pts, K, dist, R=I, P=K, MAX_ITER `iters`); OpenVSLAM uses 20."""
    xy, p = _p(np.asarray(xy, np.float32).reshape(-1, 2), np.float32)
    k1, k2, p1, p2, k3 = [float(v) for v in dist]
    out = np.zeros_like(xy)
    D = C.c_double
    lib().oc_undistort_points(p, len(xy), D(fx), D(fy), D(cx), D(cy), D(k1), D(k2), D(p1), D(p2), D(k3), int(iters), out.ctypes.data_as(C.c_void_p))
    return out


def fisheye_undistort_points(xy, fx, fy, cx, cy, dist, max_count=10, eps=1e-8):
    """camera::fisheye::undistort_keypoints: cv::fisheye::undistortPoints(pts, K, D = (k1, k2, k3, k4), R = I, P = K), default criteria."""
    xy, p = _p(np.asarray(xy, np.float32).reshape(-1, 2), np.float32)
    k1, k2, k3, k4 = [float(v) for v in dist]
    out = np.zeros_like(xy)
    D = C.c_double
    lib().oc_fisheye_undistort_points(p, len(xy), D(fx), D(fy), D(cx), D(cy), D(k1), D(k2), D(k3), D(k4), int(max_count), D(eps), out.ctypes.data_as(C.c_void_p))
    return out


def radial_division_undistort_points(xy, fx, fy, cx, cy, distortion):
    """camera::radial_division::undistort_keypoints: p_u = p_d / (1 + distortion |p_d|^2) on normalised coordinates."""
    xy, p = _p(np.asarray(xy, np.float32).reshape(-1, 2), np.float32)
    out = np.zeros_like(xy)
    D = C.c_double
    lib().oc_radial_division_undistort_points(p, len(xy), D(fx), D(fy), D(cx), D(cy), D(distortion), out.ctypes.data_as(C.c_void_p))
    return out


def bearings_perspective(xy, fx, fy, cx, cy):
    xy, p = _p(np.asarray(xy, np.float32).reshape(-1, 2), np.float32)
    out = np.zeros((len(xy), 3))
    D = C.c_double
    lib().oc_bearings_perspective(p, len(xy), D(fx), D(fy), D(cx), D(cy), out.ctypes.data_as(C.c_void_p))
    return out


def bearings_equirectangular(xy, cols, rows):
    xy, p = _p(np.asarray(xy, np.float32).reshape(-1, 2), np.float32)
    out = np.zeros((len(xy), 3))
    lib().oc_bearings_equirectangular(p, len(xy), C.c_double(cols), C.c_double(rows), out.ctypes.data_as(C.c_void_p))
    return out


def project_equirectangular(bearings, cols, rows):
    b, p = _p(np.asarray(bearings, np.float64).reshape(-1, 3), np.float64)
    out = np.zeros((len(b), 2))
    lib().oc_project_equirectangular(p, len(b), C.c_double(cols), C.c_double(rows), out.ctypes.data_as(C.c_void_p))
    return out


# ------------------------------------------------------------------ match::bow_tree (8f rank 2; oracle only so far)
def bow_tree_match_frame_and_keyframe(desc_kf, angle_kf, lm_valid_kf, bow_node_kf, desc_frm, angle_frm, bow_node_frm, lowe_ratio=0.6,
                                      check_orientation=True):
    dk, pdk = _p(desc_kf, np.uint8); ak, pak = _p(angle_kf, np.float32); vk, pvk = _p(lm_valid_kf, np.uint8); nk, pnk = _p(bow_node_kf, np.int32)
    df, pdf = _p(desc_frm, np.uint8); af, paf = _p(angle_frm, np.float32); nf, pnf = _p(bow_node_frm, np.int32)
    out = np.full(max(len(af), 1), -1, np.int32)
    n = lib().om_bow_tree_match_frame_and_keyframe(len(ak), pdk, pak, pvk, pnk, len(af), pdf, paf, pnf, C.c_float(lowe_ratio), int(check_orientation),
                                                   out.ctypes.data_as(C.c_void_p))
    return n, out[:len(af)]


def bow_tree_match_keyframes(desc_1, angle_1, lm_valid_1, bow_node_1, desc_2, angle_2, lm_valid_2, bow_node_2, lowe_ratio=0.6, check_orientation=True):
    d1, pd1 = _p(desc_1, np.uint8); a1, pa1 = _p(angle_1, np.float32); v1, pv1 = _p(lm_valid_1, np.uint8); n1, pn1 = _p(bow_node_1, np.int32)
    d2, pd2 = _p(desc_2, np.uint8); a2, pa2 = _p(angle_2, np.float32); v2, pv2 = _p(lm_valid_2, np.uint8); n2, pn2 = _p(bow_node_2, np.int32)
    out = np.full(max(len(a1), 1), -1, np.int32)
    n = lib().om_bow_tree_match_keyframes(len(a1), pd1, pa1, pv1, pn1, len(a2), pd2, pa2, pv2, pn2, C.c_float(lowe_ratio), int(check_orientation),
                                          out.ctypes.data_as(C.c_void_p))
    return n, out[:len(a1)]


def fuse_best_keypoints(frm, reproj_xy, reproj_x_right, pred_level, lm_desc, scale_factors, inv_level_sigma_sq, margin, usable=None):
    """match::fuse matching core (8f rank 2; oracle only so far) -> (count, best_idx_of_lm)."""
    rp, prp = _p(reproj_xy, np.float32); lv, plv = _p(pred_level, np.int32); d, pd = _p(lm_desc, np.uint8)
    sf, psf = _p(scale_factors, np.float32); iw, piw = _p(inv_level_sigma_sq, np.float32)
    pxr = pu = None
    if reproj_x_right is not None:
        reproj_x_right, pxr = _p(reproj_x_right, np.float32)
    if usable is not None:
        usable, pu = _p(usable, np.uint8)
    out = np.full(max(len(lv), 1), -1, np.int32)
    n = lib().om_fuse_best_keypoints(C.byref(frm.c), len(lv), pu, prp, pxr, plv, pd, psf, piw, C.c_float(margin), out.ctypes.data_as(C.c_void_p))
    return n, out[:len(lv)]
