"""Host-side mirror of openvslam::match::* (src/openvslam/match/{base,robust,...}.h; names as
recalled in SURVEY.md 8a) over the C ABI of libovs_b200.so."""
import ctypes as C
import numpy as np

from . import _lib

HAMMING_DIST_THR_LOW = 50
HAMMING_DIST_THR_HIGH = 100
MAX_HAMMING_DIST = 256


class _matcher_handle:
    def __init__(self, device=0):
        self._h = C.c_void_p()
        _lib.check(_lib.lib().ovs_matcher_create(int(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().ovs_matcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_kernel_us(self):
        v = C.c_float(0)
        _lib.check(_lib.lib().ovs_matcher_last_kernel_us(self._h, C.byref(v)))
        return v.value


def _desc(a):
    a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
    return a, a.ctypes.data_as(C.c_void_p)


class robust(_matcher_handle):
    """openvslam::match::robust (lowe_ratio_, check_orientation_)."""

    def __init__(self, lowe_ratio=0.6, check_orientation=True, device=0):
        super().__init__(device)
        self.lowe_ratio_ = float(lowe_ratio)
        self.check_orientation_ = bool(check_orientation)

    def brute_force_topk(self, query, train):
        q, pq = _desc(query); t, pt = _desc(train)
        keys = np.zeros((len(q), 4), np.uint32)
        _lib.check(_lib.lib().ovs_match_bruteforce_topk_host(self._h, pq, len(q), pt, len(t), keys.ctypes.data_as(C.c_void_p)))
        return keys

    def brute_force_nearest(self, desc1, desc2):
        d1, p1 = _desc(desc1); d2, p2 = _desc(desc2)
        bi = np.zeros(len(d1), np.int32); bd = np.zeros(len(d1), np.int32); sd = np.zeros(len(d1), np.int32)
        _lib.check(_lib.lib().ovs_match_bruteforce_host(self._h, p1, len(d1), p2, len(d2), bi.ctypes.data_as(C.c_void_p),
                                                        bd.ctypes.data_as(C.c_void_p), sd.ctypes.data_as(C.c_void_p)))
        return bi, bd, sd

    def brute_force_match(self, desc_frm, desc_keyfrm, lm_valid_2=None):
        """robust::brute_force_match(frm, keyfrm, matches) -> matches[(idx_1, idx_2)]."""
        d1, p1 = _desc(desc_frm); d2, p2 = _desc(desc_keyfrm)
        vp = None
        if lm_valid_2 is not None:
            lm_valid_2 = np.ascontiguousarray(lm_valid_2, np.uint8)
            vp = lm_valid_2.ctypes.data_as(C.c_void_p)
        cap = max(min(len(d1), len(d2)), 1)
        pairs = np.zeros((cap, 2), np.int32)
        n = C.c_int(0)
        _lib.check(_lib.lib().ovs_robust_brute_force_match_host(self._h, p1, len(d1), p2, len(d2), vp, C.c_float(self.lowe_ratio_),
                                                                pairs.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        return pairs[:n.value].copy()
