"""CPU checks that do not need a GPU: the __host__ __device__ arithmetic of csrc/orb_math.cuh / ba_math.cuh compiled into a
test shim (tests/hostcheck), and two host restatements of the tree distribution (the list-based round-1 product code, now test
infrastructure: tests/hostcheck/keypoint_tree.cpp; the array-pass prototype tree_levelsync.cpp), compared with the oracle.
The product's tree distribution is a CUDA kernel (k_tree_distribute): tests/test_tree_device_model.py pins its formulation on
the CPU, tests/test_extractor_gpu.py the kernel itself."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from openvslam_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "hostcheck")])
    lib = C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))
    lib.hc_fast_atan2.restype = C.c_float
    lib.hc_fast_atan2.argtypes = [C.c_float, C.c_float]
    lib.hc_resize_px.restype = C.c_uint8
    return lib


def _pack(c):
    return (c["x"].astype(np.uint32) | (c["y"].astype(np.uint32) << 12) | (c["score"].astype(np.uint32) << 24))


@pytest.mark.parametrize("w,h,seed", [(800, 500, 5), (300, 600, 6), (333, 333, 7)])
def test_tree_distribution_equals_oracle(hc, oracle, w, h, seed):
    img = synth.frame(w, h, seed=seed)
    cands = oracle.fast_detect(np.ascontiguousarray(img[19:-19, 19:-19]), 20)
    packed = _pack(cands)
    for N in (1, 2, 7, 50, 217, 869, 2500, 10 ** 6):
        ref = oracle.distribute_via_tree(cands, 19, w - 19, 19, h - 19, N)
        out = np.zeros(len(cands) + 8, np.int32)
        m = hc.hc_distribute(packed.ctypes.data_as(C.c_void_p), len(cands), 19, w - 19, 19, h - 19, C.c_uint(N),
                             out.ctypes.data_as(C.c_void_p))
        assert m == len(ref) and np.array_equal(out[:m], ref), (N, m, len(ref))


def test_tree_distribution_ties_and_duplicates(hc, oracle):
    # many equal scores: the first-maximum and creation-order tie rules must agree
    rng = np.random.default_rng(3)
    n = 3000
    xy = rng.choice(600 * 400, n, replace=False)
    cands = np.zeros(n, oracle.FASTPT_DTYPE)
    cands["x"] = xy % 600; cands["y"] = xy // 600; cands["score"] = rng.integers(20, 23, n)
    order = np.lexsort((cands["x"], cands["y"]))
    cands = cands[order]
    for N in (10, 100, 400, 1000):
        ref = oracle.distribute_via_tree(cands, 19, 619, 19, 419, N)
        out = np.zeros(n + 8, np.int32)
        m = hc.hc_distribute(_pack(cands).ctypes.data_as(C.c_void_p), n, 19, 619, 19, 419, C.c_uint(N), out.ctypes.data_as(C.c_void_p))
        assert m == len(ref) and np.array_equal(out[:m], ref)


def test_fast_score_arithmetic_equals_oracle(hc, oracle):
    img = synth.frame(400, 300, seed=8)
    ref = oracle.fast_score_map(img)
    ref[ref < 7] = 0
    out = np.zeros_like(img)
    hc.hc_fast_score_map(img.ctypes.data_as(C.c_void_p), 400, 300, 400, 7, out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, ref)


def test_fast_atan2_arithmetic_equals_oracle(hc, oracle):
    rng = np.random.default_rng(1)
    for _ in range(5000):
        y = float(rng.integers(-300000, 300000)); x = float(rng.integers(-300000, 300000))
        assert hc.hc_fast_atan2(y, x) == oracle.fast_atan2(y, x)


def test_descriptor_offsets_equal_oracle(hc, oracle):
    # rotated sampling offsets of all 512 pattern points, for many angles
    import re
    rows = re.findall(r"\{(-?\d+),(-?\d+),(-?\d+),(-?\d+)\}", open(os.path.join(HERE, "..", "oracle", "orb_pattern.inc")).read())
    pat = np.array(rows, np.int8).reshape(512, 2)
    rng = np.random.default_rng(2)
    blurred = np.arange(64 * 64, dtype=np.uint32).reshape(64, 64)  # unique value per pixel -> decodes offsets
    for a in rng.uniform(0, 360, 200).astype(np.float32):
        drow = np.zeros(512, np.int32); dcol = np.zeros(512, np.int32)
        hc.hc_descriptor_offsets(C.c_float(float(a)), pat.ctypes.data_as(C.c_void_p), drow.ctypes.data_as(C.c_void_p),
                                 dcol.ctypes.data_as(C.c_void_p))
        s, c = oracle.sincosf(float(np.float32(np.float64(a) * np.pi / 180.0)))
        for i in range(0, 512, 37):
            px, py = int(pat[i, 0]), int(pat[i, 1])
            r = int(np.rint(np.float32(np.float32(px * np.float32(s)) + np.float32(py * np.float32(c)))))
            cc = int(np.rint(np.float32(np.float32(px * np.float32(c)) - np.float32(py * np.float32(s)))))
            assert (drow[i], dcol[i]) == (r, cc)
        assert np.abs(drow).max() <= 18 and np.abs(dcol).max() <= 18


@pytest.mark.parametrize("w,h,seed", [(800, 500, 15), (300, 600, 16), (333, 333, 17), (1920, 960, 18)])
def test_levelsync_tree_prototype_equals_list_version(hc, oracle, w, h, seed):
    """Round-2 design prototype (tests/hostcheck/tree_levelsync.cpp): the tree distribution as array passes (partitions,
    prefix sums, one sort per finishing round) gives exactly the list-based product implementation's selection, in order."""
    img = synth.frame(w, h, seed=seed)
    cands = oracle.fast_detect(np.ascontiguousarray(img[19:-19, 19:-19]), 20)
    packed = _pack(cands)
    n = len(cands)
    for N in (1, 2, 3, 7, 50, 217, 500, 869, 1300, 2500, 4000, n - 1, n, 10 ** 6):
        a = np.zeros(n + 8, np.int32); b = np.zeros(n + 8, np.int32)
        ma = hc.hc_distribute(packed.ctypes.data_as(C.c_void_p), n, 19, w - 19, 19, h - 19, C.c_uint(max(N, 1)), a.ctypes.data_as(C.c_void_p))
        mb = hc.hc_distribute_levelsync(packed.ctypes.data_as(C.c_void_p), n, 19, w - 19, 19, h - 19, C.c_uint(max(N, 1)), b.ctypes.data_as(C.c_void_p))
        assert ma == mb and np.array_equal(a[:ma], b[:mb]), (N, ma, mb)


def test_levelsync_tree_prototype_random_clouds(hc):
    """Random candidate clouds with heavy ties and duplicates, many target counts."""
    rng = np.random.default_rng(23)
    for trial in range(60):
        n = int(rng.integers(1, 3000))
        w = int(rng.integers(60, 1500)); h = int(rng.integers(60, 1500))
        span = max(2, int(rng.integers(2, 40)))
        x = (rng.integers(0, w - 38, n) // span * span).astype(np.uint32); y = (rng.integers(0, h - 38, n) // span * span).astype(np.uint32)
        sc = rng.integers(7, 12, n).astype(np.uint32)
        order = np.lexsort((x, y))                                  # row-major like the FAST output
        packed = (x | (y << 12) | (sc << 24))[order].astype(np.uint32)
        for N in (1, int(rng.integers(1, 50)), int(rng.integers(50, 2000)), n, 10 ** 6):
            a = np.zeros(n + 8, np.int32); b = np.zeros(n + 8, np.int32)
            ma = hc.hc_distribute(packed.ctypes.data_as(C.c_void_p), n, 19, w - 19, 19, h - 19, C.c_uint(N), a.ctypes.data_as(C.c_void_p))
            mb = hc.hc_distribute_levelsync(packed.ctypes.data_as(C.c_void_p), n, 19, w - 19, 19, h - 19, C.c_uint(N), b.ctypes.data_as(C.c_void_p))
            assert ma == mb and np.array_equal(a[:ma], b[:mb]), (trial, n, N, ma, mb)
