"""The formulation of the device radix sort (tests/sort_device_model.py) against numpy's stable argsort."""
import numpy as np
import pytest

from sort_device_model import sort_pairs


@pytest.mark.parametrize("n,end_bit", [(1, 3), (31, 5), (2047, 11), (2048, 11), (2049, 11), (5000, 11), (7000, 19)])
def test_model_is_a_stable_sort(n, end_bit):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << end_bit, n).astype(np.uint32)
    if n > 4096:
        keys[: n // 3] = keys[0]
    vals = np.arange(n, dtype=np.uint64) * np.uint64(3) + np.uint64(1)
    ks, vs = sort_pairs(keys, vals, end_bit)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ks, keys[order]) and np.array_equal(vs, vals[order])
