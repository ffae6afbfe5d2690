"""least_numa_test.go's helper-level vectors (TestNormalizeScore :706-756, TestMinDistance :758-912)."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("count,is_min,want", [(1, 0, 88), (2, 0, 76), (8, 0, 4), (1, 1, 94), (2, 1, 82), (8, 1, 10)])
def test_normalize_score(oracle, count, is_min, want):
    lib = oracle.lib()
    lib.orc_nrt_normalize_least_numa.restype = C.c_int64
    assert lib.orc_nrt_normalize_least_numa(count, is_min, 8) == want  # nodeconfig.DefaultMaxNUMANodes = 8


COSTS = [[10, 12, 20, 20], [12, 10, 20, 20], [20, 20, 10, 12], [20, 20, 12, 10]]


@pytest.mark.parametrize("combos,cost,want", [
    ([[0], [1], [2], [3]], COSTS, 10.0),
    ([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]], COSTS, 11.0),
    ([[0, 1, 2], [1, 2, 3], [0, 2, 3]], COSTS, np.float32(14.888889)),
    ([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]], [[-1] * 4] * 4, 255.0),  # no Costs: maxDistanceValue
])
def test_min_avg_distance(oracle, combos, cost, want):
    lib = oracle.lib()
    lib.orc_nrt_min_avg_distance.restype = C.c_float
    c = np.ascontiguousarray(cost, dtype=np.int32)
    cb = np.ascontiguousarray(combos, dtype=np.int32)
    got = lib.orc_nrt_min_avg_distance(C.c_void_p(c.ctypes.data), 4, C.c_void_p(cb.ctypes.data), len(combos), len(combos[0]))
    assert np.float32(got) == np.float32(want)  # the Go test compares float32 with !=
