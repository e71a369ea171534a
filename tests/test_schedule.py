"""Host logic of the Gram kernels' tile schedule (no device needed): every tile of the block is computed by
exactly one task, for full and lower-triangular blocks, and the refinement keeps the per-sub-partition loads
within one tile of each other on the shapes of the benchmark workload."""
import ctypes as C

import numpy as np
import pytest

import diaglib_b200 as D


def _sched(ntp, ntq, sym):
    cover = np.zeros(ntp * ntq, np.int32)
    load = np.zeros(4, np.int32)
    nt = D.lib().diaglib_b200_k_gram_schedule(ntp, ntq, int(sym), cover.ctypes.data_as(C.c_void_p), load.ctypes.data_as(C.c_void_p))
    return nt, cover.reshape(ntq, ntp).T, load


@pytest.mark.parametrize("sym", [False, True])
def test_every_tile_is_computed_once(sym):
    for ntp in range(1, 17):
        for ntq in ([ntp] if sym else range(1, 17)):
            nt, cover, load = _sched(ntp, ntq, sym)
            if ((ntp + 1) // 2) * ((ntq + 3) // 4) > 30:   # more coarse tasks than the 15 consumer warps have slots
                assert nt == -1
                continue
            want = np.tril(np.ones((ntp, ntq), np.int32)) if sym else np.ones((ntp, ntq), np.int32)
            assert nt >= 1 and nt <= 30
            assert np.array_equal(cover, want), (ntp, ntq, sym)
            assert load.sum() == want.sum()


def test_benchmark_shapes_are_balanced():
    # C3: x^T u of ortho_vs_x (74 x 37 -> 10 x 5 tiles), first-iteration Gram (74 x 74, lower), the
    # 111 x 111 Rayleigh-Ritz Gram (lower): largest sub-partition load within one tile of the mean
    for ntp, ntq, sym in ((10, 5, False), (10, 10, True), (14, 14, True)):
        _, cover, load = _sched(ntp, ntq, sym)
        assert load.max() - cover.sum() / 4.0 <= 1.0, (ntp, ntq, sym, load)
