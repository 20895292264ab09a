"""CPU: the host-side launch plan of the tensor pass (no device needed).
Launch groups must cover the batch exactly in whole 128-query tiles and keep the SMs busy;
scan phases must cover every row tile exactly once and grow geometrically."""
import ctypes as C

import numpy as np
import pytest

from cortex_b200 import _capi


def plan(nq, rows, sms=148, sample=32, growth=8):
    L = _capi.load()
    g = np.zeros(4096, np.uint32)
    p = np.zeros(256, np.uint32)
    gn, pn, hk = C.c_uint32(g.size), C.c_uint32(p.size), C.c_double(0)
    st = L.cx_debug_tensor_plan(nq, rows, sms, sample, growth, g.ctypes.data, C.byref(gn), p.ctypes.data,
                                C.byref(pn), C.byref(hk))
    assert st == 0, L.cx_last_error()
    return g[:gn.value].tolist(), p[:pn.value].tolist(), hk.value


@pytest.mark.parametrize("nq", [1, 5, 128, 129, 1024, 5000, 16384, 18944, 18945, 100_000, 1_000_003])
def test_groups_cover_the_batch_in_whole_tiles(nq):
    groups, _, _ = plan(nq, 1_000_000)
    assert sum(groups) == nq
    assert all(g % 128 == 0 for g in groups[:-1])          # only the last group may hold a ragged tile
    assert all(0 < g <= 148 * 128 for g in groups)


def test_groups_keep_the_sms_busy():
    def cost(groups):  # time in units of "one CTA scans the whole shard"
        return sum(1.0 / (148 // ((g + 127) // 128)) for g in groups)
    g, _, _ = plan(16384, 1_000_000)                       # 128 tiles: one group would idle 20 of 148 SMs
    assert cost(g) < 0.9 and cost([16384]) == 1.0
    assert len(g) <= 5
    g, _, _ = plan(1024, 1_000_000)                        # 8 tiles x 18 row splits = 144 CTAs: one group
    assert g == [1024]
    g, _, _ = plan(148 * 128 * 3, 1_000_000)
    assert g == [148 * 128] * 3


@pytest.mark.parametrize("rows,sample,growth", [(1_000_000, 32, 8), (10_000_000, 32, 4), (5000, 20, 8),
                                                (1_000_000, 64, 0), (300, 2, 8), (50_000_000, 64, 4)])
def test_phases_cover_every_tile_once(rows, sample, growth):
    _, phases, hits = plan(1024, rows, sample=sample, growth=growth)
    n_tiles = (rows + 255) // 256
    assert sum(phases) == n_tiles and all(p > 0 for p in phases)
    if growth < 2:
        assert phases == [n_tiles]
        return
    seen = sample
    for p in phases[:-1]:
        assert p == growth * seen                          # every phase but the last: growth x rows seen so far
        seen = sum(phases[:phases.index(p) + 1])
    # expected nominations per kept key: growth per full phase (the last one absorbs a remainder of up to
    # twice that), and far fewer than a single phase would give
    assert hits <= growth * (len(phases) + 1) + 1e-9
    if n_tiles > 4 * growth * sample:
        assert hits < 0.5 * n_tiles / sample
