"""Host-side model of the overlay path's list walk (sd_kernels.cuh: win_load / win_segment_mask / warp_overlay_all): a warp
keeps a 32-entry WINDOW of an ascending position list in registers, serves both 64-row segments of a tile from it, reloads when
a segment runs off the window's end, and carries only the window's base to the next tile.  The model mirrors the kernel's
statements lane for lane; brute force over the list is the checker.  (The CUDA path itself is checked against the oracle by the
-m gpu tests on update deltas / delete masks.)"""
import numpy as np
import pytest

INF = 0x7FFFFFFF
THREADS, RPT = 256, 4
TILE = THREADS * RPT


class Win:
    def __init__(self):
        self.p = np.full(32, INF, np.int64)
        self.base = 0
        self.loads = 0


def win_load(pos, base, w):
    w.base = base
    idx = base + np.arange(32)
    w.p = np.where(idx < len(pos), pos[np.minimum(idx, max(len(pos) - 1, 0))] if len(pos) else INF, INF).astype(np.int64)
    w.loads += 1


def win_segment_mask(pos, a, w):
    while np.all(w.p < a):
        win_load(pos, w.base + 32, w)
    first = w.base + int(np.sum(w.p < a))
    mask = 0
    while True:
        inr = (w.p >= a) & (w.p < a + 64)
        for p in w.p[inr]:
            mask |= 1 << int(p - a)
        if w.p[31] >= a + 64:
            break
        win_load(pos, w.base + 32, w)
    return mask, first


@pytest.mark.parametrize("seed,density", [(1, 0.005), (2, 0.05), (3, 0.6), (4, 1.0), (5, 0.0)])
def test_window_walk_equals_brute_force(seed, density):
    r = np.random.default_rng(seed)
    nrows = 5 * TILE + 321
    pos = np.flatnonzero(r.random(nrows) < density).astype(np.int64)
    ntiles = (nrows + TILE - 1) // TILE
    for tile0 in (0, 2):                      # a chunk may start anywhere: the cursor starts at lower_bound
        for warp in range(THREADS // 32):
            w = Win()
            for tile in range(tile0, ntiles):
                a0 = tile * TILE + warp * 64
                base = int(np.searchsorted(pos, a0)) if tile == tile0 else w.base   # warp_cursor_init / the base kept in shared memory
                win_load(pos, base, w)
                for u in range(RPT // 2):
                    a = a0 + u * 2 * THREADS
                    mask, first = win_segment_mask(pos, a, w)
                    want = pos[(pos >= a) & (pos < a + 64)]
                    assert mask == sum(1 << int(p - a) for p in want)
                    if len(want):
                        assert pos[first] == want[0]         # delta values are addressed by first + popcount(mask below the row's bit)
            if density <= 0.05 and len(pos):
                assert w.loads <= 2 * (ntiles - tile0) + len(pos) // 32 + 2   # ~one load per tile (+ reloads), not 4 per tile
