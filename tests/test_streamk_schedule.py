"""Host-side model of the stream-K decomposition used by skinny_chain_kernel (llava_align_b200/csrc/decode_gemm.cu): the
kernel derives, from integer formulas alone, which CTAs share an output tile, a CTA's segment index inside a tile and the
workspace slots it may touch.  A wrong formula is a hang (a tile counter that never reaches its target) or a silent
overlap of partials, so the arithmetic is pinned here on CPU for the shapes the engine launches and for adversarial ones."""
import itertools

import pytest

SK_BM, SK_BK = 256, 64


def geom(N, K, grid):
    """sk_geom(): tiles, k-blocks, units, active CTAs, workspace segments per tile."""
    nkb = (K + SK_BK - 1) // SK_BK
    tiles = (N + SK_BM - 1) // SK_BM
    units = tiles * nkb
    active = min(units, grid)
    min_len = units // active
    maxseg = nkb // max(min_len, 1) + 2
    return nkb, tiles, units, active, maxseg


def cta_range(c, units, active):
    """sk_range(): [u0, u0 + n) of CTA c."""
    if c >= active:
        return 0, 0
    u0 = c * units // active
    return u0, (c + 1) * units // active - u0


def sharers(t, nkb, units, active):
    """c_first / c_last of tile t as the kernel computes them."""
    s_lo, s_hi = t * nkb, (t + 1) * nkb
    c_first = ((s_lo + 1) * active + units - 1) // units - 1
    c_last = (s_hi * active + units - 1) // units - 1
    return c_first, c_last


SHAPES = [  # (N, K): LLaVA-1.5 7B / 13B projections, TP shards, the tiny test model, and awkward sizes
    (12288, 4096), (4096, 4096), (22016, 4096), (4096, 11008), (32000, 4096),
    (15360, 5120), (5120, 5120), (27648, 5120), (5120, 13824),
    (1920, 5120), (5120, 640), (3456, 5120), (5120, 1728), (4000, 5120),
    (768, 256), (256, 256), (1024, 256), (256, 512), (512, 256),
    (1000, 1376), (256, 64), (32000, 512), (272, 8), (16, 4096)]


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("grid", [1, 2, 7, 132, 148, 160])
def test_streamk_ranges_partition_the_units_and_tiles_know_their_sharers(N, K, grid):
    nkb, tiles, units, active, maxseg = geom(N, K, grid)
    # 1. the CTA ranges tile [0, units) exactly, in order, none empty among the active CTAs
    pos = 0
    owners = {}  # unit -> cta
    for c in range(grid):
        u0, n = cta_range(c, units, active)
        if c < active:
            assert n >= 1 and u0 == pos
            for u in range(u0, u0 + n):
                owners[u] = c
            pos = u0 + n
        else:
            assert n == 0
    assert pos == units
    # 2. per tile: the CTAs that own at least one of its k-blocks are exactly c_first..c_last, contiguous
    for t in range(tiles):
        real = sorted({owners[u] for u in range(t * nkb, (t + 1) * nkb)})
        c_first, c_last = sharers(t, nkb, units, active)
        assert real == list(range(c_first, c_last + 1)), (t, real, c_first, c_last)
        nseg = c_last - c_first + 1
        # 3. the workspace holds maxseg slots per tile: a segment index never leaves them
        assert 1 <= nseg <= maxseg, (t, nseg, maxseg)
    # 4. a CTA meets a tile in ONE contiguous segment, and whole-tile segments are exactly the unshared tiles
    for c in range(active):
        u0, n = cta_range(c, units, active)
        i = 0
        seen = set()
        while i < n:
            u = u0 + i
            t, kb0 = divmod(u, nkb)
            ln = min(n - i, nkb - kb0)
            assert t not in seen
            seen.add(t)
            whole = kb0 == 0 and ln == nkb
            c_first, c_last = sharers(t, nkb, units, active)
            assert whole == (c_first == c_last), (c, t)
            i += ln


def test_fixup_items_are_taken_exactly_once():
    """Phase 2: sharer `myseg` of `nseg`, epilogue warp `ew` of 8, takes items myseg + nseg*ew, + nseg*8, ... of the tile's
    8 * RP/NR items; every item must be reduced by exactly one (CTA, warp) for every row-count / sharer-count combination."""
    for RP, nseg in itertools.product(range(16, 257, 16), range(2, 25)):
        NR = 16 if nseg <= 4 else 8 if nseg <= 8 else 4
        items = 8 * (RP // NR)
        taken = []
        for myseg in range(nseg):
            for ew in range(8):
                taken += list(range(myseg + nseg * ew, items, nseg * 8))
        assert sorted(taken) == list(range(items)), (RP, nseg)
