"""Host-side tiling logic of the tcgen05 GEMM / implicit-GEMM path (sidlsg_debug_tiling runs without a GPU):
every plan the choosers produce for the SD1.5 / SD2.1 UNet shapes must satisfy the kernel's structural limits
(UMMA N granularity, 64-column TMA boxes for MN-major B, 192 KB shared-memory ring, 512 TMEM columns)."""
import ctypes

import pytest

from sid_lsg_b200._lib import lib

KINDS = {"fwd": 0, "dgrad": 1, "wgrad": 2, "conv_fwd": 3, "conv_dgrad": 4, "conv_wgrad": 5}
RING_BYTES = 4 * (128 * 64 * 2 + 256 * 64 * 2)


def plan(kind, M, N, K):
    lib.load()
    out = (ctypes.c_int * 8)()
    st = lib._fns["sidlsg_debug_tiling"](KINDS[kind], M, N, K, out)
    assert st == 0, lib.last_error()
    keys = ("bm2", "block_n", "m_tiles", "n_tiles", "splits", "stages", "stage_bytes", "kb_total")
    return dict(zip(keys, out))


def unet_shapes():
    """(kind, M, N, K) of the dense contractions of one SD1.5 iteration at CFG batch 64 and 32 plus SD2.1's 1024-wide
    text projections and a few ragged shapes."""
    shapes = []
    for B in (32, 64):
        for hw, C in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
            M = B * hw
            for (n, k) in ((C, C), (3 * C, C), (8 * C, C), (C, 4 * C), (2 * C, 768), (2 * C, 1024)):
                shapes.append(("fwd", M, n, k))
                shapes.append(("dgrad", M, k, n))
                shapes.append(("wgrad", n, k, M))
            for (cin, cout) in ((C, C), (2 * C, C), (C, 2 * C), (3 * C, C)):
                if cin % 64 or cout > 2560:
                    continue
                shapes.append(("conv_fwd", M, cout, cin))
                shapes.append(("conv_dgrad", M, cin, cout))
                shapes.append(("conv_wgrad", cout, cin, M))
    shapes += [("fwd", 300, 136, 640), ("fwd", 1000, 640, 768), ("fwd", 64, 16, 64), ("dgrad", 77 * 64, 768, 640),
               ("wgrad", 136, 640, 300 // 64 * 64 + 64), ("conv_fwd", 128, 64, 64), ("conv_wgrad", 64, 64, 128)]
    return shapes


@pytest.mark.parametrize("kind,M,N,K", unet_shapes())
def test_plans_respect_kernel_limits(kind, M, N, K):
    p = plan(kind, M, N, K)
    mn_major_b = kind in ("dgrad", "wgrad", "conv_dgrad", "conv_wgrad")
    bn = p["block_n"]
    assert 16 <= bn <= 256 and bn % (64 if mn_major_b else 16) == 0, p
    per_group = -(-N // bn)
    assert p["n_tiles"] == per_group * (9 if kind == "conv_wgrad" else 1), p
    rows = 256 if p["bm2"] else 128
    assert p["m_tiles"] == -(-M // rows), p
    assert 2 <= p["stages"] <= 8 and p["stages"] * p["stage_bytes"] <= RING_BYTES, p
    b_bytes = (-(-bn // 64)) * 8192 if mn_major_b else bn * 128
    assert p["stage_bytes"] == (2 if p["bm2"] else 1) * 16384 + b_bytes, p
    assert p["splits"] >= 1, p
    if kind in ("wgrad", "conv_wgrad"):
        assert not p["bm2"] and p["splits"] <= max(1, p["kb_total"] // 8), p
    else:
        assert p["splits"] == 1, p
    if p["bm2"]:
        # two 256-column accumulators must fit the 512 TMEM columns; only deep reductions pay for the exposed epilogue
        assert bn <= 256 and (kind.startswith("conv") or p["kb_total"] >= 16), p


def test_balanced_width_for_320_columns():
    """N = 320 as 256 + 64 puts every wide tile on the even CTAs of the 148-CTA grid; the chooser must pick a split
    whose tiles are equal (160 + 160) for the K-major case."""
    p = plan("fwd", 262144, 320, 320)
    assert p["block_n"] == 160 and p["n_tiles"] == 2 and not p["bm2"], p


def test_deep_reductions_get_256_row_tiles():
    assert plan("conv_fwd", 64 * 4096, 320, 320)["bm2"] == 1
    assert plan("fwd", 16384, 1280, 5120)["bm2"] == 1
    assert plan("fwd", 262144, 2560, 320)["bm2"] == 0       # 5 k-blocks: the single-buffered epilogue would dominate


def test_split_k_fills_the_grid_without_a_ragged_last_wave():
    for (m, n, k) in ((320, 320, 262144), (640, 640, 65536), (1280, 1280, 16384)):
        p = plan("wgrad", m, n, k)
        tiles = p["m_tiles"] * p["n_tiles"] * p["splits"]
        waves = tiles / 148.0
        assert tiles >= 100, p
        frac = waves - int(waves)
        assert frac == 0 or frac >= 0.6 or waves < 1, (p, tiles)


# ---- tile cursor of the persistent GEMM (sidlsg_debug_tile_walk runs the kernel's own TileCursor on the host) ----------
def _walk(geom, cta, grid, max_tiles=4096):
    lib.load()
    g = (ctypes.c_int * 9)(*geom)
    out = (ctypes.c_int * (8 * max_tiles))()
    n = lib._fns["sidlsg_debug_tile_walk"](g, cta, grid, max_tiles, out)
    assert n >= 0, lib.last_error()
    return [tuple(out[8 * i: 8 * i + 8]) for i in range(n)]


def _decode(geom, tile):
    """the division-per-tile decode the cursor replaced (sid_lsg_b200/csrc/gemm_tc.cu history)"""
    m_tiles, n_tiles, splits, batched, nb2, kb_total, block_n, N, bm2 = geom
    base = m_tiles * n_tiles * (batched if batched else 1)
    split, r = divmod(tile, base)
    m_blk, n_blk = divmod(r, n_tiles)
    b1 = b2 = 0
    if batched:
        bidx, m_blk = divmod(m_blk, m_tiles)
        b1, b2 = divmod(bidx, nb2 if nb2 > 0 else 1)
    m0 = m_blk * (256 if bm2 else 128)
    n_in = n_blk * block_n
    per = (kb_total + splits - 1) // splits
    kb0 = split * per
    return (tile, m0, n_in, min(block_n, N - n_in), kb0, min(kb_total, kb0 + per), b1, b2)


@pytest.mark.parametrize("geom,grid", [
    ((2048, 2, 1, 0, 0, 5, 160, 320, 0), 148),      # M262144 N320 K320: two 160-wide tiles per row block
    ((1024, 3, 1, 0, 0, 10, 256, 640, 0), 148),     # N640 as 256 + 256 + 128
    ((512, 10, 1, 0, 0, 5, 256, 2560, 1), 148),     # 256-row tiles
    ((3, 3, 37, 0, 0, 4096, 128, 320, 0), 148),     # split-K weight gradient: split-major order, ragged last split
    ((5, 7, 4, 0, 0, 1023, 96, 640, 0), 111),       # stride > tiles per split, stride % n_tiles != 0
    ((2, 1, 1, 512, 8, 3, 80, 77, 0), 148),         # batched (attention scores of the exact path): B x heads problems
    ((1, 1, 1, 0, 0, 1, 16, 16, 0), 148),           # fewer tiles than CTAs
    ((7, 5, 3, 6, 2, 11, 64, 300, 0), 13),          # everything at once
])
def test_tile_cursor_matches_division_decode(geom, grid):
    m_tiles, n_tiles, splits, batched = geom[:4]
    total = m_tiles * n_tiles * splits * (batched if batched else 1)
    seen = []
    for cta in sorted({0, 1, grid // 2, grid - 1}):
        got = _walk(geom, cta, grid)
        want = [_decode(geom, t) for t in range(cta, total, grid)]
        assert got == want, (geom, grid, cta, got[:3], want[:3])
        seen += [g[0] for g in got]
    if grid <= 16:      # every tile exactly once over the whole grid
        allt = sorted(t for cta in range(grid) for (t, *_rest) in _walk(geom, cta, grid))
        assert allt == list(range(total))
