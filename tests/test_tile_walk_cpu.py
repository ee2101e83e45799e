"""CPU tests of the host-callable tile-order functions of the persistent kernels (strotss_debug_tile_walk replays
decode_tile / ss1_decode, the functions the kernels themselves call) and of the rule that sends a GEMM to skewed
tile couples.  No GPU work."""
import ctypes as C

import pytest


@pytest.fixture(scope="module")
def lib():
    from strotss_tensorflow_b200 import build, _lib
    build.build()
    return _lib.load()


def _walk(lib, walk, tiles_m, tiles_n, group_n):
    cap = tiles_m * tiles_n
    tm = (C.c_int * cap)()
    tn = (C.c_int * cap)()
    n = lib.strotss_debug_tile_walk(walk, tiles_m, tiles_n, group_n, tm, tn, cap)
    assert 0 <= n <= cap
    return [(tm[i], tn[i]) for i in range(n)]


@pytest.mark.parametrize("tiles_m,tiles_n,group_n", [(1, 1, 1), (9, 64, 4), (64, 64, 7), (16, 48, 48), (3, 5, 2), (32, 9, 100)])
def test_rectangular_raster_visits_every_tile_once_in_column_groups(lib, tiles_m, tiles_n, group_n):
    order = _walk(lib, 0, tiles_m, tiles_n, group_n)
    assert len(order) == tiles_m * tiles_n
    assert sorted(order) == [(m, n) for m in range(tiles_m) for n in range(tiles_n)]
    # L2 blocking: column groups are visited one after another, all row tiles inside a group before the next group
    groups = [n // group_n for _, n in order]
    assert groups == sorted(groups)
    # inside a group: row tiles ascending, column tile fastest
    for g in set(groups):
        sub = [(m, n) for (m, n) in order if n // group_n == g]
        assert sub == sorted(sub)


@pytest.mark.parametrize("tiles", [1, 2, 9, 64])
def test_triangle_walk_is_the_upper_block_triangle(lib, tiles):
    order = _walk(lib, 1, tiles, tiles, 4)
    assert order == [(m, n) for m in range(tiles) for n in range(m, tiles)]


@pytest.mark.parametrize("tiles_m,tiles_n,group_n", [(16, 64, 4), (16, 48, 5), (16, 16, 3), (8, 64, 64), (1, 7, 2), (4, 4, 1),
                                                      (16, 9, 4)])
def test_trapezoid_walk_of_a_symmetric_row_panel(lib, tiles_m, tiles_n, group_n):
    """Stage 1 of the self-similarity (ss1_decode, trap = 1): row tile r of a panel owns the tiles at or right of its
    diagonal tile; 904 / 648 / 392 / 136 tiles for the four 4096-row panels at N = 16384 (2080 of 4096 in total)."""
    order = _walk(lib, 2, tiles_m, tiles_n, group_n)
    want = [(m, n) for m in range(min(tiles_m, tiles_n)) for n in range(m, tiles_n)]
    assert len(order) == len(want) and sorted(order) == want
    groups = [n // group_n for _, n in order]
    assert groups == sorted(groups)


def test_trapezoid_tile_counts_at_the_bench_size(lib):
    counts = [lib.strotss_debug_tile_walk(2, 16, 64 - 16 * p, 4, None, None, 0) for p in range(4)]
    assert counts == [904, 648, 392, 136] and sum(counts) == 2080 == 64 * 65 // 2


def test_tile_walk_argument_errors(lib):
    assert lib.strotss_debug_tile_walk(3, 4, 4, 1, None, None, 0) < 0
    assert lib.strotss_debug_tile_walk(1, 4, 5, 1, None, None, 0) < 0        # the triangle needs a square grid
    assert lib.strotss_debug_tile_walk(0, 0, 4, 1, None, None, 0) < 0
    assert lib.strotss_debug_tile_walk(0, 4, 4, 1, None, None, 8) < 0        # capacity without buffers


def test_tile_couples_are_chosen_where_they_save_rounds(lib):
    pay = lib.strotss_debug_couples_pay
    K, skew, sms = 35, 16, 148
    # full-size relaxed EMD (N = M = 16384): 4096 tiles = 56 rounds of 74 pairs, 2048 couples = 28 rounds x 1.73
    assert pay(sms, 128, 64, K, skew) == 1
    # covariance backward at N = 16384 (18 row blocks x 64 column tiles) and as a 4-way row shard (x 16)
    assert pay(sms, 18, 64, K, skew) == 1 and pay(sms, 18, 16, K, skew) == 1
    # an 8-way row shard's covariance backward is one round of 72 plain tiles: couples would only serialise it
    assert pay(sms, 18, 8, K, skew) == 0
    # a handful of tiles, a single column tile, a short K loop (couple cost 2.0) or a disabled skew never pay
    assert pay(sms, 4, 9, K, skew) == 0 and pay(sms, 128, 1, K, skew) == 0
    assert pay(sms, 128, 64, 9, skew) == 0 and pay(sms, 128, 64, K, -1) == 0
    # the cost model: a couple of K blocks with skew s moves 2*s*32 + (K - s)*48 KB against 2*K*32 KB
    for s in (0, 8, 16, 35):
        cost = (2 * s * 32 + (K - s) * 48) / (K * 32)
        plain, couples = -(-64 * 64 // 74), -(-64 * 32 // 74)
        assert pay(sms, 128, 64, K, s) == int(couples * cost < plain)
