"""CPU test of the work split of the row-sharded symmetric self-similarity (csrc/ss_jobs.h, through strotss_debug_ss_jobs):
over all ranks the jobs -- every tile standing for its mirror image -- cover each 256 x 256 tile of the N x N matrix exactly
once, every rank gets the same number of tiles, and the send / receive lists of the mirrored stage-2 products match pairwise and
name exactly the foreign rows a rank computed products for.  Reference: the matrices of nn/losses.py:56-68 are symmetric."""
import ctypes as C

import numpy as np
import pytest

from strotss_tensorflow_b200 import _lib


def plan(N, world, rank, panel):
    lib = _lib.load()
    jobs = (C.c_int * 48)(); sends = (C.c_int * 24)(); recvs = (C.c_int * 24)(); cnt = (C.c_int * 3)()
    rc = lib.strotss_debug_ss_jobs(N, world, rank, panel, jobs, sends, recvs, cnt)
    if rc == 0:
        return None
    assert rc == 1
    j = [tuple(jobs[6 * k:6 * k + 6]) for k in range(cnt[0])]
    s = [tuple(sends[3 * k:3 * k + 3]) for k in range(cnt[1])]
    r = [tuple(recvs[3 * k:3 * k + 3]) for k in range(cnt[2])]
    return j, s, r


@pytest.mark.parametrize("N,world,panel", [(16384, 2, 4096), (16384, 4, 4096), (16384, 8, 4096), (4096, 2, 4096), (4096, 4, 4096),
                                           (3 * 256 * 4, 3, 4096), (5 * 512, 5, 256), (16384, 8, 1024), (6 * 512 * 3, 6, 512),
                                           (32768, 2, 4096), (7 * 256, 7, 256)])
def test_jobs_cover_every_tile_once(N, world, panel):
    T = N // 256
    cover = np.zeros((T, T), dtype=np.int32)
    per_rank = []
    foreign = {}          # rank -> set of (owner, row tile) it produced mirrored stage-2 products for
    plans = [plan(N, world, k, panel) for k in range(world)]
    assert all(p is not None for p in plans)
    per = N // world
    for k, (jobs, sends, recvs) in enumerate(plans):
        ntiles = 0
        rows_touched = set()
        assert jobs[0][0] == k * per
        for (r0, r1, c0, c1, diag, kind) in jobs:
            assert r0 % 256 == 0 and r1 % 256 == 0 and c0 % 256 == 0 and c1 % 256 == 0 and r1 > r0 and c1 > c0
            assert k * per <= r0 and r1 <= (k + 1) * per and r1 - r0 <= panel
            if diag:
                assert c0 == r0 and kind == 0
            for tm in range(r0 // 256, r1 // 256):
                for tn in range(c0 // 256, c1 // 256):
                    if diag and tn < tm:
                        continue
                    ntiles += 1
                    cover[tm, tn] += 1
                    if tn != tm or not diag:
                        assert tn != tm
                        cover[tn, tm] += 1
                        rows_touched.add(tn)
        per_rank.append(ntiles)
        foreign[k] = {t for t in rows_touched if not (k * per <= t * 256 < (k + 1) * per)}
    assert (cover == 1).all()
    assert max(per_rank) == min(per_rank) == T * (T + 1) // 2 // world
    # exchange lists: what k sends to p is what p receives from k, and it is exactly k's foreign rows
    for k, (jobs, sends, recvs) in enumerate(plans):
        sent_tiles = set()
        for (peer, r0, r1) in sends:
            assert peer != k and peer * per <= r0 < r1 <= (peer + 1) * per
            assert (k, r0, r1) in plans[peer][2]
            sent_tiles |= set(range(r0 // 256, r1 // 256))
        assert sent_tiles == foreign[k]
        for (peer, r0, r1) in recvs:
            assert (k, r0, r1) in plans[peer][1]
        assert len(sends) == len(recvs) == (world - 1) // 2 + (1 if world % 2 == 0 else 0)


@pytest.mark.parametrize("N,world,panel", [(1000, 2, 4096), (16384 + 256, 2, 4096), (4096 + 512, 2, 4096), (16384, 1, 4096),
                                           (256 * 4, 4, 4096)])
def test_ragged_shapes_fall_back(N, world, panel):
    # blocks that are not whole tiles (or, for an even world, not whole tile pairs) use rectangular row sharding
    assert plan(N, world, 0, panel) is None
