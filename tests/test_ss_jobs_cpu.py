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


def copies(N, world, rank, panel):
    lib = _lib.load()
    buf = (C.c_longlong * (8 * 64))(); total = C.c_longlong(0)
    n = lib.strotss_debug_ss_copies(N, world, rank, panel, buf, 64, C.byref(total))
    assert n >= 0
    return [tuple(buf[8 * k:8 * k + 8]) for k in range(n)], total.value


@pytest.mark.parametrize("N,world,panel", [(16384, 2, 4096), (16384, 4, 4096), (16384, 8, 4096), (4096, 2, 4096), (5 * 512, 5, 256),
                                           (16384, 8, 1024), (6 * 512 * 3, 6, 512), (32768, 2, 4096), (7 * 256, 7, 256)])
def test_sign_block_copies_fill_every_window_exactly_once(N, world, panel):
    """When the bf16 sign blocks travel instead of the fp32 products: the copies of all ranks write every element of every
    rank's window exactly once, a block lands as [source rows][receiver rows], and what the window holds for a receive-list
    entry is exactly the mirrored tiles (source rows x own rows) that entry stands for."""
    per = N // world
    plans = [plan(N, world, k, panel) for k in range(world)]
    totals = []
    windows = {}
    for k in range(world):
        cps, total = copies(N, world, k, panel)
        totals.append(total)
        jobs = plans[k][0]
        for (job, peer, off, ld, i0, i1, j0, j1) in cps:
            r0, r1, c0, c1, diag, kind = jobs[job]
            assert r0 <= i0 < i1 <= r1 and c0 <= j0 < j1 <= c1          # inside the job's P buffer
            assert peer != k and peer * per <= j0 and j1 <= (peer + 1) * per and k * per <= i0 and i1 <= (k + 1) * per
            assert not diag or j0 >= r1                                  # a trapezoid job writes every tile right of its rows
            w = windows.setdefault(peer, {})
            for i in range(i0, i1, 256):
                for j in range(j0, j1, 256):
                    key = off + (i - i0) * ld + (j - j0)
                    assert key not in w
                    w[key] = (i, j, ld)
    assert len(set(totals)) == 1                                          # every rank's window has the same size
    for q in range(world):
        # replay the receiver's view: blocks one after another in receive-list order, each [source rows][receiver rows]
        recvs = plans[q][2]
        sends_of = {k: plans[k][1] for k in range(world)}
        off = 0
        w = windows.get(q, {})
        seen = 0
        for (peer, o0, o1) in recvs:
            # source rows of this entry = rows of `peer` whose tiles against [o0, o1) it computed
            src = sorted({i for (i, j, ld) in w.values() if peer * per <= i < (peer + 1) * per and o0 <= j < o1})
            assert src, (q, peer)
            s0, s1 = src[0], src[-1] + 256
            assert len(src) == (s1 - s0) // 256
            for i in range(s0, s1, 256):
                for j in range(o0, o1, 256):
                    assert w[off + (i - s0) * (o1 - o0) + (j - o0)] == (i, j, o1 - o0)
                    seen += 1
            off += (s1 - s0) * (o1 - o0)
        assert off == totals[q] and seen == len(w)
