"""CPU tests of the multi-GPU host logic: the row partition, the packed-minimum exchange protocol
(emulated with the oracle on two row shards) and the torch.distributed plumbing on a world-size-2
gloo group.  The data-path collective itself (NCCL inside the library) needs GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import strotss_oracle as O
from strotss_tensorflow_b200 import distributed as Dm


@pytest.mark.parametrize("N,world", [(16384, 8), (16384, 2), (1000, 4), (100, 8), (129, 2), (1, 2)])
def test_shard_rows_partition(N, world):
    spans = [Dm.shard_rows(N, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == N or any(s[1] == N for s in spans)
    covered = np.zeros(N, dtype=int)
    for r0, r1 in spans:
        assert 0 <= r0 <= r1 <= N
        assert r0 % Dm.TILE_ROWS == 0 or r0 == N
        covered[r0:r1] += 1
    assert np.all(covered == 1)


def test_pack_best_orders_like_the_kernel():
    keys = [Dm.pack_best(v, i) for v, i in [(0.5, 7), (0.5, 3), (-0.25, 0), (0.75, 9), (-1.0, 2)]]
    assert Dm.unpack_best(max(keys)) == (0.75, 9)
    # equal values: the LOWEST index wins the max
    assert Dm.unpack_best(max(Dm.pack_best(0.5, 7), Dm.pack_best(0.5, 3))) == (0.5, 3)
    for v, i in [(0.0, 0), (-0.0, 1), (1.0, 16383), (-3.5, 12)]:
        vv, ii = Dm.unpack_best(Dm.pack_best(v, i))
        assert vv == v and ii == i


def test_exchange_protocol_reproduces_global_relaxed_emd():
    """Each rank sees only its prediction rows; allreduce-max of packed target-row bests plus a sum of
    the per-rank column-min partials must give the single-GPU result."""
    st, co, pr = O.synth_problem(300, 90, 40, eps=1.0, seed=3)
    C = O.cosine_distance(st, pr, np.float32).astype(np.float32)       # (M, N)
    M, N = C.shape
    world = 2
    best = [0] * M
    ry_sum = 0.0
    for rank in range(world):
        r0, r1 = Dm.shard_rows(N, world, rank)
        local = C[:, r0:r1]
        for i in range(M):
            j = int(local[i].argmin())
            best[i] = max(best[i], Dm.pack_best(float(1.0 - local[i, j]), r0 + j))       # value = dot = 1 - cost
        ry_sum += float(local.min(axis=0).sum())
    rx = np.mean([1.0 - Dm.unpack_best(k)[0] for k in best])
    assert rx == pytest.approx(C.min(axis=1).mean(), rel=1e-6)
    assert ry_sum / N == pytest.approx(C.min(axis=0).mean(), rel=1e-6)
    assert [Dm.unpack_best(k)[1] for k in best] == list(C.argmin(axis=1))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, D, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        payload = bytes(range(128)) if rank == 0 else None
        got = Dm.broadcast_bytes(payload, 128, 0)
        ok_id = got == bytes(range(128))
        # every rank "computes" its rows of a gradient; all_gather_rows rebuilds the full matrix
        full = torch.arange(N * D, dtype=torch.float32).reshape(N, D)
        r0, r1 = Dm.shard_rows(N, world, rank)
        mine = torch.zeros(N, D)
        mine[r0:r1] = full[r0:r1]
        rebuilt = Dm.all_gather_rows(mine, N)
        # the scalar block is summed over ranks exactly once
        part = torch.tensor([float(r1 - r0)])
        dist.all_reduce(part)
        out.put((rank, ok_id, bool(torch.equal(rebuilt, full)), float(part.item())))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    N, D = 300, 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, D, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_id, ok_rows, total in res:
        assert ok_id and ok_rows and total == N
