"""Row-sharded multi-GPU evaluation: host-side plumbing (one process per GPU).

The whole exchange of an evaluation happens inside the C library (csrc/api.cu): two small allreduces
-- {sum of N floats, max of the packed (value, ~index) minima of the M style rows} and a sum of a
(16 + D)-float block -- plus, for the symmetric self-similarity and covariance tiles that are dealt out
over the ranks, bf16 sign blocks and partial covariance tiles.  With CUDA IPC available everything moves
through peer windows over NVLink (`Handle.comm_transport() == 1`), otherwise over NCCL.  This module
only (a) mirrors the library's row partition, (b) ships the NCCL unique id from rank 0 to the other
ranks with torch.distributed (any backend: nccl on GPUs, gloo in the CPU tests) and (c) optionally
all-gathers the per-rank gradient rows for callers that want the full (N, D) gradient on every rank.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

TILE_ROWS = 128          # shards are multiples of the MMA tile height (BM in csrc/gemm_core.cuh)
ID_BYTES = 128           # NCCL_UNIQUE_ID_BYTES


def shard_rows(N: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [r0, r1) of the prediction owned by `rank` -- identical to shard_of() in csrc/api.cu."""
    if world <= 1:
        return 0, N
    per = -(-N // world)
    per = -(-per // TILE_ROWS) * TILE_ROWS
    r0 = min(rank * per, N)
    r1 = min(r0 + per, N)
    return r0, r1


def pack_best(value: float, index: int) -> int:
    """Python mirror of sb::pack_best (csrc/common.cuh): max over packed keys = (largest value, lowest index)."""
    import struct
    u = struct.unpack("<I", struct.pack("<f", value))[0]
    u = (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)
    return (u << 32) | ((~index) & 0xFFFFFFFF)


def unpack_best(key: int) -> Tuple[float, int]:
    import struct
    k = (key >> 32) & 0xFFFFFFFF
    u = (k & 0x7FFFFFFF) if (k & 0x80000000) else (~k & 0xFFFFFFFF)
    return struct.unpack("<f", struct.pack("<I", u))[0], (~key) & 0xFFFFFFFF


def broadcast_bytes(payload: Optional[bytes], nbytes: int, src: int = 0, group=None, device=None) -> bytes:
    """Broadcast `nbytes` bytes from rank `src` over torch.distributed (works on gloo and nccl)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    if dist.get_rank(group) == src:
        assert payload is not None and len(payload) == nbytes
        t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def attach(handle, group=None) -> Tuple[int, int]:
    """Create the NCCL communicator of `handle` across the ranks of `group`.  Returns (rank, world)."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if world == 1:
        handle.comm_init(0, 1, None)
        return 0, 1
    payload = None
    if rank == 0:
        buf = C.create_string_buffer(ID_BYTES)
        code = handle.lib.strotss_comm_unique_id(buf)
        if code != 0:
            raise _lib.StrotssError("strotss_comm_unique_id failed: NCCL (libnccl.so.2) is not loadable")
        payload = buf.raw
    uid = broadcast_bytes(payload, ID_BYTES, 0, group)
    handle.comm_init(rank, world, uid)
    return rank, world


def all_gather_rows(grad: torch.Tensor, N: int, group=None) -> torch.Tensor:
    """Fill the rows owned by the other ranks into `grad` (N x D; each rank wrote only its shard)."""
    world = dist.get_world_size(group)
    if world == 1:
        return grad
    per = shard_rows(N, world, 0)[1]
    D = grad.shape[1]
    pad = torch.zeros(per * world, D, device=grad.device, dtype=grad.dtype)
    r0, r1 = shard_rows(N, world, dist.get_rank(group))
    mine = torch.zeros(per, D, device=grad.device, dtype=grad.dtype)
    mine[: r1 - r0] = grad[r0:r1]
    dist.all_gather_into_tensor(pad, mine, group=group) if hasattr(dist, "all_gather_into_tensor") and grad.is_cuda else \
        dist.all_gather(list(pad.chunk(world, dim=0)), mine, group=group)
    return pad[:N]
