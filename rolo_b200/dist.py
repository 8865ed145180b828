"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests).

Two partitionings:
  * batch of independent scan pairs: pairs are sharded per rank, no collective (bench.py);
  * one large registration with the SOURCE points sharded: every rank holds the full target (replicated
    voxel map, built locally, no communication) and a contiguous block of the source cloud; the 32-double
    packed normal equations of every linearize / compute_error evaluation are all-reduced (sum), so each
    rank's host optimiser sees identical numbers and takes identical steps.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

RES_N = 32   # packed result length (rolo_b200/csrc/common.cuh: RES_N)


TILE = 256   # TMA tile of the fused kernel: shard starts must be tile-aligned (rg_set_source_shard)


def shard_bounds(n: int, rank: int, world: int, align: int = 1) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`. Blocks cover [0, n); with align > 1 every block start is a multiple of
    `align` (blocks then differ by at most `align` points, trailing ranks may be empty for tiny n)."""
    units = (n + align - 1) // align
    base, rem = divmod(units, world)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(lo_u * align, n), min(hi_u * align, n)


class _CudaView:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def make_reduce_hook(group=None):
    """Returns fn(buf_ptr, count, on_device, stream) -> int that all-reduces (sum) `count` doubles in place.
    Host buffers go through the default (gloo/nccl) process group via a CPU tensor; device buffers are wrapped
    without a copy and reduced by NCCL on the context's own stream."""
    import torch
    import torch.distributed as dist

    def hook(buf_ptr, count, on_device, stream):
        if on_device:
            t = torch.as_tensor(_CudaView(buf_ptr, count), device="cuda")
            ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
            with torch.cuda.stream(ext):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            arr = np.ctypeslib.as_array(C.cast(buf_ptr, C.POINTER(C.c_double)), shape=(count,))
            t = torch.from_numpy(arr)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return 0

    return hook


def connect_peers(reg, rank: int, world: int, group=None):
    """Fused NVLink reduction: all-gather the CUDA-IPC handles of the ranks' exchange buffers and map them (one node)."""
    import torch.distributed as dist
    mine = reg.shardExport()
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    reg.shardConnect(rank, world, handles)
    dist.barrier(group=group)     # nobody writes into a buffer that its owner has not zeroed yet


def sharded_registration(target, source, rank: int, world: int, device: int = -1, on_device: bool = True, group=None, transport: str = "nccl"):
    """A RotVGICP whose source is this rank's shard and whose evaluations are summed across `world` ranks, either by an
    all-reduce hook (transport="nccl": NCCL on device buffers / gloo on host buffers) or inside the fused kernel itself
    over NVLink peer memory (transport="p2p")."""
    from .registration import RotVGICP
    r = RotVGICP(device=device)
    r.setInputTarget(target)
    r.setInputSource(source)          # full cloud: covariances need un-cut neighbourhoods; only the evaluation is sharded
    if world > 1:
        lo, hi = shard_bounds(int(source.shape[0]), rank, world, align=TILE)
        if hi <= lo:
            raise ValueError("source cloud too small for this many ranks")
        r.setSourceShard(lo, hi)
        if transport == "p2p":
            connect_peers(r, rank, world, group)
        else:
            r.setReduceHook(make_reduce_hook(group), on_device)
    return r
