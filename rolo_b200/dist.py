"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests).

Two partitionings:
  * batch of independent scan pairs: pairs are sharded per rank, no collective (bench.py);
  * one large registration with the SOURCE points sharded: every rank holds the full target and a contiguous block of the
    source cloud; the 32-double packed normal equations of every linearize / compute_error evaluation are summed over the
    ranks (NCCL all-reduce, or inside the fused kernel over NVLink peer memory), so each rank's host optimiser sees identical
    numbers and takes identical steps. The SETUP can be sharded as well (shard_target): every rank estimates the covariances
    and accumulates the voxel sums of one slice of the target; the compact voxel lists are all-gathered once per align.
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


class _CudaViewT:
    """Zero-copy typed view of a raw device pointer (CUDA array interface v2)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def gather_ragged(local_keys, local_rows, group=None):
    """All-gather of per-rank lists of different lengths: local_keys (n_r,) int64 and local_rows (n_r, w) float64 tensors (CPU with
    gloo, CUDA with NCCL). Returns (keys, rows) = the concatenation over ranks in rank order, identical on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_local = int(local_keys.shape[0])
    counts = [None] * world
    dist.all_gather_object(counts, n_local, group=group)
    mx = max(max(counts), 1)
    w = int(local_rows.shape[1])
    pk = torch.zeros(mx, dtype=local_keys.dtype, device=local_keys.device); pk[:n_local] = local_keys
    pr = torch.zeros((mx, w), dtype=local_rows.dtype, device=local_rows.device); pr[:n_local] = local_rows
    gk = [torch.empty_like(pk) for _ in range(world)]; gr = [torch.empty_like(pr) for _ in range(world)]
    dist.all_gather(gk, pk, group=group); dist.all_gather(gr, pr, group=group)
    keys = torch.cat([gk[r][:counts[r]] for r in range(world)]).contiguous()
    rows = torch.cat([gr[r][:counts[r]] for r in range(world)]).contiguous()
    return keys, rows


def make_voxel_merge_hook(reg, group=None):
    """Merge step of a target-sharded registration: every rank's compact voxel list (packed key, 10 sums) is all-gathered over
    NCCL on the context's own stream and the concatenation handed back to the library (rg_voxel_import adds equal keys)."""
    import torch

    def hook(keys_ptr, sums_ptr, n_local, stream):
        ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
        with torch.cuda.stream(ext):
            if n_local > 0:
                k = torch.as_tensor(_CudaViewT(keys_ptr, n_local, "<i8"), device="cuda")
                r = torch.as_tensor(_CudaViewT(sums_ptr, n_local * 10, "<f8"), device="cuda").view(n_local, 10)
            else:
                k = torch.empty(0, dtype=torch.int64, device="cuda"); r = torch.empty((0, 10), dtype=torch.float64, device="cuda")
            keys, rows = gather_ragged(k, r, group)
            reg.voxelImport(keys.data_ptr(), rows.data_ptr(), int(keys.shape[0]))
            hook.keep = (keys, rows)          # alive until the next build (the import is queued on `ext`)
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


def sharded_registration(target, source, rank: int, world: int, device: int = -1, on_device: bool = True, group=None, transport: str = "nccl",
                         shard_target: bool = False):
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
        if shard_target:      # the SETUP is sharded too: covariances + voxel sums of one slice of the target per rank, merged once per align
            r.setTargetShard(rank, world)
            r.setVoxelMergeHook(make_voxel_merge_hook(r, group))
    return r
