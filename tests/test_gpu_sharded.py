"""Sharded single registration (BASELINE configs[3]).
* test_target_shard_setup_on_one_gpu: the target-sharded SETUP (covariances + voxel sums of one slice of the target per rank, lists
  merged through rg_voxel_import) emulated with two contexts on ONE GPU: runs on the driver's single-GPU box.
* test_sharded_equals_single_gpu: the real multi-process path (torchrun, NCCL / fused NVLink reduction, target-sharded setup) on
  however many GPUs the box has (>= 2), else skipped."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_single_gpu():
    from rolo_b200 import capi
    n = capi.load().rg_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "sharded_check.py"), "small"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 4      # fused NVLink P2P reduction, NCCL on device buffers, gloo on host buffers, P2P + target-sharded setup


def test_target_shard_setup_on_one_gpu():
    import numpy as np
    import torch
    from rolo_b200 import capi, synth
    from rolo_b200.registration import RotVGICP
    tgt, src, T_gt = synth.scan_pair("small")
    world = 2

    def make(rank=None):
        r = RotVGICP(device=0)
        r.setResolution(1.0); r.setNeighborSearchMethod(capi.DIRECT7)
        r.setInputTarget(tgt); r.setInputSource(src)
        if rank is not None:
            r.setTargetShard(rank, world)
        return r

    whole = make()
    whole.create_voxelmap()
    kw, mw, cw, nw = whole.getVoxels()
    cov_whole = whole.getTargetCovariances()
    # pass 1: every "rank" builds its slice; the hook keeps a copy of the local list and imports it alone
    parts, lists = [], {}
    for rank in range(world):
        r = make(rank)

        def hook(keys_ptr, sums_ptr, n_local, stream, r=r, rank=rank):
            k = torch.as_tensor(_View(keys_ptr, n_local, "<i8"), device="cuda").clone()
            v = torch.as_tensor(_View(sums_ptr, n_local * 10, "<f8"), device="cuda").clone()
            torch.cuda.synchronize()
            lists[rank] = (k, v)
            r.voxelImport(k.data_ptr(), v.data_ptr(), n_local)
            return 0
        r.setVoxelMergeHook(hook)
        r.create_voxelmap()
        parts.append(r)
    assert 0 < lists[0][0].shape[0] and 0 < lists[1][0].shape[0]
    # the slices partition the target: per-voxel point counts of the two partial maps add up to the whole map's
    k0, _, _, n0 = parts[0].getVoxels(); k1, _, _, n1 = parts[1].getVoxels()
    cnt = {}
    for k, n in ((k0, n0), (k1, n1)):
        for key, c in zip(map(tuple, k), n):
            cnt[key] = cnt.get(key, 0) + int(c)
    assert cnt == {tuple(k): int(c) for k, c in zip(kw, nw)}
    # the covariances a rank computed for its slice equal the whole computation's (un-cut neighbourhoods); the slices cover every point
    c0 = parts[0].getTargetCovariances(); c1 = parts[1].getTargetCovariances()
    own0 = np.all(c0 == cov_whole, axis=(1, 2)); own1 = np.all(c1 == cov_whole, axis=(1, 2))
    assert np.all(own0 | own1) and own0.sum() >= len(tgt) // 2 - 2048 and own1.sum() >= len(tgt) // 2 - 2048
    # pass 2: merge = import of the concatenated lists; the merged map equals the whole map
    keys = torch.cat([lists[0][0], lists[1][0]]).contiguous(); rows = torch.cat([lists[0][1], lists[1][1]]).contiguous()
    for r in parts:
        r.voxelImport(keys.data_ptr(), rows.data_ptr(), int(keys.shape[0]))
        km, mm, cm, nm = r.getVoxels()
        assert np.array_equal(km, kw) and np.array_equal(nm, nw)
        assert np.max(np.abs(mm - mw)) <= 1e-9 and np.max(np.abs(cm - cw)) <= 1e-9
    # and a linearize on the merged map equals the whole context's
    cw_, Hw, bw = whole.linearize(T_gt)
    cm_, Hm, bm = parts[0].linearize(T_gt)
    assert abs(cm_ - cw_) <= 1e-9 * abs(cw_) and np.max(np.abs(Hm - Hw)) <= 1e-9 * np.max(np.abs(Hw))


class _View:
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
