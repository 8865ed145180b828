"""CPU tests of the N>1 path (world_size 2, gloo): shard bounds, the all-reduce hook on host buffers, and
"sharded equals whole": per-shard normal equations summed by the hook equal the full-cloud ones."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rolo_b200.dist import shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 130000, 262144):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
            a = [shard_bounds(n, r, w, align=256) for r in range(w)]
            assert a[0][0] == 0 and a[-1][1] == n and all(a[i][1] == a[i + 1][0] for i in range(w - 1))
            assert all(lo % 256 == 0 or lo == n for lo, _ in a)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as ob
    from rolo_b200 import synth
    from rolo_b200.dist import make_reduce_hook, shard_bounds, RES_N
    hook = make_reduce_hook()
    # 1. raw buffer semantics
    buf = np.arange(RES_N, dtype=np.float64) * (rank + 1)
    assert hook(buf.ctypes.data, RES_N, False, 0) == 0
    ok1 = np.array_equal(buf, np.arange(RES_N, dtype=np.float64) * sum(range(1, world + 1)))
    # 2. sharded == whole (oracle as the per-shard evaluator; the voxel map is replicated)
    tgt, src, T_gt = synth.scan_pair("tiny")
    whole = ob.Oracle(); whole.setPolarResolution(0.175, 0.175, 2.0); whole.setInputTarget(tgt); whole.setInputSource(src)
    whole.calculateCovariances(0); whole.calculateCovariances(1)
    cs, ct = whole.getSourceCovariances(), whole.getTargetCovariances()
    lo, hi = shard_bounds(len(src), rank, world)
    part = ob.Oracle(); part.setPolarResolution(0.175, 0.175, 2.0); part.setInputTarget(tgt); part.setInputSource(src[lo:hi])
    part.setTargetCovariances(ct); part.setSourceCovariances(cs[lo:hi])
    c, H, b = part.linearize(T_gt)
    packed = np.concatenate([[c], H.ravel(), b]).astype(np.float64)
    hook(packed.ctypes.data, packed.size, False, 0)
    cw, Hw, bw = whole.linearize(T_gt)
    ref = np.concatenate([[cw], Hw.ravel(), bw])
    ok2 = np.allclose(packed, ref, rtol=1e-11, atol=1e-9)
    out[rank] = (ok1, ok2)
    dist.destroy_process_group()


def test_allreduce_hook_and_sharded_equals_whole_gloo():
    world = 2
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] == (True, True) for r in range(world)), dict(out)


def _merge_worker(rank, world, port, out):
    """Target-sharded setup on host tensors: every rank accumulates the voxel sums of its slice of the target, the ragged per-rank
    lists are all-gathered (rolo_b200.dist.gather_ragged, gloo) and merged by key; the result must equal the oracle's whole map."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as ob
    from rolo_b200 import synth
    from rolo_b200.dist import gather_ragged, shard_bounds
    tgt, src, _ = synth.scan_pair("tiny")
    o = ob.Oracle(); o.setResolution(1.0); o.setInputTarget(tgt); o.setInputSource(src)
    o.calculateCovariances(1)
    cov = o.getTargetCovariances()
    o.createVoxelmap()
    keys_w, mean_w, cov_w, n_w = o.getVoxels()
    # this rank's slice (any partition works; the library uses slices of the Morton order)
    lo, hi = shard_bounds(len(tgt), rank, world, align=32)
    pts = tgt[lo:hi, :3].astype(np.float64)
    key3 = np.floor(pts / 1.0 - 0.5).astype(np.int64)                      # vmp_voxel.hpp:199-201
    packed = ((key3[:, 0] + (1 << 20)) << 42) | ((key3[:, 1] + (1 << 20)) << 21) | (key3[:, 2] + (1 << 20))
    uk, inv = np.unique(packed, return_inverse=True)
    rows = np.zeros((len(uk), 10))
    c6 = np.stack([cov[lo:hi, 0, 0], cov[lo:hi, 0, 1], cov[lo:hi, 0, 2], cov[lo:hi, 1, 1], cov[lo:hi, 1, 2], cov[lo:hi, 2, 2]], axis=1)
    np.add.at(rows, inv, np.concatenate([pts, c6, np.ones((len(pts), 1))], axis=1))
    gk, gr = gather_ragged(torch.from_numpy(uk), torch.from_numpy(rows))
    gk, gr = gk.numpy(), gr.numpy()
    mk, minv = np.unique(gk, return_inverse=True)
    merged = np.zeros((len(mk), 10)); np.add.at(merged, minv, gr)
    # compare with the oracle's map (sorted by (x, y, z) key == sorted by the packed key)
    okk = len(mk) == len(n_w) and np.array_equal(merged[:, 9].astype(np.int64), np.asarray(n_w, dtype=np.int64))
    okm = np.allclose(merged[:, :3] / merged[:, 9:10], mean_w[:, :3], rtol=0, atol=1e-9)
    cw6 = np.stack([cov_w[:, 0, 0], cov_w[:, 0, 1], cov_w[:, 0, 2], cov_w[:, 1, 1], cov_w[:, 1, 2], cov_w[:, 2, 2]], axis=1)
    okc = np.allclose(merged[:, 3:9] / merged[:, 9:10], cw6, rtol=0, atol=1e-9)
    out[rank] = (bool(okk), bool(okm), bool(okc), int(len(gk)))
    dist.destroy_process_group()


def test_target_shard_merge_equals_whole_gloo():
    world = 2
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_merge_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r][:3] == (True, True, True) for r in range(world)), dict(out)
    assert out[0][3] == out[1][3] > 0          # every rank sees the same gathered list
