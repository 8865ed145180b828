"""BASELINE config 4: one large scan-to-submap registration — ~130 k-pt live scan vs a ~10 M-pt target (77 scans merged in
the map frame), UNIFORM 1.0 m voxels, DIRECT7; source points sharded across ranks, NCCL all-reduce of the 32-double packed
normal equations per GN iteration. Run plain (1 GPU) or under torchrun (one rank per GPU).
usage: [torchrun ...] tools/bench_submap.py [n_scans=77] [iters=20]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from rolo_b200 import capi, synth
from rolo_b200.dist import make_reduce_hook, shard_bounds, TILE
from rolo_b200.registration import RotVGICP

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
n_scans = int(_pos[0]) if len(_pos) > 0 else 77
iters = int(_pos[1]) if len(_pos) > 1 else 20
sensor = "os1-128x1024"
# 77 overlapping scans give ~1 cm point spacing: with the per-scan 2 cm range noise the 20-NN patches of the merged map
# would be thinner than the noise (meaningless normals, for the reference as well), so the map scans use 4 mm noise
SIGMA_MAP = 0.004
scene = synth.make_scene()
# map: scans along an ellipse inside the room, merged in the map frame
clouds = []
poses = []
for k in range(n_scans):
    a = 2 * np.pi * k / n_scans
    T = synth.pose((0.0, 0.0, np.degrees(a) + 90.0), (5.0 + 18.0 * np.cos(a), 3.0 + 10.0 * np.sin(a), 0.0))
    p = synth.scan_torch(scene, T, sensor, sigma=SIGMA_MAP, seed=1000 + k, device=dev)
    Tm = torch.as_tensor(T, dtype=torch.float32, device=dev)
    q = p.clone(); q[:, :3] = p[:, :3] @ Tm[:3, :3].T + Tm[:3, 3]
    clouds.append(q); poses.append(T)
target = torch.cat(clouds, dim=0).contiguous()
del clouds
# live scan near pose 0 with a small unknown offset; registration result should be T_live (map <- scan)
T_live = poses[0] @ synth.pose((1.0, -1.5, 2.0), (0.10, -0.05, 0.02))
source = synth.scan_torch(scene, T_live, sensor, seed=7, device=dev)
guess = poses[0].astype(np.float32)

r = RotVGICP(device=local)
r.setResolution(1.0); r.setNeighborSearchMethod(capi.DIRECT7)
r.setOptimizerType(capi.OPT_GAUSS_NEWTON); r.setMaximumIterations(iters); r.setRotationEpsilon(1e-30); r.setTransformationEpsilon(1e-30)
torch.cuda.synchronize()
t0 = time.perf_counter()
r.setInputTarget(target); r.setInputSource(source)
if world > 1:
    lo, hi = shard_bounds(source.shape[0], rank, world, align=TILE)
    r.setSourceShard(lo, hi)
    if "--p2p" in sys.argv:
        from rolo_b200.dist import connect_peers
        connect_peers(r, rank, world)          # reduction fused into the kernel over NVLink peer memory
    else:
        r.setReduceHook(make_reduce_hook(), True)
r.calculate_covariances(capi.SOURCE); r.calculate_covariances(capi.TARGET); r.create_voxelmap()
torch.cuda.synchronize()
t_setup = time.perf_counter() - t0
c, H, b = r.linearize(guess.astype(np.float64))          # warm-up of the kernel + NCCL
r.setProfile(True)
t1 = time.perf_counter(); reps = 0; kern_us = 0.0
for _ in range(iters * 3):
    r.linearize(guess.astype(np.float64)); reps += 1; kern_us += r.lastKernelUs()
t_lin = (time.perf_counter() - t1) / reps
r.setProfile(False)
t2 = time.perf_counter(); T = r.align(guess); t_align = time.perf_counter() - t2
erot, etr = synth.pose_error(T.astype(np.float64), T_live)
if world > 1:
    tl = torch.tensor([t_lin, t_setup], dtype=torch.float64, device=dev); dist.all_reduce(tl, op=dist.ReduceOp.MAX); t_lin, t_setup = [float(x) for x in tl]
if rank == 0:
    print(json.dumps({"config": f"scan-to-submap: {source.shape[0]} source pts (sharded x{world}) vs {target.shape[0]} target pts, UNIFORM 1.0 DIRECT7, GN x{iters}",
                      "n_gpus": world, "transport": ("p2p-fused" if "--p2p" in sys.argv else "nccl") if world > 1 else "none", "voxels": int(r._L.rg_voxel_count(r._h)), "correspondences_shard0": int(r.numCorrespondences()),
                      "setup_ms_cov_and_voxelmap": 1e3 * t_setup, "gn_iter_us_wall": 1e6 * t_lin, "gn_iters_per_s": 1.0 / t_lin,
                      "fused_kernel_us": kern_us / reps, "align_ms_incl_voxel_rebuild": 1e3 * t_align,
                      "pose_err_rad": erot, "pose_err_m": etr}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
