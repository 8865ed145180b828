"""Run under torchrun (one rank per GPU): a point-sharded registration with an NCCL all-reduce of the packed
normal equations per evaluation must reproduce the single-GPU registration.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sharded_check.py [sensor] [uniform]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from rolo_b200 import capi, synth
from rolo_b200.dist import sharded_registration
from rolo_b200.registration import RotVGICP

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gloo = dist.new_group(backend="gloo")     # host-buffer variant of the hook
sensor = sys.argv[1] if len(sys.argv) > 1 else "small"
uniform = len(sys.argv) > 2 and sys.argv[2] == "uniform"
tgt, src, T_gt = synth.scan_pair(sensor)


def configure(r):
    if uniform:
        r.setResolution(1.0); r.setNeighborSearchMethod(capi.DIRECT7)
    else:
        r.setPolarResolution(0.175, 0.175, 2.0)
    r.setOptimizerType(capi.OPT_LM)


ok = True
for on_device, transport, shard_target in ((True, "p2p", False), (True, "nccl", False), (False, "nccl", False), (True, "p2p", True)):
    r = sharded_registration(tgt, src, rank, world, device=local, on_device=on_device, group=None if on_device else gloo, transport=transport,
                             shard_target=shard_target)
    configure(r)
    t0 = time.perf_counter(); T = r.align(); dt = time.perf_counter() - t0
    Td = r.getFinalPoseDouble()
    ref = RotVGICP(device=local); configure(ref); ref.setInputTarget(tgt); ref.setInputSource(src)
    Tr = ref.align(); Trd = ref.getFinalPoseDouble()
    drot, dtr = synth.pose_error(Td, Trd)
    same_iters = r.nrIterations() == ref.nrIterations()
    # every rank must hold the same pose (they all saw the same reduced numbers)
    t = torch.tensor(Td.ravel(), device="cuda"); tmax = t.clone(); tmin = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    spread = float((tmax - tmin).abs().max())
    # a merged voxel map adds the same numbers in another order: 1e-9 instead of bit equality with the single-GPU run
    good = drot <= 1e-9 and dtr <= 1e-9 and same_iters and spread == 0.0
    ok = ok and good
    if rank == 0:
        print(f"transport={transport} on_device={on_device} shard_target={shard_target}: world={world} pose vs single-GPU {drot:.2e} rad {dtr:.2e} m, iters {r.nrIterations()} vs {ref.nrIterations()}, "
              f"rank spread {spread:.1e}, align {1e3*dt:.2f} ms -> {'OK' if good else 'FAIL'}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
