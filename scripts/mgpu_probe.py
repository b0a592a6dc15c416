"""Sharded-vs-single probe at full cloud size (torchrun, one rank per GPU): python -m torch.distributed.run ... scripts/mgpu_probe.py P [iterations]
Prints, per flag combination, the largest |sharded - single| over particles / history / mean and the iteration where the history departs."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import svn_icp_b200 as sv
from svn_icp_b200 import synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 250
I = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pb = synth.make_problem_saturated(P, sensor="64")
VARIANTS = (("default", 0), ("nccl_gather", sv.FLAG_NCCL_GATHER), ("no_sort", sv.FLAG_NO_PARTICLE_SORT), ("filter_full", sv.FLAG_FILTER_FULL))
if len(sys.argv) > 3:
    VARIANTS = tuple(v for v in VARIANTS if v[0] in sys.argv[3].split(","))
for name, fl in VARIANTS:
    prm = sv.SteinICPParam(iterations=I, KNN_count=100, max_dist=3.0, lr=1.0, SVN_full_grad=True, flags=fl)
    icp = sv.SVNICP(prm, pb.init_pose, device=local)
    uid = [sv.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    icp.init_sharding(uid[0], rank, world)
    for _ in range(3):
        icp.add_cloud(pb.source, pb.target, pb.init_pose); icp.set_initial_mean(pb.R0, pb.t0)
        assert icp.stein_align() == sv.ALIGN_SUCCESS
    got, hist, mean = icp.get_particles().reshape(6, -1).copy(), icp.get_particle_history().copy(), icp.get_transformation().copy()
    dist.barrier()
    if rank == 0:
        one = sv.SVNICP(prm, pb.init_pose, device=local)
        one.add_cloud(pb.source, pb.target, pb.init_pose); one.set_initial_mean(pb.R0, pb.t0)
        assert one.stein_align() == sv.ALIGN_SUCCESS
        p1, h1, m1 = one.get_particles().reshape(6, -1), one.get_particle_history(), one.get_transformation()
        d = np.abs(got - p1).max(axis=0)
        dh = np.abs(hist.reshape(hist.shape[0], -1) - h1.reshape(h1.shape[0], -1)).max(axis=1)
        first = int(np.argmax(dh > 1e-4)) if (dh > 1e-4).any() else -1
        print(f"P={P} ranks={world} {name}: particles {d.max():.3e} (beyond 1e-4: {(d > 1e-4).sum()}, worst particle {int(d.argmax())}), history {dh.max():.3e} "
              f"(first iteration beyond 1e-4: {first}), per-iteration history diff {np.array2string(dh, precision=1, max_line_width=1000)}, mean {np.abs(mean - m1).max():.3e}; "
              f"particles beyond 1e-5: {np.nonzero(d > 1e-5)[0][:20].tolist()}", flush=True)
        one.close()
    dist.barrier()
    icp.close()
dist.barrier()
dist.destroy_process_group()
