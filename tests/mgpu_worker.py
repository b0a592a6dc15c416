"""Worker for tests/test_gpu_multi.py: launched with torch.distributed.run, one rank per GPU.

Runs particle-sharded scans (SVN-ICP class: peer-memory record exchange over NVLink, and the ncclAllGather fallback;
SVGD-ICP class: ncclAllGather) and compares them, on rank 0, with the same scan on a single GPU.  Exit code 0 = parity.
Slices of >= 64 particles exercise the pose-space particle ordering inside each rank's slice (getters must map back)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOL_SVN = 1e-7   # identical algorithm; only the grouping of the fp32 Gauss-Newton partial sums and of the fp64 Stein sums depends on the slice size
TOL_SVGD = 2e-5  # Adam divides by the running RMS of the gradient, which amplifies that ~1e-7 relative difference (measured 3e-6)


def main():
    import torch
    import torch.distributed as dist
    import svn_icp_b200 as sv
    from svn_icp_b200 import synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    #        P    full   es     class   flags                 twice
    cases = ((96, True, False, "svn", 0, False), (37, False, False, "svn", 0, False), (64, True, True, "svn", 0, False),
             (256, True, False, "svn", 0, False), (300, False, True, "svn", 0, False), (256, True, False, "svn", sv.FLAG_NO_PARTICLE_SORT, False),
             (200, True, False, "svn", sv.FLAG_NCCL_GATHER, False), (130, True, True, "svn", sv.FLAG_NCCL_GATHER, False),
             (160, True, False, "svn", 0, True), (144, True, False, "svn", 0, "rescan"), (90, False, True, "svn", 0, "rescan"),
             (48, True, False, "svgd", 0, False), (43, True, True, "svgd", 0, False))
    for P, full, es, cls, flags, twice in cases:
        pb = synth.make_problem(P, sensor="32", scan_index=6, n_map_scans=6, seed=0xC0FFEE)
        if cls == "svn":
            prm = sv.SteinICPParam(iterations=5 if twice is True else 10, KNN_count=64, max_dist=3.0, lr=1.0, SVN_full_grad=full, check_early_stop=es,
                                   convergence_threshold=1e-2, flags=flags)
            make = lambda: sv.SVNICP(prm, pb.init_pose, device=local)
        else:  # the SVGD-ICP class shards the same way (first-order record, one all-gather per iteration)
            prm = sv.SteinICPParam(iterations=10, KNN_count=64, max_dist=3.0, lr=0.03, optimizer="Adam", check_early_stop=es,
                                   convergence_threshold=5e-2)
            make = lambda: sv.SVGDICP(prm, pb.init_pose, device=local)

        def scan(h):
            # "rescan": three full scans on the same handle (the sequence numbers of the peer exchange must run on: with numbers
            # that restart, the flags left by the previous scan satisfy every wait and the ranks only agree while they stay in step)
            for k in range(3 if twice == "rescan" else 1):
                h.add_cloud(pb.source[k::2] if twice == "rescan" else pb.source, pb.target, pb.init_pose)
                h.set_initial_mean(pb.R0, pb.t0)
                assert h.stein_align() == sv.ALIGN_SUCCESS
            if twice is True:  # stein_align again without add_cloud: the other ranks' x must carry over
                assert h.stein_align() == sv.ALIGN_SUCCESS

        icp = make()
        uid = [sv.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        icp.init_sharding(uid[0], rank, world)
        lo, hi = icp.slice()
        assert hi - lo >= 1 and (hi - lo) <= -(-P // world)
        scan(icp)
        got = icp.get_particles()
        mean, cov, its, hist = icp.get_transformation(), icp.get_cov_matrix(), icp.iterations_done(), icp.get_particle_history()
        # every rank must hold the identical full result
        t = torch.from_numpy(np.concatenate([got, mean, cov, [its]])).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(t, ref))
        if rank == 0:
            single = make()
            scan(single)
            err = np.abs(single.get_particles() - got).max()
            herr = np.abs(single.get_particle_history() - hist).max()
            tol = TOL_SVN if cls == "svn" else TOL_SVGD
            good = bool(err < tol and herr < tol + 1e-6 and its == single.iterations_done())
            print(f"{cls} P={P} full={full} es={es} flags={flags} twice={twice} ranks={world}: |sharded - single| particles {err:.3e} "
                  f"history {herr:.3e} iters {its} vs {single.iterations_done()} same_on_all_ranks={same} -> {'ok' if good and same else 'FAIL'}",
                  flush=True)
            ok &= good
        flag = torch.tensor([1 if (ok and same) else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
        icp.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
