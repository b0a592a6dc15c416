"""Time the REFERENCE ITSELF on the GPU (oracle/_ref/libsvnicp_ref_cuda.so: the reference's unmodified SVNICP/SVGDICP
sources and its vendored knn.cu against libtorch CUDA, no device swap) on the bench workload.  MEASUREMENT INFRASTRUCTURE
ONLY: spawned as a subprocess by bench.py (the reference freezes P in function-static tensors, and a failure or an
out-of-memory of the reference must not take the bench down).  Prints one JSON line.

usage: python -m oracle.ref_gpu_run P [iterations] [n_s_cap (0 = whole scan)] [particles_out.npy]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import oracle as orc
    import bench
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    I = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOAD["iterations"]
    cap = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    import torch
    pb, _ = bench.make_problem(P)
    src = pb.source if not cap else bench._subsample(pb, cap)
    ref = orc.Reference(cuda=True)
    W = bench.WORKLOAD
    prm = orc.make_params(iterations=I, knn_count=W["K"], max_dist=W["max_dist"], lr=W["lr"], svn_full_grad=W["svn_full_grad"])
    # warm-up with the SAME particle count (Q8) on a tiny problem: CUDA context, cuBLAS/cuSOLVER handles, allocator
    warm = orc.make_params(iterations=1, knn_count=W["K"], max_dist=W["max_dist"], lr=W["lr"], svn_full_grad=W["svn_full_grad"])
    ref.scan(warm, src[:512], pb.target[:4096], pb.init_pose, pb.R0, pb.t0)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.time()
    out = ref.scan(prm, src, pb.target, pb.init_pose, pb.R0, pb.t0)
    torch.cuda.synchronize()
    wall = time.time() - t0
    sec = out["seconds"]
    if len(sys.argv) > 4:  # the reference's particles [6][P] for the full-size parity figure of bench.py
        np.save(sys.argv[4], out["particles"])
    print(json.dumps(dict(ok=True, particles=P, iterations=I, n_s=int(len(src)), n_t=int(len(pb.target)), seconds_scan=float(sec[0] + sec[1] + sec[2]),
                          seconds_add_cloud=float(sec[0]), seconds_stein_align=float(sec[1]), wall=wall,
                          peak_gb=torch.cuda.max_memory_allocated() / 1e9, mean=[float(v) for v in out["mean"]], gt=[float(v) for v in pb.gt_rel],
                          state=int(out["state"]))))


if __name__ == "__main__":
    try:
        main()
    except Exception as e:  # the reference's memory model: O(P * N_s) fp64 temporaries
        print(json.dumps(dict(ok=False, error=f"{type(e).__name__}: {str(e)[:300]}")))
