"""The shipped regime (geodeAlpha.yaml:9-22: 30 particles, <= 100 iterations, early stop 5e-4, pre-conditioned step, down-sampled
source) and configs[0] for timing / ncu launch lists: python scripts/small_regime.py [graph|direct] [n_scans]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
fl = sv.FLAG_NO_GRAPH if mode == "direct" else 0
pb = synth.make_problem_saturated(30, sensor="64")
ds = np.ascontiguousarray(synth.uniform_downsample(synth.uniform_downsample(pb.source, 0.5), 1.5))
rng = np.random.default_rng(3)
import torch
for name, P, kw in (("shipped_p30", 30, dict(iterations=100, SVN_full_grad=False, check_early_stop=True, convergence_threshold=5e-4)),
                    ("config0_p100", 100, dict(iterations=30, SVN_full_grad=True))):
    init = synth.init_particles(P, rng)
    icp = sv.SVNICP(sv.SteinICPParam(KNN_count=100, max_dist=3.0, lr=1.0, flags=fl, **kw), init)
    for i in range(N + 3):
        if i == 3:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        icp.add_cloud(ds, pb.target, init); icp.set_initial_mean(pb.R0, pb.t0); icp.stein_align()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / N
    print(f"{name} {mode}: {dt*1e3:.3f} ms per scan (host clock, incl. add_cloud H2D), device {icp.get_timing()}, iterations {icp.iterations_done()}, launches {icp.launch_count()}", flush=True)
    icp.close()
