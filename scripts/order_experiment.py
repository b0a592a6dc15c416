"""Experiment: which particle -> thread order helps k_gn's warp-voted early exit most?  (configs[1] scan)
The library's own ordering is switched off (SVNICP_NO_PARTICLE_SORT=1) and the initial particles are permuted here by
different keys; the result set is the same up to order."""
import os
import sys
import numpy as np
os.environ["SVNICP_NO_PARTICLE_SORT"] = "1"
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth

P = 1000
pb = synth.make_problem_saturated(P, sensor="64")
I = 30


def run(init, label):
    icp = sv.SVNICP(sv.SteinICPParam(iterations=I, KNN_count=100, max_dist=3.0, lr=1.0, SVN_full_grad=True), init)
    icp.set_profiling(True)
    for _ in range(2):
        icp.add_cloud(pb.source, pb.target, init)
        icp.set_initial_mean(pb.R0, pb.t0)
        icp.stein_align()
    ph = icp.get_phase_times()
    print(f"{label:34s} gn_ms {ph['gn_ms']:.2f}", flush=True)
    icp.close()


def morton(x, comps, bits):
    q = [np.clip(((x[i] - x[i].min()) / (np.ptp(x[i]) + 1e-12) * (1 << bits)).astype(int), 0, (1 << bits) - 1) for i in comps]
    code = np.zeros(x.shape[1], dtype=np.int64)
    for b in range(bits):
        for j, qq in enumerate(q):
            code |= ((qq >> b) & 1) << (len(comps) * b + j)
    return np.argsort(code, kind="stable")


x = pb.init_pose
keys = {
    "as generated": np.arange(P),
    "morton (x,y,yaw) 3 bits [library]": morton(x, (0, 1, 5), 3),
    "morton (x,yaw) 4 bits": morton(x, (0, 5), 4),
    "morton (x,y,yaw,pitch) 2 bits": morton(x, (0, 1, 5, 4), 2),
    "morton (x,y,yaw,pitch,roll) 2 bits": morton(x, (0, 1, 5, 4, 3), 2),
    "morton (x,y,z,yaw) 2 bits": morton(x, (0, 1, 2, 5), 2),
    "morton (x,y,yaw) 2 bits": morton(x, (0, 1, 5), 2),
}
for label, perm in keys.items():
    run(np.ascontiguousarray(x[:, perm]), label)
