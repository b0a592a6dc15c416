"""Experiment: does the particle -> thread order matter for k_gn's warp-voted early exit?  (configs[1] scan)
Permute the initial particles by different keys; the result set is the same up to order."""
import sys
import numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth

P = 1000
pb = synth.make_problem_saturated(P, sensor="64")
I = int(sys.argv[1]) if len(sys.argv) > 1 else 30


def run(init, label, cls=sv.SVNICP, **kw):
    prm = sv.SteinICPParam(iterations=I, KNN_count=100, max_dist=3.0, **kw)
    icp = cls(prm, init)
    icp.set_profiling(True)
    for _ in range(2):
        icp.add_cloud(pb.source, pb.target, init)
        icp.set_initial_mean(pb.R0, pb.t0)
        icp.stein_align()
    ph = icp.get_phase_times()
    print(f"{label:28s} gn_ms {ph['gn_ms']:.2f} filter {ph['filter_ms']:.2f} mean {np.round(icp.get_transformation(), 5)}", flush=True)
    icp.close()


x = pb.init_pose
keys = {
    "as generated": np.arange(P),
    "sorted by yaw": np.argsort(x[5]),
    "sorted by x": np.argsort(x[0]),
    "sorted by |t|": np.argsort(np.linalg.norm(x[:3], axis=0)),
    "morton (x,y,yaw) 3 bits": None,
}
q = [np.clip(((x[i] - x[i].min()) / (np.ptp(x[i]) + 1e-12) * 8).astype(int), 0, 7) for i in (0, 1, 5)]
code = np.zeros(P, dtype=int)
for b in range(3):
    for j, qq in enumerate(q):
        code |= ((qq >> b) & 1) << (3 * b + j)
keys["morton (x,y,yaw) 3 bits"] = np.argsort(code, kind="stable")
for label, perm in keys.items():
    run(np.ascontiguousarray(x[:, perm]), "svn " + label, lr=1.0, SVN_full_grad=True)
for label in ("as generated", "morton (x,y,yaw) 3 bits"):
    run(np.ascontiguousarray(x[:, keys[label]]), "svgd " + label, cls=sv.SVGDICP, lr=0.03, optimizer="Adam")
