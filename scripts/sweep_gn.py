"""Time the per-phase cost of one configs[1] scan under the current env knobs."""
import sys, os, numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
pb = synth.make_problem_saturated(P, sensor="64")
icp = sv.SVNICP(sv.SteinICPParam(iterations=30, KNN_count=100, max_dist=3.0, lr=1.0, SVN_full_grad=True), pb.init_pose)
icp.set_profiling(True)
for _ in range(3):
    icp.add_cloud(pb.source, pb.target, pb.init_pose); icp.set_initial_mean(pb.R0, pb.t0); icp.stein_align()
ph = icp.get_phase_times(); t = icp.get_timing(); info = icp.get_scan_info()
print(os.environ.get("SVNICP_GN_STAGES"), os.environ.get("SVNICP_GN_SMEM_KB"), "TB", info["TB"], "gn_ms %.2f stein_ms %.2f setup %.2f total %.2f" % (ph["gn_ms"], ph["stein_ms"], ph["setup_ms"], t["total_ms"]))
ts = icp.get_tail_stamps()
print("tail phases us: decide %.1f | sync %.1f | median %.1f | stein %.1f | sync %.1f | update %.1f | sync %.1f | radius %.1f | sync %.1f  total %.1f" % tuple(list(np.diff(ts) / 1e3) + [(ts[-1] - ts[0]) / 1e3]))
