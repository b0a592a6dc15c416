"""How often does k_filter prune from the previous iteration's list (ball containment), and what does it save?
Run twice: default, and with SVNICP_FILTER_FULL=1 in the environment (always the full K-slot table)."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
pb = synth.make_problem_saturated(1000, sensor="64")
icp = sv.SVNICP(sv.SteinICPParam(iterations=30, KNN_count=100, max_dist=3.0, lr=1.0, SVN_full_grad=True), pb.init_pose)
icp.set_profiling(True)
for _ in range(3):
    icp.add_cloud(pb.source, pb.target, pb.init_pose); icp.set_initial_mean(pb.R0, pb.t0); icp.stein_align()
ph = icp.get_phase_times()
os.environ["SVNICP_DEBUG_REUSE"] = "1"
hits = icp.get_prune_stats()
del os.environ["SVNICP_DEBUG_REUSE"]
kept = icp.get_prune_stats()
print(f"full={os.environ.get('SVNICP_FILTER_FULL')} filter {ph['filter_ms']:.2f} gn {ph['gn_ms']:.2f} total {icp.get_timing()['total_ms']:.2f} | hit% " +
      " ".join(f"{100*h:.0f}" for h in hits[::2]) + " ... | kept " + " ".join(f"{k:.1f}" for k in kept[::6]), flush=True)
