"""One configs[1] scan for ncu: python scripts/profile_scan.py [P] [iterations] [n_scans]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
I = int(sys.argv[2]) if len(sys.argv) > 2 else 30
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pb = synth.make_problem_saturated(P, sensor="64")
icp = sv.SVNICP(sv.SteinICPParam(iterations=I, KNN_count=100, max_dist=3.0, lr=1.0, SVN_full_grad=True), pb.init_pose)
for _ in range(N):
    icp.add_cloud(pb.source, pb.target, pb.init_pose)
    icp.set_initial_mean(pb.R0, pb.t0)
    icp.stein_align()
print(icp.get_timing(), icp.get_transformation(), pb.gt_rel)
