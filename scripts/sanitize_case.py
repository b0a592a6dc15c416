"""Small scan for compute-sanitizer: both kernel modes, odd N_s, early stop, P<32 and P>=32."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
for P, pair, full, es in ((5, 0, True, False), (40, 0, False, True), (70, 1, True, False)):
    if pair: os.environ["SVNICP_GN_PAIR"] = "1"
    else: os.environ.pop("SVNICP_GN_PAIR", None)
    pb = synth.make_uniform_problem(P, 301, 3000, seed=P)
    icp = sv.SVNICP(sv.SteinICPParam(iterations=4, KNN_count=37, max_dist=3.0, lr=1.0, SVN_full_grad=full, check_early_stop=es, convergence_threshold=1e-2, debug_corr=True), pb.init_pose)
    icp.add_cloud(pb.source, pb.target, pb.init_pose); icp.set_initial_mean(pb.R0, pb.t0); icp.stein_align()
    print(P, pair, icp.get_transformation()[:3], icp.iterations_done())
    icp.close()
