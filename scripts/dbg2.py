import sys, numpy as np
sys.path.insert(0, '.')
import oracle as orc, svn_icp_b200 as sv
from svn_icp_b200 import synth
O = orc.Oracle()
np.set_printoptions(linewidth=200, precision=6)
other = synth.make_uniform_problem(64, 700, 9000, seed=9)
P = 64
for I in (1, 2, 3):
    icp = sv.SVNICP(sv.SteinICPParam(iterations=I, KNN_count=100, max_dist=3.0, debug_corr=True), other.init_pose)
    icp.add_cloud(other.source, other.target, other.init_pose); icp.set_initial_mean(other.R0, other.t0); icp.stein_align()
    o = O.align(orc.make_params(iterations=I, knn_count=100, max_dist=3.0, lr=1.0), other.source, other.target, other.init_pose, other.R0, other.t0,
                dumps=("H", "b", "delta", "x_before", "x_after", "bandwidth", "corr_idx", "corr_mask"))
    H, b, x = icp.get_gn_system(); d, h = icp.get_stein()
    xf, idx, mask = icp.get_correspondences()
    print(f"--- I={I}: last iteration compare")
    print(" x_before err", np.abs(x - o["x_before"][-1]).max())
    print(" H rel err", np.abs(H - o["H"][-1]).max() / np.abs(o["H"][-1]).max(), " b err", np.abs(b - o["b"][-1]).max(), "b scale", np.abs(o["b"][-1]).max())
    print(" idx mismatch", np.mean(idx != o["corr_idx"][-1]), "mask mismatch", np.mean(mask != o["corr_mask"][-1]), "masked", 1 - mask.mean())
    print(" bandwidth gpu", h, "oracle", o["bandwidth"][-1])
    print(" delta err", np.abs(d - o["delta"][-1]).max(), "delta scale", np.abs(o["delta"][-1]).max(), "gpu delta scale", np.abs(d).max())
    print(" particles err", np.abs(icp.get_particles().reshape(6, P) - o["particles"]).max())
    print(" prune", icp.get_prune_stats())
