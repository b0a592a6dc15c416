import sys, numpy as np
sys.path.insert(0, '.')
import oracle as orc, svn_icp_b200 as sv
from svn_icp_b200 import synth
O = orc.Oracle()
np.set_printoptions(linewidth=200, precision=6)
lidar = synth.make_problem(64, sensor="32", scan_index=6, n_map_scans=6, seed=0xC0FFEE)
print("lidar n_s", len(lidar.source), "n_t", len(lidar.target))
# --- GN check
for P in (64,):
    rng = np.random.default_rng(P); init = synth.init_particles(P, rng)
    icp = sv.SVNICP(sv.SteinICPParam(iterations=1, KNN_count=100, max_dist=3.0, debug_corr=True), init)
    icp.add_cloud(lidar.source, lidar.target, init); icp.set_initial_mean(lidar.R0, lidar.t0); icp.stein_align()
    H, b, x = icp.get_gn_system()
    R = np.stack([O.so3_exp(init[3:, p])[0] for p in range(P)]); t = np.ascontiguousarray(init[:3].T)
    q0 = O.transform_q0(lidar.source, lidar.R0, lidar.t0)
    mink, _ = O.knn_mink(q0, lidar.target, 100)
    oH, ob, ridx, rmask = O.gn(R, t, lidar.R0, lidar.t0, lidar.source, lidar.target, mink, 3.0, want_corr=True)
    xf, idx, mask = icp.get_correspondences()
    print("idx mismatch frac", np.mean(ridx != idx), "mask mismatch", np.mean(rmask != mask), "masked frac", 1-mask.mean())
    for a in (slice(0,3), slice(3,6)):
        for c in (slice(0,3), slice(3,6)):
            sc = np.abs(oH[:, a, c]).max(); print("H blk", a, c, "scale", sc, "max rel err", np.abs(H[:, a, c]-oH[:, a, c]).max()/sc)
    print("b_t err", np.abs(b[:, :3]-ob[:, :3]).max(), "scale", np.abs(ob[:, :3]).max())
    print("b_r err", np.abs(b[:, 3:]-ob[:, 3:]).max(), "scale", np.abs(ob[:, 3:]).max())
    g = np.linalg.solve(H, b[..., None])[..., 0]; og = np.linalg.solve(oH, ob[..., None])[..., 0]
    print("newton err", np.abs(g-og).max(axis=0))
    p = np.argmax(np.abs(g-og).max(axis=1)); print("worst particle", p, "H diff\n", (H[p]-oH[p]), "\nb diff", b[p]-ob[p])
    print(icp.get_scan_info(), icp.get_timing(), icp.get_prune_stats())
# --- reuse problem
other = synth.make_uniform_problem(64, 700, 9000, seed=9)
icp = sv.SVNICP(sv.SteinICPParam(iterations=6, KNN_count=100, max_dist=3.0), other.init_pose)
icp.add_cloud(other.source, other.target, other.init_pose); icp.set_initial_mean(other.R0, other.t0); icp.stein_align()
print("fresh other: mean", icp.get_transformation(), "iters", icp.iterations_done(), icp.get_scan_info(), icp.get_prune_stats())
o = O.align(orc.make_params(iterations=6, knn_count=100, max_dist=3.0, lr=1.0), other.source, other.target, other.init_pose, other.R0, other.t0)
print("oracle other mean", o["mean"], "gt", other.gt_rel)
print("hist rows", np.abs(icp.get_particle_history()).sum(axis=1))
