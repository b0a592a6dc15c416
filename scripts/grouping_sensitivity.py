"""How far do two runs of the SAME scan drift apart when only the grouping of the fp32 Gauss-Newton partial sums changes
(tile height / ring depth of k_gn: ~1e-8 relative differences in H and b)?  Calibrates the sharded-vs-single tolerance.
python scripts/grouping_sensitivity.py"""
import sys, numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
pb = synth.make_problem_saturated(1000, sensor="64")
res = {}
for name, st, kb in (("default TB=16 S=4", 0, 0), ("TB=32 S=2", 2, 110), ("TB=16 S=5", 5, 140), ("TB=8 S=4", 4, 60)):
    icp = sv.SVNICP(sv.SteinICPParam(iterations=30, KNN_count=100, max_dist=3.0, lr=1.0, SVN_full_grad=True, gn_stages=st, gn_smem_kb=kb), pb.init_pose)
    icp.add_cloud(pb.source, pb.target, pb.init_pose); icp.set_initial_mean(pb.R0, pb.t0)
    assert icp.stein_align() == sv.ALIGN_SUCCESS
    res[name] = (icp.get_particles().reshape(6, -1).copy(), icp.get_particle_history().copy(), icp.get_transformation().copy(), icp.get_scan_info()["TB"])
    icp.close()
base = res["default TB=16 S=4"]
for name, (p, h, m, tb) in res.items():
    d = np.abs(p - base[0]).max(axis=0)
    dh = np.abs(h.reshape(h.shape[0], -1) - base[1].reshape(h.shape[0], -1)).max(axis=1)
    print(f"{name} (TB={tb}): max |particles - default| {d.max():.3e} (particles beyond 1e-5: {(d > 1e-5).sum()}, beyond 1e-4: {(d > 1e-4).sum()}), "
          f"history max {dh.max():.3e} at iteration {int(dh.argmax())}, mean {np.abs(m - base[2]).max():.3e}", flush=True)
