"""k_gn launch-shape check on ONE GPU: the Gauss-Newton system (H, b) of iteration 0 depends only on a particle's own pose, so a
handle that holds a SLICE of the particles (other PG / RG shape of k_gn, as a sharded rank would run) must reproduce the rows
of the full handle up to the fp32 summation grouping.  python scripts/gn_shape_check.py [P_total]"""
import sys, numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
pb = synth.make_problem_saturated(P, sensor="64")
prm = sv.SteinICPParam(iterations=1, KNN_count=100, max_dist=3.0, lr=1.0, SVN_full_grad=True, flags=sv.FLAG_NO_PARTICLE_SORT)
def system(init):
    icp = sv.SVNICP(prm, init)
    icp.add_cloud(pb.source, pb.target, init); icp.set_initial_mean(pb.R0, pb.t0)
    assert icp.stein_align() == sv.ALIGN_SUCCESS
    H, b, x = icp.get_gn_system(); info = icp.get_scan_info(); icp.close()
    return H, b, info
init = pb.init_pose  # [6, P]
Hf, bf, info = system(init)
print("full", info, flush=True)
scale_H, scale_b = np.abs(Hf).max(), np.abs(bf).max()
for lo, hi in ((0, 125), (875, 1000), (0, 63), (0, 64), (0, 65), (100, 350), (0, 500), (3, 4), (10, 40), (0, 129)):
    Hs, bs, info = system(np.ascontiguousarray(init[:, lo:hi]))
    eh, eb = np.abs(Hs - Hf[lo:hi]).max() / scale_H, np.abs(bs - bf[lo:hi]).max() / scale_b
    worst = int(np.abs(bs - bf[lo:hi]).max(axis=1).argmax())
    print(f"slice [{lo},{hi}) TB={info['TB']} pgroups={info['n_pgroups']} slices={info['n_slices']}: rel |dH| {eh:.2e} rel |db| {eb:.2e} worst particle {worst} -> {'ok' if eh < 1e-5 and eb < 1e-5 else 'MISMATCH'}", flush=True)
