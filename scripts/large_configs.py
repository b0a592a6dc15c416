"""Sanity + timing of the larger BASELINE.json configs on one GPU (parity-test cases, not bench lines).
usage: python scripts/large_configs.py c3|c5|c1"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
rng = np.random.default_rng(7)
if which == "c3":      # 128-beam ~260k points, 4096 particles
    pb = synth.make_problem_saturated(4096, sensor="128"); I = 10
elif which == "c5":    # dense-map stress: 10M-point map (0.25 m voxels over the corridor), 16384 particles
    world = synth.make_world(); 
    src, (Rg, tg) = synth.make_scan(world, 8, "64")
    t0 = time.time(); tgt = synth.saturated_map(world, tg, 100.0, rng, voxel=0.25, cap=20, density=900.0); print("map gen s", time.time()-t0, len(tgt))
    pb = synth.make_problem_saturated(16384, sensor="64"); pb.target = tgt; I = 4
else:                  # configs[0]: 100 particles
    pb = synth.make_problem_saturated(100, sensor="64"); I = 30
P = pb.init_pose.shape[1]
print(which, "P", P, "n_s", len(pb.source), "n_t", len(pb.target), "I", I, flush=True)
icp = sv.SVNICP(sv.SteinICPParam(iterations=I, KNN_count=100, max_dist=3.0, lr=1.0, SVN_full_grad=True), pb.init_pose)
icp.set_profiling(True)
for rep in range(2):
    icp.add_cloud(pb.source, pb.target, pb.init_pose); icp.set_initial_mean(pb.R0, pb.t0)
    assert icp.stein_align() == 1
a = icp.get_particles().copy()
print("timing", icp.get_timing(), "phases", icp.get_phase_times(), "info", icp.get_scan_info())
print("mean", icp.get_transformation(), "gt", pb.gt_rel, "prune", np.round(icp.get_prune_stats(), 1))
icp.add_cloud(pb.source, pb.target, pb.init_pose); icp.set_initial_mean(pb.R0, pb.t0); icp.stein_align()
print("deterministic:", np.array_equal(a, icp.get_particles()), "finite:", np.isfinite(a).all())
err = np.abs(icp.get_transformation() - pb.gt_rel)
print("registration error", err)
