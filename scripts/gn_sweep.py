"""k_gn tuning sweep on one GPU: python scripts/gn_sweep.py  (P, stages, smem KB) -> per-scan phase times at the configs[1] cloud."""
import sys, numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
pb = synth.make_problem_saturated(1000, sensor="64")
rng = np.random.default_rng(5)
for P, stages, kb in [(1000, 0, 0), (500, 0, 0), (250, 0, 0), (125, 0, 0), (100, 0, 0), (30, 0, 0)]:
    init = synth.init_particles(P, rng)
    icp = sv.SVNICP(sv.SteinICPParam(iterations=30, KNN_count=100, max_dist=3.0, lr=1.0, gn_stages=stages, gn_smem_kb=kb), init)
    icp.set_profiling(True)
    for _ in range(3):
        icp.add_cloud(pb.source, pb.target, init); icp.set_initial_mean(pb.R0, pb.t0); icp.stein_align()
    ph = icp.get_phase_times(); info = icp.get_scan_info(); t = icp.get_timing()
    print(f"P={P} stages={stages} kb={kb} TB={info['TB']} slices={info['n_slices']}: gn {ph['gn_ms']:.2f} filter {ph['filter_ms']:.2f} fin {ph['finalize_ms']:.2f} tail {ph['stein_ms']:.2f} total {t['total_ms']:.2f} ms", flush=True)
    icp.close()
