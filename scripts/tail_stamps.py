"""Phase stamps of k_tail (tuning aid): python scripts/tail_stamps.py [P ...]"""
import sys, numpy as np
sys.path.insert(0, '.')
import svn_icp_b200 as sv
from svn_icp_b200 import synth
pb0 = synth.make_problem_saturated(16, sensor="64")
for P in [int(a) for a in sys.argv[1:]] or [1000, 256, 125, 30]:
    rng = np.random.default_rng(P)
    init = synth.init_particles(P, rng)
    for full in (True, False):
        icp = sv.SVNICP(sv.SteinICPParam(iterations=12, KNN_count=100, max_dist=3.0, lr=1.0, SVN_full_grad=full), init)
        icp.set_profiling(True)
        for _ in range(2):
            icp.add_cloud(pb0.source[::8], pb0.target, init); icp.set_initial_mean(pb0.R0, pb0.t0); icp.stein_align()
        s = icp.get_tail_stamps(); ph = icp.get_phase_times()
        d = (s[1:8] - s[0]) / 1e3
        print(f"P={P} full={full}: k_tail stamps us after-wait {d[0]:.1f} stein-sums {d[1]:.1f} [solve {d[5]:.1f} exp+update+log {d[6]:.1f}] update {d[2]:.1f} lastCTA-start {d[3]:.1f} end {d[4]:.1f} | "
              f"per-iteration us: filter {ph['filter_ms']/12*1e3:.1f} gn {ph['gn_ms']/12*1e3:.1f} finalize {ph['finalize_ms']/12*1e3:.1f} tail {ph['stein_ms']/12*1e3:.1f} total {icp.get_timing()['iterations_ms']/12*1e3:.1f}")
