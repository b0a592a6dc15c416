"""Golden vectors for the SVGD-ICP class (class_type = SVGDICP), produced by the REFERENCE ITSELF
(oracle/_ref: the unmodified SVGDICP.cpp compiled against libtorch, CPU device swap).

Run in the build container only:
    python tests/golden/make_golden_svgd.py
The committed svgd_*.npz files hold inputs and reference outputs; tests never need /root/reference.
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

# name -> (P, n_s, n_t, K, iterations, optimizer, early_stop, threshold, lr, max_dist, seed, stale, two_scans)
#   stale: the constructor sees a DIFFERENT particle set than add_cloud (pose_particles_ is not refreshed by
#   add_cloud, SVGDICP.cpp:46-62, so iteration 0 uses the constructor's); two_scans: a second add_cloud+align.
CASES = {
    "svgd_adam_p20": (20, 300, 3000, 16, 8, "Adam", 0, 1e-5, 0.03, 3.0, 21, 0, 0),
    "svgd_rmsprop_p12": (12, 300, 3000, 16, 8, "RMSprop", 0, 1e-5, 0.01, 3.0, 22, 0, 0),
    "svgd_sgd_p8": (8, 300, 3000, 16, 8, "SGD", 0, 1e-5, 0.004, 3.0, 23, 0, 0),
    "svgd_adagrad_p8": (8, 300, 3000, 16, 8, "Adagrad", 0, 1e-5, 0.03, 3.0, 24, 0, 0),
    "svgd_p1_adam": (1, 300, 3000, 16, 6, "Adam", 0, 1e-5, 0.03, 3.0, 25, 0, 0),
    "svgd_p2_adam": (2, 200, 2000, 8, 4, "Adam", 0, 1e-5, 0.03, 1.0, 26, 0, 0),          # bandwidth 0 -> NaN
    "svgd_stale_p10": (10, 300, 3000, 16, 6, "Adam", 0, 1e-5, 0.03, 3.0, 27, 1, 1),
    "svgd_earlystop_p8": (8, 300, 3000, 16, 60, "Adam", 1, 1.2e-2, 0.03, 3.0, 28, 0, 0),
    "svgd_noopt_p4": (4, 100, 1000, 8, 3, "LBFGS", 0, 1e-5, 0.03, 3.0, 29, 1, 0),
    "svgd_small_map_p5": (5, 60, 10, 16, 3, "Adam", 0, 1e-5, 0.03, 3.0, 30, 0, 0),
}


def run_case(name: str) -> None:
    sys.path.insert(0, ROOT)
    import oracle as orc
    from svn_icp_b200 import synth

    P, n_s, n_t, K, I, opt, es, thr, lr, md, seed, stale, two = CASES[name]
    pb = synth.make_uniform_problem(P, n_s, max(n_t, n_s), seed=seed, box=8.0)
    if n_t < n_s:
        pb.target = pb.target[:n_t].copy()
    rng = np.random.default_rng(seed + 1000)
    ctor_pose = pb.init_pose + (rng.normal(size=pb.init_pose.shape) * 0.05 if stale else 0.0)
    init2 = pb.init_pose + rng.normal(size=pb.init_pose.shape) * 0.02 if two else None
    prm = orc.make_params(iterations=I, lr=lr, max_dist=md, check_early_stop=bool(es), convergence_threshold=thr, knn_count=K)
    ref = orc.Reference()
    out = ref.svgd_scan(prm, opt, pb.source, pb.target, ctor_pose, pb.init_pose, pb.R0, pb.t0, init_pose2=init2)
    first = ref.svgd_scan(prm, opt, pb.source, pb.target, ctor_pose, pb.init_pose, pb.R0, pb.t0) if two else out
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        source=pb.source, target=pb.target, init_pose=pb.init_pose, ctor_pose=ctor_pose,
        init_pose2=init2 if init2 is not None else np.zeros(0), R0=pb.R0, t0=pb.t0,
        params=np.array([I, lr, md, es, thr, K], dtype=np.float64), optimizer=np.array(opt),
        ref_particles=out["particles"], ref_mean=out["mean"], ref_var=out["var"], ref_cov=out["cov"],
        ref_weights=out["weights"], ref_history=out["history"], ref_state=np.array([out["state"]]),
        ref_runtime=out["runtime"], ref_first_particles=first["particles"],
    )
    print(name, "ok; state", out["state"], "finish_iter", out["runtime"][2], "mean", out["mean"])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
        for n in CASES:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), n])
