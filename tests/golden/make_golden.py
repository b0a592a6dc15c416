"""Generate golden vectors by running the REFERENCE ITSELF (oracle/_ref, built by oracle/build_ref.sh
from the unmodified sources under /root/reference) on small seeded problems.

Run in the build container only (needs /root/reference for the build):
    python tests/golden/make_golden.py
Each case runs in a fresh subprocess because the reference freezes the particle count and the
early-stop threshold in function-static tensors (SVNICP.cpp:42,167).  The committed .npz files hold
the inputs AND the reference's outputs, so tests never need /root/reference.
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

# name -> (P, n_s, n_t, K, iterations, svn_full_grad, early_stop, threshold, lr, max_dist, seed)
CASES = {
    "svn_full_p24": (24, 300, 3000, 16, 6, 1, 0, 1e-5, 1.0, 3.0, 11),
    "svn_precond_p24": (24, 300, 3000, 16, 6, 0, 0, 1e-5, 1.0, 3.0, 12),
    "svn_p1": (1, 300, 3000, 16, 5, 1, 0, 1e-5, 1.0, 3.0, 13),
    # P == 2: the lower median of {0,0,d,d} is 0 -> bandwidth 0 -> NaN translations.  Reference behaviour, kept.
    "svn_p2_lr05": (2, 200, 2000, 8, 4, 1, 0, 1e-5, 0.5, 1.0, 14),
    "svn_p3_lr05": (3, 200, 2000, 8, 4, 1, 0, 1e-5, 0.5, 1.0, 17),
    # threshold chosen so that the stop fires mid-run (epoch 5): rows 5.. of the history stay zero
    "svn_earlystop_p8": (8, 300, 3000, 16, 40, 0, 1, 1.5e-2, 1.0, 3.0, 15),
    "svn_earlystop_never_p8": (8, 300, 3000, 16, 12, 1, 1, 5e-4, 1.0, 3.0, 15),
    "svn_small_map_p5": (5, 60, 10, 16, 3, 1, 0, 1e-5, 1.0, 3.0, 16),  # N_t < K: zero padded candidates
}


def run_case(name: str) -> None:
    sys.path.insert(0, ROOT)
    import oracle as orc
    from svn_icp_b200 import synth

    P, n_s, n_t, K, I, full, es, thr, lr, md, seed = CASES[name]
    pb = synth.make_uniform_problem(P, n_s, max(n_t, n_s), seed=seed, box=8.0)
    if n_t < n_s:  # tiny map (N_t < K): keep the first n_t map points
        pb.target = pb.target[:n_t].copy()
    prm = orc.make_params(iterations=I, lr=lr, max_dist=md, check_early_stop=bool(es), convergence_threshold=thr,
                          knn_count=K, svn_full_grad=bool(full))
    ref = orc.Reference()
    full_run = ref.scan(prm, pb.source, pb.target, pb.init_pose, pb.R0, pb.t0)
    steps = min(I, 4)
    st = ref.scan_steps(prm, pb.source, pb.target, pb.init_pose, pb.R0, pb.t0, steps)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        source=pb.source, target=pb.target, init_pose=pb.init_pose, R0=pb.R0, t0=pb.t0,
        params=np.array([I, lr, md, es, thr, K, full], dtype=np.float64),
        ref_particles=full_run["particles"], ref_mean=full_run["mean"], ref_var=full_run["var"],
        ref_cov=full_run["cov"], ref_weights=full_run["weights"], ref_history=full_run["history"],
        ref_state=np.array([full_run["state"]]),
        step_x_after=st["x_after"], step_H=st["H"], step_b=st["b"], step_tgt_paired=st["tgt_paired"].astype(np.float64),
        step_cand_idx=st["cand_idx"],
    )
    print(name, "ok; mean", full_run["mean"])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
        for n in CASES:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), n])
