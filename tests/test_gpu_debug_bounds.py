"""The -DSVN_DEBUG_BOUNDS build (device-side bounds checks in k_filter*, k_gn, k_finalize, k_tail; compute-sanitizer is not
available on the target pool): the same scans must run clean -- a failed check makes svnicp_align return an error -- and give
the same result as the release build."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, %r)
import svn_icp_b200 as sv
from svn_icp_b200 import synth
out = []
for P, full, es, K in ((37, True, False, 40), (130, False, True, 100), (1, True, False, 17), (600, True, False, 64)):
    pb = synth.make_problem(P, sensor="32", scan_index=6, n_map_scans=4, seed=0xC0FFEE)
    icp = sv.SVNICP(sv.SteinICPParam(iterations=8, KNN_count=K, max_dist=3.0, lr=1.0, SVN_full_grad=full, check_early_stop=es), pb.init_pose)
    for _ in range(2):
        icp.add_cloud(pb.source[::3], pb.target, pb.init_pose)
        icp.set_initial_mean(pb.R0, pb.t0)
        assert icp.stein_align() == sv.ALIGN_SUCCESS
    out.append(icp.get_particles().tolist())
    icp.close()
print(json.dumps(out))
"""


def _run(lib):
    env = dict(os.environ)
    if lib:
        env["SVNICP_B200_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [np.asarray(x) for x in json.loads(r.stdout.strip().splitlines()[-1])]


def test_debug_bounds_build_runs_clean_and_matches_release():
    from svn_icp_b200 import build as b
    dbg = b.LIB.replace(".so", "_dbg.so")
    if not os.path.exists(dbg):  # built by __graft_entry__.build(); build here when the test runs on its own
        dbg = b.build(debug_bounds=True)
    got, ref = _run(dbg), _run(None)
    for g, r in zip(got, ref):
        np.testing.assert_array_equal(g, r)  # the checks do not touch the arithmetic
