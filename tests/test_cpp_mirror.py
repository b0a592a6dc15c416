"""The header-only C++ mirror of the reference classes (svn_icp_b200/include/svnicp/SVNICP.hpp): compiles against the C ABI
(CPU check) and, on a GPU, produces the same bits as the Python mirror."""
import os
import struct
import subprocess

import numpy as np
import pytest

import svn_icp_b200 as sv
from svn_icp_b200 import synth
from conftest import ROOT


def build_mirror(tmp_path):
    from svn_icp_b200 import build
    lib = build.build()
    exe = os.path.join(tmp_path, "mirror_main")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "svn_icp_b200", "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_main.cpp"), "-o", exe, lib,
                           "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


def test_mirror_compiles_and_links(tmp_path):
    exe = build_mirror(str(tmp_path))
    assert os.path.exists(exe)
    # plain C consumers can include the ABI header too
    c = os.path.join(tmp_path, "abi.c")
    open(c, "w").write('#include "svnicp_b200.h"\nint main(void){svnicp_params p; svnicp_default_params(&p); return p.KNN_count==100?0:1;}\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(tmp_path, "abi"),
                           sv.LIB_PATH, "-Wl,-rpath," + os.path.dirname(sv.LIB_PATH)])
    assert subprocess.call([os.path.join(tmp_path, "abi")]) == 0


def test_mirror_batch_and_handoff_compile_and_run(tmp_path):
    """SVNICPBatch / compose_pose / ScanPreprocessor::deskew_pointcloud of the C++ mirror compile against the C ABI; compose_pose
    is host arithmetic and runs here, the batch constructor must fail loudly without a GPU (no CPU fallback)."""
    from svn_icp_b200 import build
    lib = build.build()
    src = os.path.join(tmp_path, "check.cpp")
    open(src, "w").write("""
#include "svnicp/SVNICP.hpp"
#include "svnicp/ScanPreprocessor.hpp"
#include "svnicp/VoxelHashMap.hpp"
int main() {
  svnicp::InitialMean g;
  auto q = svnicp::compose_pose(g, std::vector<double>{0.1, 0.2, 0.3, 0.01, 0.02, 0.03});
  if (q.t[0] != 0.1 || q.t[2] != 0.3) return 1;
  int devs = 0;
  svnicp::SteinICPParam p;
  std::vector<double> ip(2 * 6 * 4, 0.0);
  try {
    svnicp::SVNICPBatch b(p, 2, 4, ip);
    if (b.size() != 2) return 2;
    devs = 1;
  } catch (const svnicp::Error &) {
  }
  return devs ? 10 : 0;
}
""")
    exe = os.path.join(tmp_path, "check")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "svn_icp_b200", "include"), src, "-o", exe, lib,
                           "-Wl,-rpath," + os.path.dirname(lib)])
    import torch
    assert subprocess.call([exe]) == (10 if torch.cuda.is_available() else 0)


def test_stein_msgs_producer_mapping(tmp_path):
    """stein_msgs field mapping (svnicp/stein_msgs_compat.hpp): [6][P] slices -> x,y,z,roll,pitch,yaw, as
    OdometryPipeline.cpp:942-987 does.  Pure host code: runs without a GPU."""
    from svn_icp_b200 import build
    lib = build.build()
    exe = os.path.join(tmp_path, "msgs_main")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "svn_icp_b200", "include"),
                           os.path.join(ROOT, "tests", "cpp", "msgs_main.cpp"), "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)])
    assert subprocess.call([exe]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("cls", ["SVNICP", "SVGDICP"])
def test_mirror_matches_python_mirror(tmp_path, cls):
    exe = build_mirror(str(tmp_path))
    P, I, K = 48, 7, 40
    pb = synth.make_uniform_problem(P, 900, 9000, seed=4)
    prob, res = os.path.join(tmp_path, "problem.bin"), os.path.join(tmp_path, "result.bin")
    with open(prob, "wb") as f:
        f.write(struct.pack("6q", len(pb.source), len(pb.target), P, I, K, 1))
        f.write(struct.pack("2d", 1.0 if cls == "SVNICP" else 0.03, 3.0))
        for a in (pb.source, pb.target, pb.init_pose, pb.R0, pb.t0):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    out = subprocess.run([exe, prob, res, cls], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = np.fromfile(res, dtype=np.uint8)
    d = np.frombuffer(raw[: 8 * (48 + 7 * P)].tobytes(), dtype=np.float64)
    hist = np.frombuffer(raw[8 * (48 + 7 * P):].tobytes(), dtype=np.float32).reshape(I, 6 * P)
    if cls == "SVNICP":
        icp = sv.SVNICP(sv.SteinICPParam(iterations=I, KNN_count=K, max_dist=3.0, lr=1.0, SVN_full_grad=True), pb.init_pose)
    else:
        icp = sv.SVGDICP(sv.SteinICPParam(iterations=I, KNN_count=K, max_dist=3.0, lr=0.03, SVN_full_grad=True), pb.init_pose)
    icp.add_cloud(pb.source, pb.target, pb.init_pose)
    icp.set_initial_mean(pb.R0, pb.t0)
    icp.stein_align()
    np.testing.assert_array_equal(d[:6], icp.get_transformation())
    np.testing.assert_array_equal(d[6:12], icp.get_distribution())
    np.testing.assert_array_equal(d[12:48], icp.get_cov_matrix())
    np.testing.assert_array_equal(d[48:48 + 6 * P], icp.get_particles())
    np.testing.assert_array_equal(d[48 + 6 * P:], icp.get_particle_weight())
    np.testing.assert_array_equal(hist, icp.get_particle_history())


@pytest.mark.gpu
def test_mirror_local_map_feeds_registration(tmp_path):
    """svnicp::VoxelHashMap (C++ mirror) -> GetMap(pose, range) device cloud -> add_cloud, as the node does."""
    exe = build_mirror(str(tmp_path))
    P, I, K = 32, 5, 24
    pb = synth.make_uniform_problem(P, 600, 6000, seed=9)
    prob, res = os.path.join(tmp_path, "problem.bin"), os.path.join(tmp_path, "result.bin")
    with open(prob, "wb") as f:
        f.write(struct.pack("6q", len(pb.source), len(pb.target), P, I, K, 1))
        f.write(struct.pack("2d", 1.0, 3.0))
        for a in (pb.source, pb.target, pb.init_pose, pb.R0, pb.t0):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    out = subprocess.run([exe, prob, res, "SVNICP", "MAP"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    d = np.fromfile(res, dtype=np.uint8)[: 8 * 48].view(np.float64)
    # same thing through the Python mirror: the map stores float32 points, so the target is the float32-rounded cloud
    m = sv.VoxelHashMap(1.0, 1e6, 32, capacity_voxels=1 << 16)
    m.AddPointCloud(pb.target.astype(np.float32), np.eye(3), np.zeros(3))
    tgt = m.GetMap(pb.t0, 1e5)
    assert len(tgt) == len(pb.target)  # cap 32 was enough for every voxel of this problem
    icp = sv.SVNICP(sv.SteinICPParam(iterations=I, KNN_count=K, max_dist=3.0, lr=1.0, SVN_full_grad=True), pb.init_pose)
    icp.add_cloud(pb.source, tgt, pb.init_pose)
    icp.set_initial_mean(pb.R0, pb.t0)
    icp.stein_align()
    # point ORDER of the two map instances may differ (hash slots); it only matters for exact distance ties
    np.testing.assert_allclose(d[:6], icp.get_transformation(), atol=1e-12, rtol=0)
