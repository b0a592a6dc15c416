"""Scan pre-processing (crop_pointcloud + pcl::UniformSampling, OdometryPipeline.cpp:555-560, :684-704).
CPU: known-answer tests of the sequential restatement (oracle/preprocess_oracle.c; the down-sampling rule is restated from
PCL's published source -- PCL is absent here, so that part of the parity is unpinned and says so).
GPU: the device path (svnicp_pre_*) against the restatement: crop bit-exact in order, down-sampling bit-exact as a set."""
import numpy as np
import pytest

import oracle as orc


def sort_rows(a):
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))] if len(a) else a


@pytest.fixture(scope="module")
def po():
    return orc.PreprocessOracle()


def test_crop_known_answers(po):
    pts = np.array([[0.1, 0, 0], [1.0, 0, 0], [3, 4, 0], [60, 80, 0], [0, 0, 2]], dtype=np.float32)
    out, mx = po.crop(pts, 1.0, 100.0)
    # strict inequalities on both sides (OdometryPipeline.cpp:700): |p| = 1 and |p| = 100 are dropped; order kept
    np.testing.assert_array_equal(out, pts[[2, 4]])
    assert mx == 10000.0  # the SQUARED norm of the farthest input point (the reference's scan_max_range_ quirk, :699)


def test_uniform_known_answers(po):
    # leaf 0.5: inv = 2; points 0..2 share leaf ijk = (2, 0, 0); PCL compares the distance to the INDEX vector (2, 0, 0), in metres
    pts = np.array([[1.10, 0.1, 0.1], [1.40, 0.2, 0.3], [1.25, 0.0, 0.0], [5.0, 5.0, 5.0]], dtype=np.float32)
    out = po.downsample_uniform(pts, 0.5)
    assert len(out) == 2
    # distances to (2,0,0): 0.83, 0.49, 0.5625 -> the second point wins although the third is nearest to the leaf centre
    np.testing.assert_array_equal(sort_rows(out), sort_rows(pts[[1, 3]]))
    # ties keep the earlier point
    tie = np.array([[1.25, 0.25, 0.0], [1.25, 0.0, 0.25]], dtype=np.float32)
    np.testing.assert_array_equal(po.downsample_uniform(tie, 0.5), tie[:1])
    # negative coordinates floor toward -inf: (-0.1) and (0.1) are different leaves
    neg = np.array([[-0.1, 0, 0], [0.1, 0, 0]], dtype=np.float32)
    assert len(po.downsample_uniform(neg, 0.5)) == 2


def test_two_stage_reduction_sizes(po):
    from svn_icp_b200 import synth
    world = synth.make_world(3)
    pts, _ = synth.make_scan(world, 2, "16", 3)
    pts = pts.astype(np.float32)
    c, _ = po.crop(pts, 1.0, 100.0)
    a = po.downsample_uniform(c, 0.5)
    b = po.downsample_uniform(a, 1.5)
    assert len(pts) >= len(c) > len(a) > len(b) > 50
    # every output point is an input point
    S = {tuple(r) for r in c}
    assert all(tuple(r) in S for r in b)


@pytest.mark.gpu
def test_gpu_preprocess_matches_restatement(po):
    import svn_icp_b200 as sv
    from svn_icp_b200 import synth
    world = synth.make_world(0xC0FFEE)
    pts, (R, t) = synth.make_scan(world, 8, "64", 0xC0FFEE)
    pts = pts.astype(np.float32)
    pre = sv.ScanPreprocessor(len(pts))
    p1, n1 = pre.crop_pointcloud(pts, 1.0, 60.0)
    c_ref, mx = po.crop(pts, 1.0, 60.0)
    np.testing.assert_array_equal(pre.download(p1, n1), c_ref)      # same points, same order
    assert pre.scan_max_range_ == mx
    p2, n2 = pre.downsample_uniform(p1, 0.5, n=n1, on_device=True)  # voxelized_cloud_toMap (:559)
    a_ref = po.downsample_uniform(c_ref, 0.5)
    np.testing.assert_array_equal(sort_rows(pre.download(p2, n2)), sort_rows(a_ref))
    p3, n3 = pre.downsample_uniform(p2, 1.5, n=n2, on_device=True)  # voxelized_cloud (:560)
    # the second stage depends on the ORDER of its input only through exact distance ties: feed the restatement the device's order
    b_ref = po.downsample_uniform(pre.download(p2, n2), 1.5)
    np.testing.assert_array_equal(sort_rows(pre.download(p3, n3)), sort_rows(b_ref))
    assert n1 > n2 > n3 > 100
    # host input gives the same result as device input
    q, m = pre.downsample_uniform(c_ref, 0.5)
    np.testing.assert_array_equal(sort_rows(pre.download(q, m)), sort_rows(a_ref))


@pytest.mark.gpu
def test_gpu_preprocess_feeds_map_and_registration():
    """crop -> downsample(0.5) -> map.AddPointCloudDevice / downsample(1.5) -> to_f64 -> add_cloud_device: no host round trip."""
    import svn_icp_b200 as sv
    from svn_icp_b200 import synth
    world = synth.make_world(0xC0FFEE)
    local_map = sv.VoxelHashMap(1.0, 100.0, 20)
    pre = sv.ScanPreprocessor(200000)
    for k in range(4):
        pts, (R, t) = synth.make_scan(world, k, "32", 0xC0FFEE)
        p1, n1 = pre.crop_pointcloud(pts.astype(np.float32), 1.0, 100.0)
        p2, n2 = pre.downsample_uniform(p1, 0.5, n=n1, on_device=True)
        local_map.AddPointCloudDevice(p2, n2, False, R, t)
    pts, (Rg, tg) = synth.make_scan(world, 4, "32", 0xC0FFEE)
    p1, n1 = pre.crop_pointcloud(pts.astype(np.float32), 1.0, 100.0)
    p2, n2 = pre.downsample_uniform(p1, 0.5, n=n1, on_device=True)
    p3, n3 = pre.downsample_uniform(p2, 1.5, n=n2, on_device=True)
    src64 = pre.to_f64(p3, n3)
    tgt, n_t = local_map.GetMapDevice(tg, 110.0)
    rng = np.random.default_rng(5)
    init = synth.init_particles(48, rng)
    icp = sv.SVNICP(sv.SteinICPParam(iterations=20, KNN_count=32, max_dist=3.0, lr=1.0), init)
    icp.add_cloud_device(src64, n3, tgt, n_t, init)
    icp.set_initial_mean(Rg, tg)  # ground-truth guess: the correction must stay small
    assert icp.stein_align() == sv.ALIGN_SUCCESS
    m = icp.get_transformation()
    assert np.all(np.isfinite(m)) and np.linalg.norm(m[:3]) < 0.15 and np.linalg.norm(m[3:]) < 5e-3
