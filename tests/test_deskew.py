"""De-skewing (OdometryPipeline::deskew_pointcloud, OdometryPipeline.cpp:357-447) and the ICP-mode pose hand-off
(updater_, :37-45 + tensor2gtsamPose3, ICPUtils.cpp:84-98).

The per-point arithmetic of the reference lives in GTSAM (absent here, version unpinned by the reference's CMake): the C
restatement of GTSAM 4.2's closed forms is pinned against the mathematical definition -- scipy's matrix exponential /
logarithm of the 4x4 twist -- on CPU; the CUDA kernels are then compared with the restatement on the GPU.
"""
import numpy as np
import pytest
from scipy.linalg import expm, logm

import oracle as orc
import svn_icp_b200 as sv
from svn_icp_b200 import synth


def twist_matrix(xi):
    w, v = xi[:3], xi[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = v
    return M


@pytest.fixture(scope="module")
def pre_oracle():
    return orc.PreprocessOracle()


@pytest.mark.parametrize("scale", [1e-12, 1e-6, 1e-3, 0.05, 0.7, 2.5])
def test_pose3_expmap_logmap_against_the_matrix_exponential(pre_oracle, scale):
    rng = np.random.default_rng(int(scale * 1e6) + 1)
    for _ in range(20):
        xi = rng.normal(size=6) * np.r_[scale, scale, scale, 1.0, 1.0, 1.0]
        if np.linalg.norm(xi[:3]) > 3.0:  # keep the angle below pi so that Logmap(Expmap(xi)) == xi
            xi[:3] *= 3.0 / np.linalg.norm(xi[:3])
        R, t = pre_oracle.pose3_expmap(xi)
        T = expm(twist_matrix(xi))
        np.testing.assert_allclose(R, T[:3, :3], atol=1e-13, rtol=0)
        # GTSAM's closed form t = (w x v - R (w x v) + w (w.v)) / theta^2 cancels for small angles (relative error ~ eps / theta),
        # and its near-zero branch is first order (|w||v|/2): both are properties of the restated formula, not of the restatement
        tol = 1e-11 + 2e-15 / scale if scale > 1e-7 else 1e-11
        np.testing.assert_allclose(t, T[:3, 3], atol=tol, rtol=0)
        back = pre_oracle.pose3_logmap(R, t)
        np.testing.assert_allclose(back, xi, atol=max(1e-9, 10 * tol), rtol=0)
        if scale > 1e-3:
            L = np.real(logm(T))
            np.testing.assert_allclose(back, np.r_[L[2, 1], L[0, 2], L[1, 0], L[:3, 3]], atol=1e-9, rtol=0)


def make_case(n=5000, seed=3):
    rng = np.random.default_rng(seed)
    cloud = (rng.normal(size=(n, 3)) * [30, 30, 3]).astype(np.float32)
    stamps = rng.uniform(1.7e9, 1.7e9 + 0.1, n)  # absolute seconds, as a FLOAT64 `timestamp` field would carry them
    Rs, ts = synth.rot_from_rotvec(np.array([0.01, -0.02, 0.3])), np.array([10.0, -3.0, 0.5])
    Rf, tf = Rs @ synth.rot_from_rotvec(np.array([0.002, 0.001, 0.03])), ts + Rs @ np.array([0.8, 0.05, -0.01])
    return cloud, stamps, (Rs, ts), (Rf, tf)


def test_deskew_oracle_properties(pre_oracle):
    cloud, stamps, start, finish = make_case()
    out, moved = pre_oracle.deskew(cloud, stamps, start, finish)
    assert moved and out.shape == cloud.shape
    # the point stamped in the middle of the scan does not move; the ends move by +- half the inter-pose motion
    xi = pre_oracle.pose3_logmap(start[0].T @ finish[0], start[0].T @ (finish[1] - start[1]))
    s = (stamps - stamps.min()) / (stamps.max() - stamps.min())
    for i in (int(np.argmin(np.abs(s - 0.5))), int(np.argmin(s)), int(np.argmax(s))):
        T = expm(twist_matrix((s[i] - 0.5) * xi))
        want = T[:3, :3] @ cloud[i].astype(np.float64) + T[:3, 3]
        np.testing.assert_allclose(out[i], want.astype(np.float32), atol=2e-6 * np.abs(want).max(), rtol=0)
    # equal stamps -> unchanged cloud (:415), zero motion -> unchanged cloud
    same, moved0 = pre_oracle.deskew(cloud, np.full(len(cloud), 3.0), start, finish)
    assert not moved0
    np.testing.assert_array_equal(same, cloud)
    still, _ = pre_oracle.deskew(cloud, stamps, start, start)
    np.testing.assert_array_equal(still, cloud)
    # KITTI branch: stamps come from the azimuth of the tilted point
    k, moved_k = pre_oracle.deskew(cloud, None, start, finish, kitti=True)
    assert moved_k and np.isfinite(k).all() and np.abs(k - cloud).max() < 3.0


def test_pose_compose_matches_the_homogeneous_product():
    """svnicp_pose_compose is host arithmetic inside the C-ABI library (no GPU needed)."""
    rng = np.random.default_rng(0)
    for _ in range(10):
        R0, t0 = synth.rot_from_rotvec(rng.normal(size=3)), rng.normal(size=3) * 20
        mean = rng.normal(size=6) * [0.3, 0.3, 0.1, 0.01, 0.01, 0.02]
        R, t = sv.pose_compose(R0, t0, mean)
        Tc = np.eye(4)
        Tc[:3, :3] = expm(twist_matrix(np.r_[mean[3:], 0, 0, 0]))[:3, :3]  # Rot3::Expmap(mean[3:6])
        Tc[:3, 3] = mean[:3]
        T0 = np.eye(4)
        T0[:3, :3], T0[:3, 3] = R0, t0
        T = T0 @ Tc  # OdometryPipeline.cpp:44: initial_guess.matrix() * correction_pose.matrix()
        np.testing.assert_allclose(R, T[:3, :3], atol=1e-13, rtol=0)
        np.testing.assert_allclose(t, T[:3, 3], atol=1e-12, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("kitti", [False, True])
def test_deskew_gpu_matches_the_restatement(pre_oracle, kitti):
    cloud, stamps, start, finish = make_case(n=143000, seed=9)
    want, moved_want = pre_oracle.deskew(cloud, None if kitti else stamps, start, finish, kitti=kitti)
    pre = sv.ScanPreprocessor(len(cloud))
    ptr, n, moved = pre.deskew_pointcloud(cloud, None if kitti else stamps, start, finish, kitti=kitti)
    got = pre.download(ptr, n)
    assert moved == moved_want and n == len(cloud)
    # fp64 on both sides, rounded to float once: libm vs CUDA sin/cos may differ in the last float ulp of the result
    ulp = np.spacing(np.abs(want).astype(np.float32))
    assert np.all(np.abs(got - want) <= ulp), float(np.abs(got - want).max())
    assert np.mean(got != want) < 1e-3
    # chains without leaving the device: deskew -> crop
    ptr2, n2 = pre.crop_pointcloud(ptr, 1.0, 100.0, n=n, on_device=True)
    assert 0 < n2 <= n
    # equal stamps: unchanged cloud, moved = False
    ptr3, n3, moved3 = pre.deskew_pointcloud(cloud, np.zeros(len(cloud)), start, finish)
    assert not moved3
    np.testing.assert_array_equal(pre.download(ptr3, n3), cloud)
