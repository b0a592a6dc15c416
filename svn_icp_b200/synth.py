"""Seeded synthetic LiDAR odometry data.

Not part of the registration path: bench.py and tests/ use it to make the scans both the CUDA
path and the CPU oracle consume.  Everything is float64 numpy.

World: ground plane z=0, two walls y=+-15 m (8 m high), 200 axis-aligned boxes and 100 vertical
cylinders in a 400 x 60 m corridor.  Sensors: "64" = 64 beams (-24.8..+2.0 deg) x 2250 azimuths
(geometry constants of the reference's include/segmentation/ImageProjection.h:63-67), "128" =
128 beams (-25..+13.1 deg) x 2048 azimuths (:70-75).  Range noise N(0, 0.02 m), 100 m range cut
(config/geodeAlpha.yaml:21).  Local map: 1 m voxel hash, <= 20 points per voxel, first come first
kept, voxel index = trunc(p / voxel) (reference VoxelHashMap.cpp:22-41), voxels kept when their
first point is within map_range of the pose (:51-61).
"""
from __future__ import annotations

import dataclasses
import numpy as np

SENSORS = {
    "64": dict(beams=64, el0=-24.8, el1=2.0, cols=2250),
    "128": dict(beams=128, el0=-25.0, el1=13.1, cols=2048),
    # small sensors for CPU-sized tests
    "16": dict(beams=16, el0=-24.8, el1=2.0, cols=360),
    "32": dict(beams=32, el0=-24.8, el1=2.0, cols=900),
}
SENSOR_HEIGHT = 1.8
MAX_RANGE = 100.0
MIN_RANGE = 1.0


def rot_from_rotvec(r):
    r = np.asarray(r, dtype=np.float64)
    a = np.linalg.norm(r)
    if a < 1e-12:
        return np.eye(3)
    n = r / a
    K = np.array([[0, -n[2], n[1]], [n[2], 0, -n[0]], [-n[1], n[0], 0]])
    return np.cos(a) * np.eye(3) + (1 - np.cos(a)) * np.outer(n, n) + np.sin(a) * K


def rotvec_from_rot(R):
    v = np.clip(0.5 * (np.trace(R) - 1.0), -1.0, 1.0)
    a = np.arccos(v)
    s = np.sin(a)
    if abs(s) <= 1e-12:
        return np.zeros(3)
    return 0.5 / s * a * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])


@dataclasses.dataclass
class World:
    boxes_lo: np.ndarray  # [nb,3]
    boxes_hi: np.ndarray  # [nb,3]
    cyl_c: np.ndarray  # [nc,2]
    cyl_r: np.ndarray  # [nc]
    cyl_h: np.ndarray  # [nc]
    wall_y: float = 15.0
    wall_h: float = 8.0


def make_world(seed: int = 0xC0FFEE, n_boxes: int = 200, n_cyl: int = 100, length: float = 400.0) -> World:
    rng = np.random.default_rng(seed)
    c = np.stack([rng.uniform(-length / 2, length / 2, n_boxes), rng.uniform(-14.0, 14.0, n_boxes)], 1)
    # keep a 3 m wide driving lane around y=0 free
    c[:, 1] = np.where(np.abs(c[:, 1]) < 3.0, np.sign(c[:, 1] + 1e-9) * (3.0 + np.abs(c[:, 1])), c[:, 1])
    size = rng.uniform(0.5, 4.0, (n_boxes, 3))
    lo = np.concatenate([c - size[:, :2] / 2, np.zeros((n_boxes, 1))], 1)
    hi = np.concatenate([c + size[:, :2] / 2, size[:, 2:3]], 1)
    cc = np.stack([rng.uniform(-length / 2, length / 2, n_cyl), rng.uniform(-14.0, 14.0, n_cyl)], 1)
    cc[:, 1] = np.where(np.abs(cc[:, 1]) < 3.0, np.sign(cc[:, 1] + 1e-9) * (3.0 + np.abs(cc[:, 1])), cc[:, 1])
    return World(lo, hi, cc, rng.uniform(0.2, 0.6, n_cyl), rng.uniform(2.0, 8.0, n_cyl))


def ray_dirs(sensor: str) -> np.ndarray:
    s = SENSORS[sensor]
    el = np.deg2rad(np.linspace(s["el0"], s["el1"], s["beams"]))
    az = np.linspace(-np.pi, np.pi, s["cols"], endpoint=False)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (len(el), len(az)))], -1)
    return d.reshape(-1, 3)


def cast(world: World, origin: np.ndarray, dirs: np.ndarray) -> np.ndarray:
    """Range of the first hit along each ray (inf if none)."""
    o = np.asarray(origin, dtype=np.float64)
    n = dirs.shape[0]
    best = np.full(n, np.inf)
    dx, dy, dz = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        # ground
        t = -o[2] / dz
        best = np.where((t > 0) & (t < best), t, best)
        # walls
        for wy in (-world.wall_y, world.wall_y):
            t = (wy - o[1]) / dy
            z = o[2] + t * dz
            ok = (t > 0) & (z >= 0) & (z <= world.wall_h) & (t < best)
            best = np.where(ok, t, best)
        inv = 1.0 / dirs
        # boxes: slab test, only those that can be within range
        near = np.linalg.norm(0.5 * (world.boxes_lo[:, :2] + world.boxes_hi[:, :2]) - o[:2], axis=1) < MAX_RANGE + 6
        for lo, hi in zip(world.boxes_lo[near], world.boxes_hi[near]):
            t0 = (lo - o) * inv
            t1 = (hi - o) * inv
            tmin = np.max(np.minimum(t0, t1), axis=1)
            tmax = np.min(np.maximum(t0, t1), axis=1)
            ok = (tmax >= tmin) & (tmin > 0) & (tmin < best)
            best = np.where(ok, tmin, best)
        # cylinders
        a = dx * dx + dy * dy
        nearc = np.linalg.norm(world.cyl_c - o[:2], axis=1) < MAX_RANGE + 2
        for c, r, h in zip(world.cyl_c[nearc], world.cyl_r[nearc], world.cyl_h[nearc]):
            ox, oy = o[0] - c[0], o[1] - c[1]
            b = ox * dx + oy * dy
            cc = ox * ox + oy * oy - r * r
            disc = b * b - a * cc
            t = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / a
            z = o[2] + t * dz
            ok = (disc > 0) & (t > 0) & (z >= 0) & (z <= h) & (t < best)
            best = np.where(ok, t, best)
    return best


def pose_at(k: int, dt: float = 0.1, v: float = 8.0, yaw_rate: float = 0.05, x0: float = -150.0):
    """Ground-truth sensor pose (R, t) of scan k: 8 m/s forward, 0.05 rad/s yaw, 10 Hz."""
    yaw = 0.0
    p = np.array([x0, 0.0, SENSOR_HEIGHT])
    for _ in range(k):
        p = p + v * dt * np.array([np.cos(yaw), np.sin(yaw), 0.0])
        yaw += yaw_rate * dt
    return rot_from_rotvec([0, 0, yaw]), p


def make_scan(world: World, k: int, sensor: str = "64", seed: int = 0xC0FFEE, noise: float = 0.02):
    """Points of scan k in the SENSOR frame, plus the ground-truth pose."""
    R, t = pose_at(k)
    d_s = ray_dirs(sensor)
    d_w = d_s @ R.T
    rng = np.random.default_rng(seed + k)
    r = cast(world, t, d_w)
    r = r + rng.normal(0.0, noise, r.shape)
    ok = np.isfinite(r) & (r < MAX_RANGE) & (r > MIN_RANGE)
    return d_s[ok] * r[ok, None], (R, t)


def voxel_cap(points: np.ndarray, voxel: float, cap: int) -> np.ndarray:
    """Keep the first `cap` points of each voxel (first come first kept)."""
    if len(points) == 0:
        return points
    key = np.trunc(points / voxel).astype(np.int64)
    key = (key[:, 0] + (1 << 20)) << 42 | (key[:, 1] + (1 << 20)) << 21 | (key[:, 2] + (1 << 20))
    order = np.argsort(key, kind="stable")
    ks = key[order]
    start = np.r_[0, np.flatnonzero(ks[1:] != ks[:-1]) + 1]
    rank = np.arange(len(ks)) - np.repeat(start, np.diff(np.r_[start, len(ks)]))
    keep = np.sort(order[rank < cap])
    return points[keep]


def uniform_downsample(points: np.ndarray, leaf: float) -> np.ndarray:
    """One point per leaf-sized voxel (stand-in for pcl::UniformSampling, OdometryPipeline.cpp:684-690)."""
    return voxel_cap(points, leaf, 1)


@dataclasses.dataclass
class ScanProblem:
    source: np.ndarray  # [N_s,3] sensor frame
    target: np.ndarray  # [N_t,3] world frame (local map)
    R0: np.ndarray  # [3,3] initial guess rotation
    t0: np.ndarray  # [3]
    init_pose: np.ndarray  # [6,P] fresh particles (component major)
    gt_rel: np.ndarray  # [6] ground-truth relative correction [t; Log R] with T_gt = T0 * T_rel
    R_gt: np.ndarray
    t_gt: np.ndarray


PARTICLE_UB = np.array([0.3, 0.2, 0.1, 0.004, 0.004, 0.012])  # OdometryPipeline.cpp:661-667


def init_particles(P: int, rng: np.random.Generator, ub=PARTICLE_UB) -> np.ndarray:
    """initialize_particles (ICPUtils.cpp:45-58): uniform in [-ub, ub]; P == 1 -> zeros."""
    if P == 1:
        return np.zeros((6, 1))
    return (2 * ub[:, None]) * rng.random((6, P)) - ub[:, None]


def make_problem(P: int, sensor: str = "64", scan_index: int = 8, n_map_scans: int = 8, seed: int = 0xC0FFEE,
                 downsample: bool = False, voxel: float = 1.0, map_cap: int = 20, map_range: float = 100.0,
                 max_source: int | None = None, world: World | None = None) -> ScanProblem:
    world = world or make_world(seed)
    rng = np.random.default_rng(seed ^ 0x5EED ^ scan_index)
    pts = []
    for k in range(scan_index - n_map_scans, scan_index):
        s, (R, t) = make_scan(world, k, sensor, seed)
        pts.append(s @ R.T + t)
    allp = voxel_cap(np.concatenate(pts), voxel, map_cap)
    src, (Rg, tg) = make_scan(world, scan_index, sensor, seed)
    # GetMap(pose, range): keep voxels near the pose; emit in an arbitrary (hash) order
    allp = allp[np.linalg.norm(allp - tg, axis=1) < map_range]
    target = allp[rng.permutation(len(allp))]
    if downsample:
        src = uniform_downsample(uniform_downsample(src, 0.5 * voxel), 1.5 * voxel)
    if max_source is not None and len(src) > max_source:
        src = src[np.sort(rng.choice(len(src), max_source, replace=False))]
    # initial guess = ground truth perturbed by N(0, diag(0.05,0.05,0.02 m, 0.002 rad x3))
    pert = rng.normal(0.0, [0.05, 0.05, 0.02, 0.002, 0.002, 0.002])
    R0 = Rg @ rot_from_rotvec(pert[3:])
    t0 = tg + pert[:3]
    Rrel = R0.T @ Rg
    trel = R0.T @ (tg - t0)
    return ScanProblem(np.ascontiguousarray(src), np.ascontiguousarray(target), R0, t0, init_particles(P, rng),
                       np.r_[trel, rotvec_from_rot(Rrel)], Rg, tg)


def saturated_map(world: World, centre: np.ndarray, map_range: float, rng: np.random.Generator, voxel: float = 1.0,
                  cap: int = 20, density: float = 40.0, noise: float = 0.02) -> np.ndarray:
    """Local map of a region that has been driven through: every surface voxel within map_range holds `cap` points
    (what the reference's VoxelHashMap converges to, VoxelHashMap.cpp:22-41).  Surfaces are sampled directly
    (ground, walls, box faces, cylinder mantles) at `density` points / m^2, then capped per voxel."""
    cx, cy = centre[0], centre[1]
    x0, x1 = cx - map_range, cx + map_range
    parts = []

    def plane(n, ax_u, lo_u, hi_u, ax_v, lo_v, hi_v, ax_w, w):
        p = np.empty((n, 3))
        p[:, ax_u] = rng.uniform(lo_u, hi_u, n)
        p[:, ax_v] = rng.uniform(lo_v, hi_v, n)
        p[:, ax_w] = w
        return p

    wy = world.wall_y
    parts.append(plane(int(density * (x1 - x0) * 2 * wy), 0, x0, x1, 1, -wy, wy, 2, 0.0))
    for s in (-wy, wy):
        parts.append(plane(int(density * (x1 - x0) * world.wall_h), 0, x0, x1, 2, 0.0, world.wall_h, 1, s))
    for lo, hi in zip(world.boxes_lo, world.boxes_hi):
        if hi[0] < x0 or lo[0] > x1:
            continue
        d = hi - lo
        parts.append(plane(int(density * d[0] * d[1]) + 1, 0, lo[0], hi[0], 1, lo[1], hi[1], 2, hi[2]))
        for s in (lo[0], hi[0]):
            parts.append(plane(int(density * d[1] * d[2]) + 1, 1, lo[1], hi[1], 2, lo[2], hi[2], 0, s))
        for s in (lo[1], hi[1]):
            parts.append(plane(int(density * d[0] * d[2]) + 1, 0, lo[0], hi[0], 2, lo[2], hi[2], 1, s))
    for c, r, h in zip(world.cyl_c, world.cyl_r, world.cyl_h):
        if c[0] + r < x0 or c[0] - r > x1:
            continue
        n = int(density * 2 * np.pi * r * h) + 1
        a = rng.uniform(0, 2 * np.pi, n)
        parts.append(np.stack([c[0] + r * np.cos(a), c[1] + r * np.sin(a), rng.uniform(0, h, n)], 1))
    pts = np.concatenate(parts)
    pts = pts + rng.normal(0.0, noise, pts.shape)
    pts = pts[np.linalg.norm(pts - np.asarray(centre), axis=1) < map_range]
    pts = pts[rng.permutation(len(pts))]
    return voxel_cap(pts, voxel, cap)


def make_problem_saturated(P: int, sensor: str = "64", scan_index: int = 8, seed: int = 0xC0FFEE, map_range: float = 100.0,
                           voxel: float = 1.0, map_cap: int = 20, world: World | None = None) -> ScanProblem:
    """BASELINE.json configs[1]/[2] workload: raw scan (~120k / ~260k points) against a saturated voxel-hash local map."""
    world = world or make_world(seed)
    rng = np.random.default_rng(seed ^ 0xBEEF ^ scan_index)
    src, (Rg, tg) = make_scan(world, scan_index, sensor, seed)
    target = saturated_map(world, tg, map_range, rng, voxel, map_cap)
    target = target[rng.permutation(len(target))]
    pert = rng.normal(0.0, [0.05, 0.05, 0.02, 0.002, 0.002, 0.002])
    R0 = Rg @ rot_from_rotvec(pert[3:])
    t0 = tg + pert[:3]
    Rrel = R0.T @ Rg
    trel = R0.T @ (tg - t0)
    return ScanProblem(np.ascontiguousarray(src), np.ascontiguousarray(target), R0, t0, init_particles(P, rng),
                       np.r_[trel, rotvec_from_rot(Rrel)], Rg, tg)


def make_uniform_problem(P: int, n_s: int, n_t: int, seed: int = 1, box: float = 20.0,
                         motion=(0.15, -0.08, 0.03, 0.0, 0.0, 0.01)) -> ScanProblem:
    """Small planted-motion problem (the survey's probe): target = random surface
    samples, source = subset of the target moved by the inverse of `motion`."""
    rng = np.random.default_rng(seed)
    # points on a few planes so that ICP is well conditioned
    n_each = n_t // 3
    a = np.stack([rng.uniform(-box, box, n_each), rng.uniform(-box, box, n_each), np.zeros(n_each)], 1)
    b = np.stack([rng.uniform(-box, box, n_each), np.full(n_each, box), rng.uniform(0, 6, n_each)], 1)
    c = np.stack([np.full(n_t - 2 * n_each, -box), rng.uniform(-box, box, n_t - 2 * n_each), rng.uniform(0, 6, n_t - 2 * n_each)], 1)
    target = np.concatenate([a, b, c])
    target = target[rng.permutation(len(target))]
    target = target + rng.normal(0, 0.01, target.shape)
    Rm = rot_from_rotvec(motion[3:])
    tm = np.asarray(motion[:3], dtype=np.float64)
    pick = rng.choice(n_t, n_s, replace=False)
    # target = Rm src + tm  ->  src = Rm^T (target - tm)
    src = (target[pick] - tm) @ Rm + rng.normal(0, 0.005, (n_s, 3))
    return ScanProblem(np.ascontiguousarray(src), np.ascontiguousarray(target), np.eye(3), np.zeros(3),
                       init_particles(P, rng), np.asarray(motion, dtype=np.float64), Rm, tm)
