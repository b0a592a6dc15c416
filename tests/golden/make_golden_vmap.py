"""Golden vectors for the local map (svnicp::VoxelHashMap), produced by the REFERENCE's own VoxelHashMap.cpp compiled over
stand-in PCL / Eigen / tsl types (oracle/_ref/libvmap_ref.so, see oracle/build_ref.sh and oracle/ref_shim_map).

Run in the build container only:  python tests/golden/make_golden_vmap.py
The fixture holds a short synthetic drive (sensor-frame float32 clouds + poses) and, after every AddPointCloud, the map size and
the lexicographically sorted GetMap() / GetMap(pose, range) outputs (point order is hash-map iteration order: not part of the
contract).
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def sort_rows(a):
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))] if len(a) else a


def main():
    sys.path.insert(0, ROOT)
    import oracle as orc
    from svn_icp_b200 import synth

    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
    world = synth.make_world(0xC0FFEE)
    rng = np.random.default_rng(7)
    voxel, max_range, cap, get_range = 1.0, 35.0, 6, 25.0
    ref = orc.ReferenceMap(voxel, max_range, cap)
    out = dict(params=np.array([voxel, max_range, cap, get_range]))
    n_scans = 8
    for k in range(n_scans):
        pts, (R, t) = synth.make_scan(world, 4 * k, "16", 0xC0FFEE)  # 4 scans apart: 3.2 m per step, voxels age out
        pts = pts[rng.permutation(len(pts))[:4000]].astype(np.float32)
        ref.AddPointCloud(pts, R, t)
        out[f"cloud{k}"], out[f"R{k}"], out[f"t{k}"] = pts, R, t
        out[f"size{k}"] = np.array([ref.Size()])
        out[f"all{k}"] = sort_rows(ref.GetMap()).astype(np.float32)
        out[f"near{k}"] = sort_rows(ref.GetMap(t, get_range)).astype(np.float32)
        print(k, "voxels", ref.Size(), "points", len(out[f"all{k}"]), "near", len(out[f"near{k}"]))
    out["n_scans"] = np.array([n_scans])
    np.savez_compressed(os.path.join(HERE, "vmap_sequence.npz"), **out)


if __name__ == "__main__":
    main()
