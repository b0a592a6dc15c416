"""N > 1 host logic on CPU: world_size-2 (and 3) `gloo` processes emulate the particle-sharded scan with the oracle's
per-slice Gauss-Newton systems, ONE all-gather of the packed record per iteration, a sliced Stein step and the deferred
early-stop decision -- the same sequence csrc/capi.cu drives on the GPUs -- and must reproduce the single-process
oracle bit for bit."""
import os
import socket

import numpy as np
import pytest

from svn_icp_b200 import sharding, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P, full, early_stop, thr, out_dir):
    import torch
    import torch.distributed as dist
    import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = orc.Oracle()
    O.set_num_threads(1)
    pb = synth.make_uniform_problem(P, 250, 2500, seed=21, box=8.0)
    K, I, md, lr = 12, 9, 3.0, 1.0
    lo, hi, L = sharding.slice_of(P, rank, world)
    q0 = O.transform_q0(pb.source, pb.R0, pb.t0)
    cand, _ = O.knn_mink(q0, pb.target, K)
    R = np.stack([O.so3_exp(pb.init_pose[3:, p])[0] for p in range(P)])
    t = np.ascontiguousarray(pb.init_pose[:3].T)
    dnorm = np.zeros(P)
    history = np.zeros((I, 6, P), dtype=np.float32)
    stop, iters_done, it = False, 0, 0

    def gather(local_rows):
        buf = np.zeros((L, sharding.REC))
        buf[: hi - lo] = local_rows
        out = [torch.zeros(L, sharding.REC, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, torch.from_numpy(buf))  # the single collective of the iteration
        return torch.cat(out).numpy()[:P]

    def local_record(with_gn):
        rows = np.zeros((hi - lo, sharding.REC))
        rows[:, 0:3] = t[lo:hi]
        rows[:, 3:6] = np.stack([O.so3_log(R[p]) for p in range(lo, hi)])
        rows[:, sharding.REC_DNORM] = dnorm[lo:hi]
        if with_gn:
            H, b = O.gn(R[lo:hi], t[lo:hi], pb.R0, pb.t0, pb.source, pb.target, cand, md)
            rows[:, sharding.REC_B:sharding.REC_B + 6] = b
            for r in range(6):
                for c in range(r, 6):
                    rows[:, sharding.REC_H + sharding.tri_index(r, c)] = H[:, r, c]
        return rows

    def decide(rec, epilogue):
        nonlocal stop, iters_done
        if stop:
            return
        mean_norm = rec[:, sharding.REC_DNORM].sum() / P
        if early_stop and it > 0 and mean_norm < thr:
            stop, iters_done = True, it
            return
        if it > 0:
            history[it - 1] = rec[:, 0:6].T.astype(np.float32)
        if epilogue:
            iters_done = it

    for _ in range(I):
        rec = gather(local_record(True))
        decide(rec, False)
        if stop:
            continue  # later iterations are no-ops on every rank (the collective still runs)
        x = rec[:, 0:6]
        H = np.zeros((P, 6, 6))
        for r in range(6):
            for c in range(6):
                H[:, r, c] = rec[:, sharding.REC_H + sharding.tri_index(r, c)]
        b = rec[:, sharding.REC_B:sharding.REC_B + 6]
        delta, _ = O.stein_step(x, H, b, full=full, lr=lr)  # every rank needs only its slice of delta
        R[lo:hi], t[lo:hi] = O.pose_update(R[lo:hi], t[lo:hi], delta[lo:hi])
        dnorm[lo:hi] = np.linalg.norm(delta[lo:hi], axis=1)
        it += 1
    rec = gather(local_record(False))
    decide(rec, True)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), particles=rec[:, 0:6].T, history=history, iters=iters_done)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,P,full,early_stop", [(2, 9, True, False), (2, 8, False, True), (3, 10, True, True)])
def test_sharded_scan_equals_single_process(tmp_path, world, P, full, early_stop):
    import torch.multiprocessing as mp
    import oracle as orc

    thr = 2.5e-2 if P == 8 else 4.5e-2  # fires at epoch 4 (P=8) / epoch 5 (P=10): mid-run on purpose
    mp.spawn(_worker, args=(world, _free_port(), P, full, early_stop, thr, str(tmp_path)), nprocs=world, join=True)
    O = orc.Oracle()
    pb = synth.make_uniform_problem(P, 250, 2500, seed=21, box=8.0)
    prm = orc.make_params(iterations=9, knn_count=12, max_dist=3.0, lr=1.0, svn_full_grad=full, check_early_stop=early_stop,
                          convergence_threshold=thr)
    o = O.align(prm, pb.source, pb.target, pb.init_pose, pb.R0, pb.t0)
    res = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for r in res:
        np.testing.assert_array_equal(r["particles"], res[0]["particles"])  # every rank ends with the full result
        assert int(r["iters"]) == o["iters_done"]
        # same arithmetic; the only difference is that the record carries H as its upper triangle (exactly symmetric)
        # while the oracle's J^T(wJ) is symmetric only up to rounding
        np.testing.assert_allclose(r["particles"], o["particles"], rtol=0, atol=1e-11)
        np.testing.assert_allclose(r["history"], o["history"], rtol=0, atol=2e-7)
    if early_stop:
        assert o["iters_done"] < 9, "pick a threshold that makes the stop fire"


def test_slice_rules():
    assert sharding.slice_of(1000, 0, 8) == (0, 125, 125)
    assert sharding.slice_of(1000, 7, 8) == (875, 1000, 125)
    assert sharding.slice_of(10, 2, 3) == (8, 10, 4)
    covered = []
    for r in range(4):
        lo, hi, L = sharding.slice_of(4096 + 3, r, 4)
        covered += list(range(lo, hi))
    assert covered == list(range(4099))
    with pytest.raises(ValueError):
        sharding.slice_of(3, 0, 4)
    with pytest.raises(ValueError):
        sharding.slice_of(9, 3, 4)  # ceil(9/4)=3 -> rank 3 would be empty
    assert sharding.tri_index(0, 0) == 0 and sharding.tri_index(5, 5) == 20 and sharding.tri_index(3, 1) == sharding.tri_index(1, 3)
    assert sharding.allgather_bytes_per_iteration(4096, 8) == 4096 * 40 * 8


def _svgd_worker(rank, world, port, P, optimizer, lr, early_stop, thr, out_dir):
    """The SVGD-ICP class sharded the same way (csrc/svgd_class.cu + capi.cu): per-slice first-order gradient, ONE
    all-gather of (kernel position, gradient, |pose difference|), redundant RBF step, per-slice optimizer update."""
    import torch
    import torch.distributed as dist
    import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = orc.Oracle()
    O.set_num_threads(1)
    pb = synth.make_uniform_problem(P, 250, 2500, seed=33, box=8.0)
    K, I, md = 12, 8, 3.0
    lo, hi, L = sharding.slice_of(P, rank, world)
    q0 = O.transform_q0(pb.source, pb.R0, pb.t0)
    cand, _ = O.knn_mink(q0, pb.target, K)
    x = np.ascontiguousarray(pb.init_pose.T)          # parameters [P][6]; only [lo:hi) is live on this rank
    prev = np.ascontiguousarray(pb.init_pose.T) + 0.01  # pose_particles_ as the constructor left it (stale on purpose)
    state = np.zeros((hi - lo, 6, 2))
    dnorm = np.zeros(P)
    history = np.zeros((I, 6, P), dtype=np.float32)
    stop, iters_done, it = False, 0, 0

    def gather(rows):
        buf = np.zeros((L, sharding.REC))
        buf[: hi - lo] = rows
        out = [torch.zeros(L, sharding.REC, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, torch.from_numpy(buf))
        return torch.cat(out).numpy()[:P]

    def local_record(with_grad):
        rows = np.zeros((hi - lo, sharding.REC))
        rows[:, 0:6] = prev[lo:hi] if (with_grad and it == 0) else x[lo:hi]   # kernel position (k_finalize_first)
        rows[:, sharding.REC_DNORM] = dnorm[lo:hi]
        if with_grad:
            rows[:, sharding.REC_B:sharding.REC_B + 6] = O.svgd_grad(x[lo:hi], pb.R0, pb.t0, pb.source, pb.target, cand, md)
        return rows

    def decide(rec, epilogue):
        nonlocal stop, iters_done
        if stop:
            return
        if early_stop and it > 0 and rec[:, sharding.REC_DNORM].sum() / P < thr:
            stop, iters_done = True, it
            return
        if it > 0:
            history[it - 1] = rec[:, 0:6].T.astype(np.float32)
        if epilogue:
            iters_done = it

    for _ in range(I):
        rec = gather(local_record(True))
        decide(rec, False)
        if stop:
            continue
        kpos, g = np.ascontiguousarray(rec[:, 0:6]), np.ascontiguousarray(rec[:, sharding.REC_B:sharding.REC_B + 6])
        stein, _ = O.svgd_step(kpos, -g)
        xl = np.ascontiguousarray(x[lo:hi]).reshape(-1)
        O.opt_step(optimizer, lr, it + 1, xl, -stein[lo:hi].reshape(-1), state.reshape(-1, 2))
        x[lo:hi] = xl.reshape(-1, 6)
        dnorm[lo:hi] = np.linalg.norm(x[lo:hi] - kpos[lo:hi], axis=1)
        it += 1
    rec = gather(local_record(False))
    decide(rec, True)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), particles=rec[:, 0:6].T, history=history, iters=iters_done)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,P,optimizer,lr,early_stop", [(2, 9, "Adam", 0.03, False), (2, 8, "RMSprop", 0.01, True), (3, 10, "Adagrad", 0.03, False)])
def test_sharded_svgd_class_equals_single_process(tmp_path, world, P, optimizer, lr, early_stop):
    import torch.multiprocessing as mp
    import oracle as orc

    thr = 8.0e-2  # RMSprop case: mean |pose difference| = 0.23, 0.084, 0.18, 0.18, 0.071 -> the stop fires at epoch 4
    mp.spawn(_svgd_worker, args=(world, _free_port(), P, optimizer, lr, early_stop, thr, str(tmp_path)), nprocs=world, join=True)
    O = orc.Oracle()
    pb = synth.make_uniform_problem(P, 250, 2500, seed=33, box=8.0)
    prm = orc.make_svgd_params(iterations=8, lr=lr, max_dist=3.0, knn_count=12, optimizer=optimizer, check_early_stop=early_stop,
                               convergence_threshold=thr)
    o = O.svgd_align(prm, pb.source, pb.target, pb.init_pose + 0.01, pb.init_pose, pb.R0, pb.t0)
    res = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for r in res:
        np.testing.assert_array_equal(r["particles"], res[0]["particles"])
        assert int(r["iters"]) == o["iters_done"]
        np.testing.assert_allclose(r["particles"], o["particles"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(r["history"], o["history"], rtol=0, atol=2e-7)
    if early_stop:
        assert 1 < o["iters_done"] < 8, "pick a threshold that makes the stop fire mid-run"
