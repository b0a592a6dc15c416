#!/usr/bin/env python
"""bench.py -- scans/sec of the SVN-ICP registration inner loop on N B200s (one process per GPU).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one scan: add_cloud -> set_initial_mean -> stein_align -> getters (the call sequence of the
reference's caller, OdometryPipeline.cpp:582-607) on BASELINE.json configs[1]: synthetic 64-beam LiDAR scan
(~120k points) against its voxel-hash local map, 1000 particles, 30 iterations, early stop off, K = 100.
  value       scans/sec with the clouds already resident in HBM (device pointers), CUDA-event timed, max over ranks
  e2e         the same scan through the public API with HOST (pinned) clouds: H2D copies and result D2H inside the timing
  roofline    correspondence+reduction pass (k_filter + k_gn): algorithmic bytes / measured device time vs measured HBM peak
  cpu_baseline  the CPU checker timed on this box's host cores on a bounded sample (reported baseline, not the target)
With N > 1 the particles are sharded across the ranks (strong scaling of the same scan; one ncclAllGather per iteration).
--impl reference times the reference's own CPU implementation (oracle/_ref, libtorch on all host cores; the C port if
_ref is absent) on a bounded sample of the same workload and prints the same JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(sensor="64", particles=1000, iterations=30, K=100, max_dist=3.0, lr=1.0, svn_full_grad=True,
                scan_index=8, n_map_scans=8)
METRIC = "scans/sec at 1000 particles (64-beam ~120k-pt synthetic scan, K=100, 30 SVN iterations)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


def make_problem(P, seed=0xC0FFEE):
    from svn_icp_b200 import synth
    t = time.time()
    pb = synth.make_problem_saturated(P, sensor=WORKLOAD["sensor"], scan_index=WORKLOAD["scan_index"], seed=seed)
    return pb, time.time() - t


def _sampled_scan_time(run, ns_full, I, ns1, ns2):
    """Affine model of the CPU scan time from four bounded runs: T(n_s, iters) = setup(n_s) + iters * (a + b n_s).
    setup (the brute-force K-NN) and b (correspondence + Gauss-Newton) scale with N_s, a (the P x P Stein step) does not.
    run(ns, iters) -> seconds.  Returns (t_scan_full, description)."""
    T22, T21, T12, T11 = run(ns2, 2), run(ns2, 1), run(ns1, 2), run(ns1, 1)
    it2, it1 = max(T22 - T21, 1e-9), max(T12 - T11, 1e-9)
    b = max((it2 - it1) / (ns2 - ns1), 0.0)
    a = max(it2 - b * ns2, 0.0)
    setup2 = max(T21 - it2, 0.0)
    t_scan = setup2 * ns_full / ns2 + I * (a + b * ns_full)
    desc = (f"runs at N_s={ns1},{ns2} x iterations=1,2: setup {setup2:.2f}s@{ns2} pts, per iteration {a:.3f}s (Stein, N_s-independent) + "
            f"{b * 1e3:.3f} ms/point; extrapolated to N_s={ns_full}, {I} iterations")
    return t_scan, desc


def _subsample(pb, ns, seed=1):
    sel = np.sort(np.random.default_rng(seed).choice(len(pb.source), ns, replace=False))
    return pb.source[sel]


def cpu_baseline_port(pb):
    """The C restatement (OpenMP, all host cores) on a bounded sample of the same workload (all particles, the full map,
    subsets of the source points, 1-2 iterations), extrapolated with the affine model above."""
    import oracle as orc
    O = orc.Oracle()
    K, I = WORKLOAD["K"], WORKLOAD["iterations"]

    def run_once(ns, iters):
        prm = orc.make_params(iterations=iters, knn_count=K, max_dist=WORKLOAD["max_dist"], lr=WORKLOAD["lr"], svn_full_grad=WORKLOAD["svn_full_grad"])
        src = _subsample(pb, ns)
        t0 = time.time()
        O.align(prm, src, pb.target, pb.init_pose, pb.R0, pb.t0)
        return time.time() - t0

    def run(ns, iters):  # best of two: the four timings are differenced, so one-off stalls (thread-pool start, page faults) matter
        return min(run_once(ns, iters), run_once(ns, iters))

    run_once(256, 1)  # warm-up: OpenMP thread pool, page-in of the map
    t_scan, desc = _sampled_scan_time(run, len(pb.source), I, 1024, 2048)
    return dict(value=1.0 / t_scan, unit="scans/sec", cores=O.num_threads(), kind="port",
                sample=f"C port (OpenMP): all {pb.init_pose.shape[1]} particles, full {len(pb.target)}-point map; " + desc)


def run_reference(args, rank, world):
    """--impl reference: the reference's own sources on the host cores (oracle/_ref; the C port if absent), bounded sample per step."""
    if rank != 0:
        return
    import oracle as orc
    P = WORKLOAD["particles"]
    pb, _ = make_problem(P)
    K, I = WORKLOAD["K"], WORKLOAD["iterations"]
    ns_full = len(pb.source)
    if orc.ref_available():
        eng = orc.Reference()
        kind, cores = "reference", eng.num_threads()
        call = lambda prm, src: eng.scan(prm, src, pb.target, pb.init_pose, pb.R0, pb.t0)
    else:
        eng = orc.Oracle()
        kind, cores = "port", eng.num_threads()
        call = lambda prm, src: eng.align(prm, src, pb.target, pb.init_pose, pb.R0, pb.t0)

    def run(ns, iters):
        prm = orc.make_params(iterations=iters, knn_count=K, max_dist=WORKLOAD["max_dist"], lr=WORKLOAD["lr"], svn_full_grad=WORKLOAD["svn_full_grad"])
        src = _subsample(pb, ns)
        t0 = time.time()
        call(prm, src)
        return time.time() - t0

    ns1, ns2 = 256, 512  # measured on the 16-core GPU box: 10 steps + 3 warm-ups of this sample take about 40 s
    for _ in range(max(args.warmup, 1)):
        run(ns1, 1)
    vals, desc = [], ""
    for _ in range(args.steps):  # a step = one bounded sample (4 short runs) -> one extrapolated scan time
        t_scan, desc = _sampled_scan_time(run, ns_full, I, ns1, ns2)
        vals.append(t_scan)
    t_scan = float(np.mean(vals))
    val = 1.0 / t_scan
    sample = f"{kind} (libtorch CPU, device-swapped reference sources): all {P} particles, full {len(pb.target)}-point map; " + desc
    line = dict(metric=METRIC, value=val, unit="scans/sec", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t_scan * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                impl="reference", config=dict(workload="configs[1]: 64-beam synthetic scan, 1000 particles", n_s=ns_full, n_t=len(pb.target), **WORKLOAD),
                cpu_baseline=dict(value=val, unit="scans/sec", cores=cores, kind=kind, sample=sample),
                e2e=dict(value=val, unit="scans/sec", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--particles", type=int, default=WORKLOAD["particles"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--config", type=int, default=1, choices=(1, 2),
                    help="BASELINE.json configs index: 1 = the contract workload (default); 2 = 128-beam ~260k-point scan, 4096 particles "
                         "(the sharded parity/scaling case; not the driver's bench line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    WORKLOAD["particles"] = args.particles
    global METRIC
    workload_name = "configs[1]: 64-beam synthetic scan, 1000 particles, particle-sharded when N>1"
    if args.config == 2:
        WORKLOAD.update(sensor="128", particles=4096 if args.particles == 1000 else args.particles)
        METRIC = "scans/sec at 4096 particles (128-beam ~260k-pt synthetic scan, K=100, 30 SVN iterations)"
        workload_name = "configs[2]: 128-beam synthetic scan, 4096 particles, particle-sharded when N>1"
        args.no_variants = True
        args.no_cpu_baseline = True
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import svn_icp_b200 as sv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    P, I, K = WORKLOAD["particles"], WORKLOAD["iterations"], WORKLOAD["K"]
    pb, gen_s = make_problem(P)
    n_s, n_t = len(pb.source), len(pb.target)
    rng = np.random.default_rng(123)
    from svn_icp_b200 import synth
    particles = [synth.init_particles(P, rng) for _ in range(W + args.steps + 2)]

    prm = sv.SteinICPParam(iterations=I, KNN_count=K, max_dist=WORKLOAD["max_dist"], lr=WORKLOAD["lr"], SVN_full_grad=WORKLOAD["svn_full_grad"],
                           check_early_stop=False)
    icp = sv.SVNICP(prm, particles[0], device=local_rank)
    if world > 1:
        uid = [sv.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        icp.init_sharding(uid[0], rank, world)
    stream = torch.cuda.current_stream()
    icp.set_stream(stream.cuda_stream)

    src_dev = torch.from_numpy(pb.source).to(dev)
    tgt_dev = torch.from_numpy(pb.target).to(dev)
    src_pin = torch.from_numpy(pb.source).pin_memory()
    tgt_pin = torch.from_numpy(pb.target).pin_memory()

    def scan_device(i):
        icp.add_cloud_device(src_dev.data_ptr(), n_s, tgt_dev.data_ptr(), n_t, particles[i])
        icp.set_initial_mean(pb.R0, pb.t0)
        assert icp.stein_align() == sv.ALIGN_SUCCESS
        return icp.get_transformation(), icp.get_distribution(), icp.get_cov_matrix(), icp.get_particles()

    def scan_host(i):
        icp.add_cloud_pinned(src_pin.data_ptr(), n_s, tgt_pin.data_ptr(), n_t, particles[i])
        icp.set_initial_mean(pb.R0, pb.t0)
        assert icp.stein_align() == sv.ALIGN_SUCCESS
        return icp.get_transformation(), icp.get_distribution(), icp.get_cov_matrix(), icp.get_particles()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, first, count):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(count):
            out = fn(first + i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for i in range(W):
        scan_device(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, out = timed(scan_device, W, args.steps)
    launches = icp.launch_count() * args.steps
    clocks = sampler.stop() if rank == 0 else None
    scan_host(0)
    ms_e2e, out_h = timed(scan_host, W, args.steps)

    # per-phase device times of one profiled scan (events around each launch group): the roofline numerators
    icp.set_profiling(True)
    scan_device(W)
    ph = icp.get_phase_times()
    info = icp.get_scan_info()
    prune = icp.get_prune_stats()
    icp.set_profiling(False)
    lo, hi = icp.slice()
    P_g = hi - lo
    hbm_peak, peak_src, sm_max = peaks()
    iters = max(ph["iterations"], 1)
    # algorithmic bytes / flops per iteration per GPU (restated in DESIGN.md)
    b_alg = 16.0 * n_s * (1 + K) + 156.0 * P_g
    f_alg = float(P_g) * n_s * (130.0 + 8.0 * K)
    t_pass = (ph["filter_ms"] + ph["gn_ms"]) / iters * 1e-3
    # the HBM-streaming kernel alone: from iteration ~11 on the default path prunes the previous iteration's short lists
    # instead of streaming the K-slot table (k_filter_reuse), so its bytes/time is measured on a scan that streams the
    # table every iteration (SVNICP_FILTER_FULL=1, read at add_cloud)
    os.environ["SVNICP_FILTER_FULL"] = "1"
    icp_ff = sv.SVNICP(prm, particles[0], device=local_rank)
    if world > 1:
        uid = [sv.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        icp_ff.init_sharding(uid[0], rank, world)
    icp_ff.set_stream(stream.cuda_stream)
    icp_ff.set_profiling(True)
    for _ in range(2):
        icp_ff.add_cloud_device(src_dev.data_ptr(), n_s, tgt_dev.data_ptr(), n_t, particles[W])
        icp_ff.set_initial_mean(pb.R0, pb.t0)
        icp_ff.stein_align()
    ph_ff = icp_ff.get_phase_times()
    icp_ff.close()
    del os.environ["SVNICP_FILTER_FULL"]
    t_filter = ph_ff["filter_ms"] / max(ph_ff["iterations"], 1) * 1e-3
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    roofline = dict(bound="hbm", kernel="k_filter + k_gn (correspondence + Gauss-Newton reduction pass, per iteration)",
                    achieved=b_alg / t_pass / 1e9, peak=hbm_peak, unit="GB/s", frac=b_alg / t_pass / 1e9 / hbm_peak, traffic=None,
                    peak_source=peak_src, algorithmic_bytes_per_launch=b_alg, ms_per_launch=t_pass * 1e3,
                    filter_only=dict(achieved=16.0 * n_s * (1 + K) / t_filter / 1e9, frac=16.0 * n_s * (1 + K) / t_filter / 1e9 / hbm_peak,
                                     ms_per_launch=t_filter * 1e3,
                                     note="k_filter streaming the K-slot table every iteration (SVNICP_FILTER_FULL=1 scan)"),
                    fp32=dict(achieved_tflops=f_alg / t_pass / 1e12, peak_tflops=fp32_peak, frac=f_alg / t_pass / 1e12 / fp32_peak,
                              note="brute-force-equivalent flops P*N_s*(130+8K); exact pruning skips most of them"))
    # dram__bytes_{read,write}.sum per launch from the committed ncu --set full captures (profiles/), if present
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        roofline["traffic"] = tr["k_filter_bytes_per_launch"] + tr["k_gn_bytes_per_launch_late"]
        roofline["traffic_detail"] = tr
    except Exception:
        pass

    # secondary variants of the same scan: the reference's own pre-processing (uniform down-sampling
    # 0.5 m then 1.5 m voxels, OdometryPipeline.cpp:559-560) and reference-style early stop (geodeAlpha.yaml:9,17-19)
    variants = {}
    if world == 1 and not args.no_variants:
        ds = synth.uniform_downsample(synth.uniform_downsample(pb.source, 0.5), 1.5)
        ds_dev = torch.from_numpy(np.ascontiguousarray(ds)).to(dev)

        def scan_ds(i):
            icp.add_cloud_device(ds_dev.data_ptr(), len(ds), tgt_dev.data_ptr(), n_t, particles[i])
            icp.set_initial_mean(pb.R0, pb.t0)
            icp.stein_align()
            return icp.get_transformation()

        scan_ds(0)
        ms_ds, mean_ds = timed(scan_ds, 1, args.steps)
        variants["downsampled_source"] = dict(n_s=int(len(ds)), scans_per_sec=args.steps / (ms_ds * 1e-3), ms_per_scan=ms_ds / args.steps,
                                              mean=[float(v) for v in mean_ds])
        prm_es = sv.SteinICPParam(iterations=100, KNN_count=K, max_dist=WORKLOAD["max_dist"], lr=WORKLOAD["lr"],
                                  SVN_full_grad=WORKLOAD["svn_full_grad"], check_early_stop=True, convergence_threshold=5e-4)
        icp_es = sv.SVNICP(prm_es, particles[0], device=local_rank)
        icp_es.set_stream(stream.cuda_stream)

        def scan_es(i):
            icp_es.add_cloud_device(src_dev.data_ptr(), n_s, tgt_dev.data_ptr(), n_t, particles[i])
            icp_es.set_initial_mean(pb.R0, pb.t0)
            icp_es.stein_align()
            return icp_es.get_transformation()

        scan_es(0)
        ms_es, _ = timed(scan_es, 1, args.steps)
        variants["early_stop_thr5e-4_max100"] = dict(scans_per_sec=args.steps / (ms_es * 1e-3), ms_per_scan=ms_es / args.steps,
                                                     iterations_executed=icp_es.iterations_done())
        icp_es.close()
        # the other registration class behind the same interface (class_type = SVGDICP): same scan,
        # same particle count, the reference's shipped SVGD settings (Adam, lr 0.03; stein_icp params)
        prm_gd = sv.SteinICPParam(iterations=I, KNN_count=K, max_dist=WORKLOAD["max_dist"], lr=0.03, optimizer="Adam", check_early_stop=False)
        icp_gd = sv.SVGDICP(prm_gd, particles[0], device=local_rank)
        icp_gd.set_stream(stream.cuda_stream)

        def scan_gd(i):
            icp_gd.add_cloud_device(src_dev.data_ptr(), n_s, tgt_dev.data_ptr(), n_t, particles[i])
            icp_gd.set_initial_mean(pb.R0, pb.t0)
            icp_gd.stein_align()
            return icp_gd.get_transformation()

        scan_gd(0)
        ms_gd, mean_gd = timed(scan_gd, 1, args.steps)
        icp_gd.set_profiling(True)
        scan_gd(1)
        ph_gd = icp_gd.get_phase_times()
        variants["svgd_icp_class_adam"] = dict(scans_per_sec=args.steps / (ms_gd * 1e-3), ms_per_scan=ms_gd / args.steps,
                                               phases_ms_per_scan=ph_gd, mean=[float(v) for v in mean_gd])
        icp_gd.close()

    # The reference ITSELF on this GPU: its unmodified SVNICP sources + its vendored knn.cu against
    # libtorch CUDA (oracle/_ref/libsvnicp_ref_cuda.so, built by oracle/build_ref_cuda.sh), same scan, same particles, in a
    # subprocess (its O(P*N_s) fp64 temporaries peak at ~83 GB; a failure there must not take this bench down).  Also the
    # full-size parity figure: our particles against the reference's on the BASELINE-size problem.
    reference_on_gpu = None
    if world == 1 and not args.no_variants:
        ref_so = os.path.join(ROOT, "oracle", "_ref", "libsvnicp_ref_cuda.so")
        if os.path.exists(ref_so):
            out_npy = os.path.join(ROOT, "gpurun_out", "ref_gpu_particles.npy")
            os.makedirs(os.path.dirname(out_npy), exist_ok=True)
            try:
                r = subprocess.run([sys.executable, "-m", "oracle.ref_gpu_run", str(P), str(I), "0", out_npy], cwd=ROOT, capture_output=True,
                                   text=True, timeout=420)
                reference_on_gpu = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as exc:  # noqa: BLE001
                reference_on_gpu = dict(ok=False, error=f"{type(exc).__name__}: {str(exc)[:200]}")
            if reference_on_gpu.get("ok"):
                icp.add_cloud_device(src_dev.data_ptr(), n_s, tgt_dev.data_ptr(), n_t, pb.init_pose)
                icp.set_initial_mean(pb.R0, pb.t0)
                icp.stein_align()
                ours = icp.get_particles().reshape(6, P)
                theirs = np.load(out_npy)
                reference_on_gpu.update(
                    scans_per_sec=1.0 / reference_on_gpu["seconds_scan"],
                    speedup_device=(args.steps / (ms_dev * 1e-3)) * reference_on_gpu["seconds_scan"],
                    parity_full_size=dict(max_abs_particle_diff=float(np.nanmax(np.abs(ours - theirs))),
                                          max_abs_mean_diff=float(np.max(np.abs(icp.get_transformation() - np.array(reference_on_gpu["mean"])))),
                                          note="our scan vs the reference's own GPU run, same inputs, 30 iterations, all particles"))
        else:
            reference_on_gpu = dict(ok=False, error="oracle/_ref/libsvnicp_ref_cuda.so not built (oracle/build_ref_cuda.sh needs /root/reference)")

    if rank == 0:
        h2d = (n_s + n_t) * 24 + 6 * P * 8
        d2h = (48 + 6 * P) * 8
        line = dict(metric=METRIC, value=args.steps / (ms_dev * 1e-3), unit="scans/sec", n_gpus=world, steps=args.steps, warmup=W,
                    ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32 geometry / f64 reduction+Stein",
                    data="synthetic",
                    config=dict(workload=workload_name, n_s=n_s, n_t=n_t,
                                l2="candidate table 16*N_s*K bytes > 126 MB L2 is re-streamed every iteration; inputs are not L2 resident",
                                particles_per_gpu=P_g, **WORKLOAD),
                    e2e=dict(value=args.steps / (ms_e2e * 1e-3), unit="scans/sec", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=int(launches), roofline=roofline, clocks=clocks,
                    phases_ms_per_scan=ph, scan_info=info, prune_mean_kept=[round(float(x), 2) for x in prune], variants=variants,
                    check=dict(mean=[float(v) for v in out[0]], gt=[float(v) for v in pb.gt_rel]), datagen_s=gen_s)
        if reference_on_gpu is not None:
            line["reference_on_gpu"] = reference_on_gpu
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port(pb)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
