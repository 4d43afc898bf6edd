#!/usr/bin/env python
"""Benchmark of the EqF vision-update hot path: vision-updates/sec at N landmarks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--landmarks 256] [--impl b200|reference]

One *step* = one `processVisionData` call including the IMU integration since the previous image
(10 IMU samples), preceded -- as eqvio_sim does (src/main_sim.cpp:136-142) -- by
`augmentLandmarkStates`, on a synthetic VIOSimulator stream (simdata/, wave trajectory, 200/20 Hz,
pinhole 752x480, fastRiccati, Euclidean chart, discrete lifts, equivariant output).  The default
workload is BASELINE.json configs[1]: N = 256 landmarks, fp64 Sigma, one sequence per GPU.

What is timed (CUDA events, after W warm-up steps, L2 flushed between steps outside the brackets):
  value  updates/s from the device time of each update, measured with CUDA events recorded on the
         filter's own stream from just after the step's pixels are staged in HBM to the last kernel
         of the correction (propagation + preprocessing + correction, the reference's LoopTimer
         labels).  Whole-job: sum over ranks of K / max over ranks of the summed device time.
  e2e    updates/s through the C ABI with HOST buffers, driven by a C++ host loop (eqvio_replay -- the
         reference's host is C++): per step 10x processIMUData, augmentLandmarkStates,
         processVisionData (H2D of pixels / ids / IMU inside), and a D2H read of the state estimate;
         host wall clock per frame (every frame ends synchronised with its estimate on the host), on
         K further frames of the same stream, L2 flushed between frames outside the bracket.
         e2e.python_driver is the same loop driven through the ctypes mirror (CUDA events around each
         synchronised step); with several sequences per GPU only that driver runs.
         e2e.real_data_flow: K more frames through the C++ loop WITHOUT augmentLandmarkStates (how
         eqvio_opt drives the filter on real data: ids are lost / added inside processVisionData).
Multi-GPU (torchrun, one rank per GPU): independent sequences (seed = rank), no collective on the
data path; one NCCL all-gather of the trajectories at the end (timed into e2e).  scaling = weak.

--impl reference times the CPU restatement of the reference's dense Eigen path (oracle/, numpy fp64,
same evaluation order incl. the doubly evaluated gain) on the host cores with the same stream.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


class StdoutGuard:
    """stdout carries the ONE JSON line only: while the benchmark runs, file descriptor 1 points at stderr (NCCL prints its
    version banner to stdout from C, torchrun children inherit chatty libraries); emit() restores it for the line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)


def all_host_threads():
    """Context manager for the all-threads CPU arms.  torchrun exports OMP_NUM_THREADS=1 to its children, which would silently turn
    the reference arm into a one-thread run: when the environment restricts the BLAS pools, lift them to the physical core count;
    otherwise leave the pools at their own default."""
    import contextlib

    if not any(os.environ.get(v) for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")):
        return contextlib.nullcontext()
    try:
        import psutil
        from threadpoolctl import threadpool_limits

        return threadpool_limits(limits=psutil.cpu_count(logical=False) or os.cpu_count() or 1)
    except Exception:
        return contextlib.nullcontext()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--landmarks", type=int, default=256)
    ap.add_argument("--coord", type=int, default=0)
    ap.add_argument("--sequences-per-gpu", type=int, default=1, help="independent sequences (replicas) run concurrently per GPU")
    ap.add_argument("--batched-sequences", type=int, default=16,
                    help="extra leg of the single-GPU, single-sequence run: this many independent sequences replayed concurrently on "
                         "the GPU (BASELINE configs[4]: 16 Monte-Carlo instances per GPU), reported as \"batched\"; 0 = skip")
    ap.add_argument("--no-graph", action="store_true", help="issue per-kernel launches instead of replaying CUDA graphs")
    ap.add_argument("--no-stage", action="store_true", help="chunk factor kernel gathers Sigma[L_c, L_c] itself instead of the TMA tensor copy")
    ap.add_argument("--no-lookahead", action="store_true", help="one in-order downdate launch per chunk (no band / rest split)")
    ap.add_argument("--downdate", default="f64", choices=["f64", "tc"],
                    help="f64: DMMA fp64 downdate (default); tc: tcgen05 split-bf16 operands, fp32 accumulate in TMEM (BASELINE configs[2])")
    ap.add_argument("--device-sim", action="store_true", help="generate the input streams with the device VIOSimulator (eqvio_b200.simulator)")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="updates in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--profile-steps", type=int, default=10, help="extra untimed-for-throughput steps with per-kernel events")
    return ap.parse_args()


def settings_dict(coord):
    """Benchmark filter settings: struct defaults of VIOFilter::Settings with fastRiccati on."""
    return dict(fastRiccati=1, coordinateChoice=coord)


def workload_name(N, coord):
    return (f"VIOSimulator wave, N={N} landmarks, fp64 Sigma (dim {21 + 3 * N}), {('Euclidean', 'InvDepth', 'Normal')[coord]} chart, "
            "fastRiccati, 10 IMU samples per update")


# ---- algorithmic work per update (DESIGN.md "Roofline bookkeeping") -------------------------
def alg_counts(N, n):
    dim, m = 21 + 3 * N, 2 * n
    return dict(
        dim=dim, m=m,
        syrk_flops=float(dim) * dim * m,  # Sigma -= Y^T Y on the lower triangle (2 m dim^2 / 2)
        trail_flops=float(m) * m * m / 3 + float(m) * m * dim,  # Cholesky of S + trsm of W
        prop_bytes=2.0 * 8 * dim * dim,  # read + write Sigma once
        upd_flops=float(m) ** 3 / 3 + float(m) * m * dim + 2.0 * m * dim * dim / 2 + 72.0 * dim * dim,
        upd_bytes=8.0 * (4 * dim * dim + 4 * m * dim))


class ClockSampler:
    """nvidia-smi-equivalent sampling of SM clocks and throttle reasons (NVML).  sample() is called from the
    benchmark loop itself every few steps, right after a step's closing event (the GPU is still busy with the
    tail of that step / the L2 flush, and NVML queries stay outside the timed brackets)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def result(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=[], samples=0)
        return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(self.samples))


def oracle_stream(stream, settings_kw):
    """simdata stream -> the oracle's containers (cpu_baseline / reference arm only)."""
    from oracle import eqf
    from oracle.camera import PinholeCamera
    from oracle.liegroups import SE3

    st = eqf.Settings()
    for k, v in settings_kw.items():
        cur = getattr(st, k)
        setattr(st, k, bool(v) if isinstance(cur, bool) else v)
    c = stream.camera
    cam = PinholeCamera(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"])
    init = eqf.VIOState(eqf.VIOSensorState.fromFlat(stream.init_sensor), stream.init_p, stream.init_ids)
    st.cameraOffset = SE3(stream.init_sensor[16:20], stream.init_sensor[20:23])
    return st, cam, init


def time_cpu(stream, settings_kw, warmup, steps, structured=False):
    """The reference's dense evaluation order on the host cores: returns (updates/s, per-stage seconds).
    structured=True times the minimal-flop CPU formulation instead (sparse A and C, one Cholesky of S, Sigma -= Y^T Y),
    so that the GPU speed-up is not credited with purely algorithmic gains."""
    from oracle import eqf

    st, cam, init = oracle_stream(stream, settings_kw)
    flt = eqf.VIOFilter(st, init, 0.0)
    flt.filterState.mirrorLazyEvaluation = not structured
    flt.filterState.structuredEvaluation = structured
    t_total = 0.0
    done = 0
    for k, fr in enumerate(stream.frames[: 1 + warmup + steps]):
        if k == 1 + warmup:
            flt.timing = {"propagation": 0.0, "preprocessing": 0.0, "correction": 0.0}
        t0 = time.perf_counter()
        for row in fr.imu:
            flt.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
        meas = eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, cam)
        flt.augmentLandmarkStates(meas.getIds(), eqf.VIOState(None, fr.provided_p, fr.ids))
        flt.processVisionData(meas)
        flt.stateEstimate()
        if k >= 1 + warmup:
            t_total += time.perf_counter() - t0
            done += 1
    return done / t_total, dict(flt.timing), done


def cpu_variants(stream, settings_kw, warmup, dense_ups):
    """Side figures of cpu_baseline: the dense reference order on ONE thread (the reference build never enables OpenMP,
    so its Eigen GEMMs are single-core) and the structured + Cholesky formulation on all threads.  Bounded samples."""
    out = {}
    n1 = int(max(2, min(10, 4.0 * dense_ups / 8.0)))  # ~4 s assuming 1 thread is <= 8x slower
    try:
        from threadpoolctl import threadpool_limits

        with threadpool_limits(limits=1):
            ups1, _, d1 = time_cpu(stream, settings_kw, 1, n1)
        out["single_thread"] = dict(value=ups1, cores=1, sample=f"{d1} updates, dense reference order")
    except Exception as e:  # threadpoolctl missing: say so instead of guessing
        out["single_thread"] = dict(value=None, note=repr(e))
    ns = int(max(3, min(30, 4.0 * dense_ups)))
    try:
        from threadpoolctl import threadpool_limits

        # one thread as well: the pair (single_thread, structured_cholesky) isolates the algorithmic gain; with all threads the
        # skinny products of this form run slower than on one (OpenBLAS threading overhead, two BLAS pools under numpy + scipy)
        with threadpool_limits(limits=1):
            ups_s, st_s, ds = time_cpu(stream, settings_kw, 1, ns, structured=True)
        out["structured_cholesky"] = dict(value=ups_s, cores=1, sample=f"{ds} updates; block-structured A / C, one Cholesky of S, "
                                          "Sigma -= Y^T Y (F_alg of SURVEY 8d; numpy + LAPACK), same results to 1e-11; at N = 256 "
                                          "Python / numpy overheads are a large part of it",
                                          stage_ms={k: 1000.0 * v / ds for k, v in st_s.items()})
    except Exception as e:
        out["structured_cholesky"] = dict(value=None, note=repr(e))
    return out


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world, guard):
    if rank != 0:
        return
    from simdata import SimConfig, record_stream

    N = args.landmarks
    skw = settings_dict(args.coord)
    frames = 1 + args.warmup + args.steps
    stream = record_stream(SimConfig.benchmark(N, 0, duration=20.0 if frames <= 399 else float((frames + 1) // 20 + 2)), frames)
    with all_host_threads():
        ups, stages, done = time_cpu(stream, skw, args.warmup, args.steps)
        cores = blas_threads()
    sample = f"{done} consecutive updates of the same stream after {args.warmup} warm-up updates"
    variants = cpu_variants(stream, skw, min(args.warmup, 2), ups)
    line = dict(impl="reference", metric="vision-updates/sec", value=ups, unit="updates/s", n_gpus=args.gpus, steps=done,
                warmup=args.warmup, ms_per_step=1000.0 / ups, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload=workload_name(N, args.coord), landmarks=N, impl_detail="oracle port of the dense Eigen path "
                            "(numpy fp64 + OpenBLAS, reference evaluation order incl. doubly evaluated gain); the reference "
                            "itself cannot be built here (Eigen3/OpenCV/yaml-cpp absent)"),
                cpu_baseline=dict(value=ups, unit="updates/s", cores=cores, kind="port", sample=sample,
                                  stage_ms={k: 1000.0 * v / done for k, v in stages.items()}, **variants),
                e2e=dict(value=ups, unit="updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    guard.emit(line)


def run_b200(args, rank, local_rank, world, guard):
    import torch

    import __graft_entry__ as entry

    entry.build()
    import eqvio_b200 as eb
    from eqvio_b200.replicas import gather_trajectories, shard_instances, trajectory_row
    from simdata import SimConfig, record_stream

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    N, K, W, P, R = args.landmarks, args.steps, args.warmup, args.profile_steps, args.sequences_per_gpu
    skw = settings_dict(args.coord)
    K3 = K if 1 + W + 3 * K + P <= 399 else 0  # third pass (real-data flow) only when the 20 s lap has frames left
    total_frames = 1 + W + 2 * K + K3 + P  # warm-up | timed (Python driver) | C++ host loop | C++ loop, real-data flow | per-kernel profile
    # the reference's simulated lap is 20 s (399 updates after the t = 0 image); longer runs keep circling the same trajectory
    duration = 20.0 if total_frames <= 399 else float((total_frames + 1) // 20 + 2)
    # weak scaling: every rank owns R independent sequences (instance id = seed), contiguous blocks of ids
    total_instances = R * world
    mine = shard_instances(total_instances, world, rank)
    if args.device_sim:  # IMU / vision streams of all local instances from one device launch
        from eqvio_b200.simulator import DeviceSimulator

        dsim = DeviceSimulator([SimConfig.benchmark(N, inst, duration=duration) for inst in mine], device=local_rank)
        streams = dsim.record_streams(total_frames)
        dsim.close()
    else:
        streams = [record_stream(SimConfig.benchmark(N, inst, duration=duration), total_frames) for inst in mine]
    st = eb.Settings(**skw)
    filters = []
    for sm in streams:
        xi0 = eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids)
        flt = eb.VIOFilter(st, xi0, 0.0, capacity=N + 8, device=local_rank)
        flt.enableStageTiming(True)
        if args.no_graph:
            flt.setTuning(graph=0)
        if args.no_lookahead:
            flt.setTuning(lookahead=0)
        if args.no_stage:
            flt.setTuning(stageS=0)
        if args.downdate == "tc":
            flt.setTuning(downdate=1)
        filters.append(flt)
    cam = eb.Camera(**streams[0].camera)
    flush_buf = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    pool = None
    if len(filters) > 1:
        # one host thread per sequence: every C-ABI call releases the GIL, the filters own independent streams, so
        # the host-side work of the replicas overlaps across cores and their kernels overlap on the GPU
        from concurrent.futures import ThreadPoolExecutor

        pool = ThreadPoolExecutor(max_workers=min(len(filters), max(1, (os.cpu_count() or 2) - 1)))
        torch.cuda.set_device(local_rank)

    def step_one(i, k):
        flt, fr = filters[i], streams[i].frames[k]
        flt.processIMUArray(fr.imu)
        flt.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        return flt.stateEstimate()

    def step(k):
        """One vision update of every local sequence; returns the state estimates."""
        if pool is None:
            return [step_one(0, k)]
        return list(pool.map(lambda i: step_one(i, k), range(len(filters))))

    def h2d_bytes(fr):
        return fr.imu.nbytes + fr.y.nbytes + 2 * 4 * len(fr.ids) + 8 * 13  # IMU rows, pixels, index maps, frame header scalars

    step(0)  # t = 0 image: augments only
    for k in range(1, 1 + W):
        step(k)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if dist:  # bring the collective path up (communicator, buffers) before anything is timed
        gather_trajectories({inst: np.zeros((K, 11)) for inst in mine}, total_instances, device=torch.device("cuda", local_rank))
    launches0 = sum(f.launchCount() for f in filters)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    dev_ms = 0.0
    stage_acc = dict(propagation=0.0, preprocessing=0.0, correction=0.0)
    traj = {inst: np.zeros((K, 11)) for inst in mine}
    h2d = d2h = 0
    wall_in = 0.0
    for kk in range(K):
        k = 1 + W + kk
        if flush_buf is not None:
            flush_buf.fill_(kk & 0xFF)
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev[kk][0].record()
        ests = step(k)
        ev[kk][1].record()
        wall_in += time.perf_counter() - t0
        if kk % max(1, K // 8) == 0:
            if flush_buf is not None:
                flush_buf.fill_(kk & 0xFF)  # keep the GPU busy while NVML is queried
            sampler.sample()
        per_filter = []
        for flt in filters:
            sm_ = flt.stageMs()
            for key in stage_acc:
                stage_acc[key] += sm_[key] / len(filters)
            per_filter.append(sm_["propagation"] + sm_["preprocessing"] + sm_["correction"])
        # several sequences per GPU run concurrently on their own streams: their device times overlap, so the step is
        # charged its host-side bracket (all sequences done) instead of a sum of per-filter device times
        dev_ms += max(per_filter) if len(filters) == 1 else 1000.0 * (time.perf_counter() - t0)
        for inst, est, sm in zip(mine, ests, streams):
            traj[inst][kk] = trajectory_row(sm.frames[k].stamp, est)
            h2d += h2d_bytes(sm.frames[k])
            d2h += 8 * (23 + 3 * len(est.ids)) + 8 * 3 * N + 4 * (2 + N)  # state estimate + gate scalars + flag/status words
    torch.cuda.synchronize()
    launches = sum(f.launchCount() for f in filters) - launches0
    e2e_ms = sum(a.elapsed_time(b) for a, b in ev)
    # the single collective of the path: all-gather of the trajectories at the end (timed into e2e)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist:
        dist.barrier()  # the slowest rank's steps are already charged through the max over ranks: do not charge the skew twice
        torch.cuda.synchronize()
    g0.record()
    all_traj = gather_trajectories(traj, total_instances, device=torch.device("cuda", local_rank) if dist else None)
    g1.record()
    torch.cuda.synchronize()
    if dist:
        e2e_ms += g0.elapsed_time(g1)
    assert all_traj.shape == (total_instances, K, 11) and np.isfinite(all_traj).all()

    # e2e through a C++ host loop over the C ABI (eqvio_replay: the reference's host is C++): the next K frames, per frame
    # processIMUData x 10, augmentLandmarkStates, processVisionData, stateEstimate on host buffers, host wall clock per
    # frame (every frame ends synchronised with its estimate on the host), L2 flushed between frames outside the bracket.
    # Stage-event recording is off here (it is instrumentation for `value`).
    cpp_ms = 0.0
    real_ms = 0.0
    if len(filters) == 1:
        filters[0].enableStageTiming(False)
        fms, est_s = filters[0].replay(streams[0].frames[1 + W + K:1 + W + 2 * K], cam, flushBytes=0 if args.no_l2_flush else 256 << 20)
        assert np.isfinite(est_s).all()
        cpp_ms = float(fms.sum())
        if dist:
            cpp_ms += g0.elapsed_time(g1)  # the one collective of the path counts into e2e
        if K3:
            # the same loop WITHOUT augmentLandmarkStates, as eqvio_opt drives the filter on real data: lost ids are pruned and
            # new ids added inside processVisionData (planned frames)
            class _NoAug:
                def __init__(self, fr):
                    self.stamp, self.imu, self.ids, self.y, self.provided_p = fr.stamp, fr.imu, fr.ids, fr.y, None

            fms3, est3 = filters[0].replay([_NoAug(fr) for fr in streams[0].frames[1 + W + 2 * K:1 + W + 2 * K + K3]], cam,
                                           flushBytes=0 if args.no_l2_flush else 256 << 20)
            assert np.isfinite(est3).all()
            real_ms = float(fms3.sum())
        filters[0].enableStageTiming(True)
    elif len(filters) > 1:
        # several sequences per GPU: one C++ host thread per sequence (eqvio_replay_batch), wall clock over the whole batch of
        # R x K frames; no L2 flush between frames here (the sequences run concurrently; R covariance pairs exceed L2 anyway
        # from R ~ 12 at N = 256)
        for flt_ in filters:
            flt_.enableStageTiming(False)
        _, est_b, wall_b = eb.replayBatch(filters, [sm_.frames[1 + W + K:1 + W + 2 * K] for sm_ in streams], cam)
        for flt_ in filters:
            flt_.enableStageTiming(True)
        assert np.isfinite(est_b).all()
        cpp_ms = float(wall_b)
    if dist:
        dist.barrier()

    # per-kernel profile on extra steps of the first sequence (event pairs around every launch of a class; the
    # graph path is off while profiling)
    flt = filters[0]
    flt.enableKernelProfile(True)
    flt.kernelProfile(reset=True)
    prof_stage = dict(propagation=0.0, preprocessing=0.0, correction=0.0)
    nprof = 0
    for k in range(1 + W + 2 * K + K3, 1 + W + 2 * K + K3 + P):
        if flush_buf is not None:
            flush_buf.fill_(1)
            torch.cuda.synchronize()
        step(k)
        for key, v in flt.stageMs().items():
            prof_stage[key] += v
        nprof += 1
    prof = flt.kernelProfile(reset=True)
    flt.enableKernelProfile(False)
    n_meas = len(streams[0].frames[1 + W].ids)
    n_state = flt.numLandmarks()

    # Batched leg (single GPU, single-sequence run only): B independent sequences -- the Monte-Carlo instances of BASELINE
    # configs[4], 16 per GPU -- replayed concurrently through the same C ABI with HOST buffers, one C++ host thread per sequence
    # (eqvio_replay_batch), wall clock over the whole batch.  A reported side figure: `value` / `e2e` stay the single sequence.
    batched = None
    B = args.batched_sequences
    if world == 1 and R == 1 and B > 1:
        Kb, Wb = min(K, 40 if N <= 256 else 15), 3
        bstreams = [record_stream(SimConfig.benchmark(N, 1000 + b, duration=20.0), 1 + Wb + Kb) for b in range(B)]
        bfilters = []
        for sm_ in bstreams:
            bf = eb.VIOFilter(st, eb.VIOState(eb.VIOSensorState.fromFlat(sm_.init_sensor), sm_.init_p, sm_.init_ids), 0.0, capacity=N + 8,
                              device=local_rank)
            if args.no_graph:
                bf.setTuning(graph=0)
            bfilters.append(bf)
        eb.replayBatch(bfilters, [sm_.frames[:1 + Wb] for sm_ in bstreams], cam)  # t = 0 image + warm-up (graph capture)
        l0 = sum(f_.launchCount() for f_ in bfilters)
        _, est_bb, wall_bb = eb.replayBatch(bfilters, [sm_.frames[1 + Wb:] for sm_ in bstreams], cam)
        assert np.isfinite(est_bb).all()
        batched = dict(sequences=B, steps_per_sequence=Kb, value=B * Kb / (wall_bb * 1e-3), unit="updates/s", ms_per_batch_step=wall_bb / Kb,
                       gpu_launches=int(sum(f_.launchCount() for f_ in bfilters) - l0),
                       timing="host wall clock over the batch, host buffers in and state estimates out (same contract as e2e)")
        for bf in bfilters:
            bf.close()

    t = torch.tensor([dev_ms, e2e_ms, cpp_ms, real_ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, cpp_ms_max, real_ms_max = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # fp64 has no entry in MEASURED_PEAKS.json (tcgen05 has no fp64 kind; the path runs on the FP64
        # DMMA pipe): calibrate a cuBLAS DGEMM here and use it as the tensor-bound denominator.
        a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
        b = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        f64_peak = 2.0 * 4096**3 / (best * 1e-3) / 1e12

        cnt = alg_counts(n_state, n_meas)
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        kern = {}
        for name, d in prof.items():
            if d["launches"] == 0:
                continue
            per_update_ms = d["ms"] / max(nprof, 1)
            e = dict(ms_per_update=per_update_ms, launches_per_update=d["launches"] / max(nprof, 1),
                     avg_launch_us=1000.0 * d["ms"] / d["launches"])
            if name == "downdate" and args.downdate == "tc":
                # six bf16 products per fp64-equivalent product; the tile kernel is bound by the fp64 Sigma traffic
                # (read + write of the lower tiles and their mirrors), not by the tensor pipe
                e.update(bound="hbm", achieved=2.0 * 8 * cnt["dim"] ** 2 * (cnt["m"] / 64.0) / (per_update_ms * 1e-3) / 1e9, peak=hbm_peak,
                         unit="GB/s", tensor_tflops=6.0 * cnt["syrk_flops"] / (per_update_ms * 1e-3) / 1e12,
                         tensor_peak_tflops=peaks.get("bf16_tflops", 1590.0))
            elif name == "downdate":
                e.update(bound="tensor", achieved=cnt["syrk_flops"] / (per_update_ms * 1e-3) / 1e12, peak=f64_peak, unit="TFLOP/s")
            elif name == "chunk_factor":
                e.update(bound="latency", note="64 sequential pivots per chunk; fp64 CUDA-core work, one 4x4 register tile per thread")
            elif name == "chol_trail":
                e.update(bound="tensor", achieved=cnt["trail_flops"] / (per_update_ms * 1e-3) / 1e12, peak=f64_peak, unit="TFLOP/s")
            elif name == "prop_ll":
                e.update(bound="hbm", achieved=cnt["prop_bytes"] / (per_update_ms * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s")
            if "achieved" in e:
                e["frac"] = e["achieved"] / e["peak"]
            kern[name] = e
        dom = max(kern, key=lambda k_: kern[k_]["ms_per_update"]) if kern else None
        dom_r = max((k_ for k_ in kern if "achieved" in kern[k_]), key=lambda k_: kern[k_]["ms_per_update"], default=None)
        roofline = None
        if dom_r:
            d = kern[dom_r]
            tr = traffic.get(dom_r, {}).get(str(N))
            roofline = dict(kernel=dom_r, bound=d["bound"], achieved=d["achieved"], peak=d["peak"], unit=d["unit"], frac=d["frac"],
                            traffic=tr, avg_launch_us=d["avg_launch_us"],
                            peak_source=(hbm_src if d["bound"] == "hbm" else
                                         f"cuBLAS DGEMM 4096^3 measured in this run ({f64_peak:.1f} TFLOP/s); MEASURED_PEAKS.json "
                                         "has no fp64 figure"),
                            dominant_by_time=dom, kernels=kern, profile_stage_ms={k_: v / max(nprof, 1) for k_, v in prof_stage.items()},
                            update=dict(flops=cnt["upd_flops"], bytes=cnt["upd_bytes"],
                                        flops_frac=cnt["upd_flops"] * K * R / (dev_ms / 1e3) / 1e12 / f64_peak,
                                        hbm_frac=cnt["upd_bytes"] * K * R / (dev_ms / 1e3) / 1e9 / hbm_peak))
        value = world * R * K / (dev_ms_max * 1e-3)
        e2e_py = world * R * K / (e2e_ms_max * 1e-3)
        e2e = world * R * K / (cpp_ms_max * 1e-3) if cpp_ms_max > 0 else e2e_py
        value_note = None
        if R > 1 and cpp_ms_max > 0:
            # concurrent sequences: per-filter device brackets overlap, and the Python-threaded loop is GIL-bound; the C++ batch is
            # the only run that shows what the GPU sustains, so it also stands for `value` (its 13 KB uploads per update included)
            value = max(value, e2e)
            value_note = "R > 1: wall clock of the C++ batch (one host thread per sequence), uploads included"
        line = dict(metric="vision-updates/sec", value=value, unit="updates/s", n_gpus=world, steps=K, warmup=W,
                    ms_per_step=dev_ms_max / K, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=workload_name(N, args.coord), landmarks=N, measured_per_update=n_meas, state_dim=cnt["dim"],
                                sequences_per_gpu=R, l2="not flushed" if args.no_l2_flush else
                                "flushed between steps (256 MiB write) outside the per-step event brackets",
                                value_timing="CUDA events on each filter's stream around the device work of processVisionData "
                                "(frame upload -> status download; + augmentLandmarkStates kernels); per step the slowest of the "
                                "GPU's concurrent sequences counts",
                                launch_mode="per-kernel launches" if args.no_graph else "steady frames replayed as a cached CUDA graph",
                                parallelism=f"replicas: {R} sequence(s) per GPU x {world} GPU(s), no data-path collective, "
                                "one all-gather of trajectories at the end"),
                    value_note=value_note,
                    e2e=dict(value=e2e, unit="updates/s", h2d_bytes_per_step=h2d // K, d2h_bytes_per_step=d2h // K,
                             ms_per_step=(cpp_ms_max if cpp_ms_max > 0 else e2e_ms_max) / K,
                             driver=("C++ host loop over the C ABI (eqvio_replay), host wall clock per synchronised frame" if R == 1 else
                                     "one C++ host thread per sequence over the C ABI (eqvio_replay_batch), wall clock of the batch") if cpp_ms_max > 0
                             else "Python ctypes driver, CUDA events around each synchronised step",
                             real_data_flow=(dict(value=world * R * K3 / (real_ms_max * 1e-3), ms_per_step=real_ms_max / K3,
                                                  note="same C++ loop without augmentLandmarkStates: ids lost / added inside "
                                                  "processVisionData (eqvio_opt's flow)") if real_ms_max > 0 else None),
                             python_driver=dict(value=e2e_py, ms_per_step=e2e_ms_max / K, host_ms_per_step=1000.0 * wall_in / K,
                                                timing="CUDA events around each synchronised step")),
                    gpu_launches=int(launches), launches_per_step=launches / K,
                    stage_ms={k_: v / K for k_, v in stage_acc.items()},
                    stage_ms_note=("per-stage event brackets of plain launches" if args.no_graph else
                                   "the timed steps replay one CUDA graph per update: a single bracket, booked under correction; the "
                                   "propagation / preprocessing / correction split of the same update with plain launches is "
                                   "roofline.profile_stage_ms"),
                    clocks=sampler.result(), roofline=roofline)
        if batched:
            line["batched"] = batched
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is a single-GPU-run figure (rank 0 at N = 1 only)
            est_s = 17.0 * cnt["dim"] ** 3 / 50e9 + 0.02  # ~17 dim^3 flops of the dense path at a conservative 50 GFLOP/s
            sample_n = args.cpu_sample or int(max(3, min(K, 20.0 / est_s)))
            with all_host_threads():
                ups, stages, done = time_cpu(streams[0], skw, min(W, 2), sample_n)
                ncores = blas_threads()
            line["cpu_baseline"] = dict(value=ups, unit="updates/s", cores=ncores, kind="port",
                                        sample=f"{done} consecutive updates of sequence 0 (same inputs) after {min(W, 2)} warm-up "
                                        "updates; oracle port of the dense Eigen path in the reference's evaluation order "
                                        "(numpy + OpenBLAS; includes ~20 ms/update of Python overhead)",
                                        stage_ms={k_: 1000.0 * v / done for k_, v in stages.items()})
            line["cpu_baseline"].update(cpu_variants(streams[0], skw, min(W, 2), ups))
        guard.emit(line)
    for f in filters:
        f.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    guard = StdoutGuard()
    if args.impl == "reference":
        run_reference(args, rank, world, guard)
    else:
        run_b200(args, rank, local_rank, world, guard)


if __name__ == "__main__":
    main()
