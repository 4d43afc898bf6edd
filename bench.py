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
         e2e.python_driver is the same loop driven through the ctypes mirror; e2e.real_data_flow the
         C++ loop WITHOUT augmentLandmarkStates (eqvio_opt's flow on real data).
  sweep  (single-GPU default run only) the same measurement, shortened, at N = 64 and N = 1024 -- north_star's sizes --
         with per-kernel roofline fractions, the parity error against the oracle after the warm-up updates and a CPU figure.
Multi-GPU (torchrun, one rank per GPU): independent sequences (seed = rank), no collective on the
data path; the single collective of the path is the all-gather of the trajectories that closes a simulated lap
(399 updates, main_sim.cpp:128-184): it is timed (e2e.collective_ms) and charged to e2e in proportion to the K
updates timed, K / 399 of it.  `batched` = BASELINE configs[4]: 16 noisy Monte-Carlo instances per GPU
(128 on 8 GPUs) replayed concurrently, poses gathered at the end.  scaling = weak.

--impl reference times the reference's CPU path: oracle/cpu_update.c, a C restatement of its dense linear algebra in
its own evaluation order (incl. the doubly evaluated gain) on the numpy-bundled OpenBLAS, all host threads; operands
come from the numpy oracle outside the timed region and every C result is checked against the oracle's.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LAP_UPDATES = 399  # vision updates of the reference's simulated 20 s lap (20 Hz, the t = 0 image only augments)


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


def host_threads():
    try:
        import psutil

        return psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        return os.cpu_count() or 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--landmarks", type=int, default=256)
    ap.add_argument("--coord", type=int, default=0)
    ap.add_argument("--riccati", choices=["fast", "accurate", "discrete"], default="fast",
                    help="Riccati variant of the filters (default fast = the benchmark workload; the others are side measurements, "
                    "use with --no-cpu-baseline --no-sweep --batched-sequences 0)")
    ap.add_argument("--sequences-per-gpu", type=int, default=1, help="independent sequences (replicas) run concurrently per GPU")
    ap.add_argument("--batched-correction", type=int, default=0, choices=[0, 2],
                    help="correction form of the batched leg's filters: 0 = sequential chunks (default there: with 16 sequences "
                         "sharing one GPU the throughput is bound by SM occupancy, and the block sweep's followers hold SMs while they "
                         "wait for their diagonal step), 2 = block sweep (the single-sequence default)")
    ap.add_argument("--batched-sequences", type=int, default=16,
                    help="extra leg: this many independent NOISY sequences per GPU replayed concurrently (BASELINE configs[4]: 16 "
                         "Monte-Carlo instances per GPU), reported as \"batched\"; 0 = skip")
    ap.add_argument("--no-sweep", action="store_true", help="skip the N = 64 / N = 1024 side measurements of the default single-GPU run")
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


RICCATI = "fast"  # --riccati: side measurements of the per-sample Riccati variants (not the headline workload)


def settings_dict(coord):
    """Benchmark filter settings: struct defaults of VIOFilter::Settings with fastRiccati on."""
    if RICCATI == "accurate":  # the struct default: integrateRiccatiStateAccurate per IMU sample
        return dict(fastRiccati=0, coordinateChoice=coord)
    if RICCATI == "discrete":
        return dict(fastRiccati=0, useDiscreteStateMatrix=1, coordinateChoice=coord)
    return dict(fastRiccati=1, coordinateChoice=coord)


def workload_name(N, coord):
    return (f"VIOSimulator wave, N={N} landmarks, fp64 Sigma (dim {21 + 3 * N}), {('Euclidean', 'InvDepth', 'Normal')[coord]} chart, "
            + {"fast": "fastRiccati", "accurate": "fastRiccati=false (matrix exponential per IMU sample)",
               "discrete": "fastRiccati=false, useDiscreteStateMatrix (per IMU sample)"}[RICCATI] + ", 10 IMU samples per update")


# ---- algorithmic work per update (DESIGN.md "Roofline bookkeeping") -------------------------
def alg_counts(N, n, chunk_rows=64):
    """Flops the sequential-chunk algorithm EXECUTES per update: the downdates m dim^2 (lower triangle: 2 m dim^2 / 2), the
    structured propagation 72 dim^2, and per chunk of r rows its elimination r^3 / 3 and the substitution of dim right-hand
    sides r^2 dim.  The m^3 / 3 + m^2 dim of factoring the full S and solving for the full W do not occur."""
    dim, m = 21 + 3 * N, 2 * n
    chunks = [min(chunk_rows, m - r0) for r0 in range(0, m, chunk_rows)]
    factor_flops = sum(r ** 3 / 3.0 + float(r) * r * dim for r in chunks)
    # block sweep (eqvio_b200/csrc/blockchol.cuh, the default up to 768 rows): 64-row blocks over [S; W^T], W padded to 64-row tiles.
    # Per block k with q = nT - k - 1 blocks to its right and TW tiles of W: diagonal step 64^3 / 3 (factorization) + 2 * 64^3 / 2
    # (look-ahead substitution + symmetric product); panels (q + TW) tiles x 64^3 (substitution, triangular: half of a full
    # product); trailing tiles (q (q + 1) / 2 + TW q + TW (TW + 1) / 2) x 2 * 64^3.  Flops = 2 x FMA.
    nT, TW, b3 = -(-m // 64), -(-(24 + 3 * N) // 64), 64.0 ** 3
    bc_diag = sum(b3 / 3.0 + (b3 if k > 0 else 0.0) for k in range(nT))
    bc_panel = sum((nT - k - 1 + TW) * b3 for k in range(nT))
    bc_trail = sum(((nT - k - 1) * (nT - k) / 2.0 + TW * (nT - k - 1) + TW * (TW + 1) / 2.0) * 2.0 * b3 for k in range(nT))
    block = m <= 768
    return dict(
        block_sweep=block, bc_diag_flops=bc_diag, bc_panel_flops=bc_panel, bc_trail_flops=bc_trail, blocks=nT,
        dim=dim, m=m, chunks=len(chunks),
        syrk_flops=float(dim) * dim * m,
        factor_flops=factor_flops,
        trail_flops=float(m) * m * m / 3 + float(m) * m * dim,  # batch sweep only (Cholesky of S + trsm of W)
        prop_bytes=2.0 * 8 * dim * dim,  # read + write Sigma once
        upd_flops=(bc_diag + bc_panel + bc_trail if block else float(dim) * dim * m + factor_flops) + 72.0 * dim * dim,
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


# ---- CPU arm --------------------------------------------------------------------------------------------------
def oracle_filter(stream, settings_kw):
    """simdata stream -> an oracle VIOFilter + camera (cpu_baseline / reference arm / sweep parity only)."""
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
    flt = eqf.VIOFilter(st, init, 0.0)
    flt.filterState.structuredEvaluation = True  # operands are recorded, not timed: the cheap evaluation of the same update
    return flt, cam


def cpu_arm(stream, settings_kw, warmup, sample_updates, variants=True):
    """CPU baseline on a bounded sample: `sample_updates` consecutive updates of the stream after `warmup` updates.  The numpy
    oracle produces the operands (untimed); oracle/cpu_update.c runs and times the dense stages in the reference's evaluation
    order on all host threads (the headline figure), on one thread (the reference build never enables OpenMP: its Eigen
    products are single-core) and in the block-structured + Cholesky form."""
    from oracle import cpu_baseline as cb

    flt, cam = oracle_filter(stream, settings_kw)
    cb.record_updates(flt, stream.frames[:1 + warmup], cam)
    ups = cb.record_updates(flt, stream.frames[1 + warmup:1 + warmup + sample_updates], cam)
    nthreads = host_threads()
    dense = cb.run_updates(ups, structured=False, threads=nthreads)
    out = dict(value=dense["updates_per_s"], unit="updates/s", cores=dense["threads"], kind="port",
               sample=f"{dense['updates']} consecutive updates of sequence 0 (same inputs) after {warmup} warm-up updates; C restatement "
               "(oracle/cpu_update.c, OpenBLAS) of the reference's dense evaluation order incl. the doubly evaluated gain; "
               "operands from the numpy oracle outside the timed region; O(N) Lie-group / Jacobian glue not included",
               stage_ms=dense["stage_ms"], worst_rel_error_vs_oracle=dense["worst_rel_error_vs_oracle"])
    if variants:
        one = cb.run_updates(ups[:max(1, min(len(ups), 3))], structured=False, threads=1)
        out["single_thread"] = dict(value=one["updates_per_s"], cores=1, stage_ms=one["stage_ms"],
                                    sample=f"{one['updates']} updates, dense reference order (the reference's own build is single-threaded)")
        s1 = cb.run_updates(ups, structured=True, threads=1)
        sa = cb.run_updates(ups, structured=True, threads=nthreads)
        out["structured_cholesky"] = dict(value=s1["updates_per_s"], cores=1, stage_ms=s1["stage_ms"], all_threads_value=sa["updates_per_s"],
                                          all_threads_cores=sa["threads"], worst_rel_error_vs_oracle=s1["worst_rel_error_vs_oracle"],
                                          sample="block-structured A / C, one Cholesky of S, Sigma -= Y^T Y (the flops the GPU path executes)")
    return out


def cpu_sample_size(N, K):
    dim = 21 + 3 * N
    est_s = 17.0 * dim ** 3 / 200e9 + 0.01  # ~17 dim^3 flops of the dense path at a conservative 200 GFLOP/s on all threads
    return int(max(2, min(K, 10.0 / est_s)))


def run_reference(args, rank, world, guard):
    if rank != 0:
        return
    from simdata import SimConfig, record_stream

    N = args.landmarks
    skw = settings_dict(args.coord)
    steps = min(args.steps, max(2, cpu_sample_size(N, args.steps) * 2))  # each step is one update of a bounded sample
    frames = 1 + args.warmup + steps
    stream = record_stream(SimConfig.benchmark(N, 0, duration=20.0 if frames <= 399 else float((frames + 1) // 20 + 2)), frames)
    cb = cpu_arm(stream, skw, args.warmup, steps)
    ups = cb["value"]
    line = dict(impl="reference", metric="vision-updates/sec", value=ups, unit="updates/s", n_gpus=args.gpus, steps=steps,
                warmup=args.warmup, ms_per_step=1000.0 / ups, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload=workload_name(N, args.coord), landmarks=N, impl_detail="C restatement of the reference's dense Eigen "
                            "path (oracle/cpu_update.c on the numpy-bundled OpenBLAS, reference evaluation order incl. the doubly "
                            "evaluated gain, all host threads); the reference itself cannot be built here (Eigen3/OpenCV/yaml-cpp absent)"),
                cpu_baseline=cb,
                e2e=dict(value=ups, unit="updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    guard.emit(line)


# ---- GPU arm --------------------------------------------------------------------------------------------------
def fp64_peak(torch):
    """cuBLAS DGEMM 4096^3 as the fp64 denominator (MEASURED_PEAKS.json has HBM and bf16 only; tcgen05 has no fp64 kind, the
    path runs on the FP64 DMMA pipe like cuBLAS does)."""
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
    return 2.0 * 4096 ** 3 / (best * 1e-3) / 1e12, best


def kernel_rooflines(prof, nprof, cnt, N, hbm_peak, f64_peak, traffic, tc=False, bf16_peak=1590.0):
    """Per kernel class: time per update, launches, and achieved / peak of the bound that applies -- flops for the factor and
    downdate kernels (fp64 pipe, denominator = measured DGEMM), bytes for the propagation (HBM)."""
    kern = {}
    for name, d in prof.items():
        if d["launches"] == 0:
            continue
        ms = d["ms"] / max(nprof, 1)
        e = dict(ms_per_update=ms, launches_per_update=d["launches"] / max(nprof, 1), avg_launch_us=1000.0 * d["ms"] / d["launches"])
        if name == "downdate" and tc:
            e.update(bound="hbm", achieved=2.0 * 8 * cnt["dim"] ** 2 * (cnt["m"] / 64.0) / (ms * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s",
                     tensor_tflops=6.0 * cnt["syrk_flops"] / (ms * 1e-3) / 1e12, tensor_peak_tflops=bf16_peak)
        elif name == "downdate":
            e.update(bound="tensor", achieved=cnt["syrk_flops"] / (ms * 1e-3) / 1e12, peak=f64_peak, unit="TFLOP/s")
        elif name == "chunk_factor":
            e.update(bound="tensor", achieved=cnt["factor_flops"] / (ms * 1e-3) / 1e12, peak=f64_peak, unit="TFLOP/s",
                     note="latency-bound: a chain of 64 dependent pivots per launch; the fraction is of the fp64 (DGEMM) peak")
        elif name == "bc_diag":
            e.update(bound="tensor", achieved=cnt["bc_diag_flops"] / (ms * 1e-3) / 1e12, peak=f64_peak, unit="TFLOP/s",
                     note="latency-bound: the chain of 64 dependent pivots per block (plus the look-ahead products on one SM); the "
                     "fraction is of the fp64 (DGEMM) peak.  Event-bracketed launches of the per-kernel profile run one after the "
                     "other: in the timed steps consecutive diagonal steps overlap (see DESIGN.md 4)")
        elif name == "bc_panel":
            e.update(bound="tensor", achieved=cnt["bc_panel_flops"] / (ms * 1e-3) / 1e12, peak=f64_peak, unit="TFLOP/s")
        elif name == "bc_trail":
            e.update(bound="tensor", achieved=cnt["bc_trail_flops"] / (ms * 1e-3) / 1e12, peak=f64_peak, unit="TFLOP/s")
        elif name == "chol_trail":
            e.update(bound="tensor", achieved=cnt["trail_flops"] / (ms * 1e-3) / 1e12, peak=f64_peak, unit="TFLOP/s")
        elif name == "prop_ll":
            e.update(bound="hbm", achieved=cnt["prop_bytes"] / (ms * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s")
        if "achieved" in e:
            e["frac"] = e["achieved"] / e["peak"]
        e["traffic"] = traffic.get(name, {}).get(str(N))
        kern[name] = e
    return kern


class Sequence:
    """One simulated sequence and its filter."""

    def __init__(self, eb, stream, N, device, args):
        self.stream = stream
        xi0 = eb.VIOState(eb.VIOSensorState.fromFlat(stream.init_sensor), stream.init_p, stream.init_ids)
        self.flt = eb.VIOFilter(eb.Settings(**settings_dict(args.coord)), xi0, 0.0, capacity=N + 8, device=device)
        if args.no_graph:
            self.flt.setTuning(graph=0)
        if args.no_lookahead:
            self.flt.setTuning(lookahead=0)
        if args.no_stage:
            self.flt.setTuning(stageS=0)
        if args.downdate == "tc":
            self.flt.setTuning(downdate=1)

    def step(self, eb, cam, k):
        fr = self.stream.frames[k]
        self.flt.processIMUArray(fr.imu)
        self.flt.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        self.flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        return self.flt.stateEstimate()


def h2d_bytes(fr):
    return fr.imu.nbytes + fr.y.nbytes + 2 * 4 * len(fr.ids) + 8 * 13  # IMU rows, pixels, index maps, frame header scalars


def measure_single(eb, torch, args, N, K, W, P, device, flush_buf, sampler=None, stream=None, parity=False):
    """One sequence on one GPU: `value` (CUDA events per update), `e2e` (C++ host loop, host buffers), the real-data flow and the
    per-kernel profile.  Returns a dict of raw figures (ms sums, counts) for the caller to combine across ranks."""
    from eqvio_b200.replicas import trajectory_row
    from simdata import SimConfig, record_stream

    K3 = K if 1 + W + 3 * K + P <= LAP_UPDATES else 0
    total = 1 + W + 2 * K + K3 + P
    if stream is None:
        stream = record_stream(SimConfig.benchmark(N, 0, duration=20.0 if total <= LAP_UPDATES else float((total + 1) // 20 + 2)), total)
    seq = Sequence(eb, stream, N, device, args)
    flt = seq.flt
    flt.enableStageTiming(True)
    cam = eb.Camera(**stream.camera)
    out = dict(stream=stream, seq=seq)
    for k in range(0, 1 + W):
        seq.step(eb, cam, k)
    torch.cuda.synchronize()
    snap = None
    if parity:  # state after the t = 0 image + W warm-up updates; compared with the oracle AFTER the timed steps (no idle gap before them)
        fs0, est0 = flt.viewEqFState(), flt.stateEstimate()
        snap = (fs0.Sigma.copy(), np.concatenate([est0.sensor.flat(), est0.p.reshape(-1)]), np.asarray(est0.ids).copy())
    launches0 = flt.launchCount()
    graphs0 = flt.graphStats()
    dev_ms = wall_in = 0.0
    stage_acc = dict(propagation=0.0, preprocessing=0.0, correction=0.0)
    traj = np.zeros((K, 11))
    h2d = d2h = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for kk in range(K):
        k = 1 + W + kk
        if flush_buf is not None:
            flush_buf.fill_(kk & 0xFF)
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev[kk][0].record()
        est = seq.step(eb, cam, k)
        ev[kk][1].record()
        wall_in += time.perf_counter() - t0
        if sampler is not None and kk % max(1, K // 8) == 0:
            if flush_buf is not None:
                flush_buf.fill_(kk & 0xFF)  # keep the GPU busy while NVML is queried
            sampler.sample()
        sm_ = flt.stageMs()
        for key in stage_acc:
            stage_acc[key] += sm_[key]
        dev_ms += sm_["propagation"] + sm_["preprocessing"] + sm_["correction"]
        fr = stream.frames[k]
        traj[kk] = trajectory_row(fr.stamp, est)
        h2d += h2d_bytes(fr)
        d2h += 8 * (23 + 3 * len(est.ids)) + 8 * 3 * N + 4 * (2 + N)  # state estimate + gate scalars + flag/status words
    torch.cuda.synchronize()
    py_ms = sum(a.elapsed_time(b) for a, b in ev)
    graphs1 = flt.graphStats()
    out.update(dev_ms=dev_ms, py_ms=py_ms, wall_in=wall_in, stage_acc=stage_acc, traj=traj, h2d=h2d // K, d2h=d2h // K,
               launches=flt.launchCount() - launches0,
               graphs=dict(captured_in_timed_steps=graphs1[0] - graphs0[0], replayed_in_timed_steps=graphs1[1] - graphs0[1]))
    # e2e through the C++ host loop (stage-event recording is instrumentation for `value`: off here)
    flt.enableStageTiming(False)
    fms, est_s = flt.replay(stream.frames[1 + W + K:1 + W + 2 * K], cam, flushBytes=0 if args.no_l2_flush else 256 << 20)
    assert np.isfinite(est_s).all()
    out["cpp_ms"] = float(fms.sum())
    out["real_ms"] = 0.0
    if K3:
        class _NoAug:
            def __init__(self, fr):
                self.stamp, self.imu, self.ids, self.y, self.provided_p = fr.stamp, fr.imu, fr.ids, fr.y, None

        fms3, est3 = flt.replay([_NoAug(fr) for fr in stream.frames[1 + W + 2 * K:1 + W + 2 * K + K3]], cam,
                                flushBytes=0 if args.no_l2_flush else 256 << 20)
        assert np.isfinite(est3).all()
        out["real_ms"] = float(fms3.sum())
    out["K3"] = K3
    if snap is not None:
        # the CUDA path against the oracle after the t = 0 image + W warm-up updates (same inputs, both from the same prior)
        from oracle import eqf

        ofl, ocam = oracle_filter(stream, settings_dict(args.coord))
        for fr in stream.frames[:1 + W]:
            for row in fr.imu:
                ofl.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
            meas = eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, ocam)
            ofl.augmentLandmarkStates(meas.getIds(), eqf.VIOState(None, fr.provided_p, fr.ids))
            ofl.processVisionData(meas)
        oest, S_ref = ofl.stateEstimate(), ofl.viewEqFState().Sigma
        x_o = np.concatenate([oest.sensor.flat(), oest.p.reshape(-1)])
        same_ids = bool(np.array_equal(snap[2], np.asarray(oest.ids)))
        out["parity"] = dict(updates=W, ids_equal=same_ids,
                             sigma_rel_fro=float(np.linalg.norm(snap[0] - S_ref) / np.linalg.norm(S_ref)) if same_ids else None,
                             state_rel_fro=float(np.linalg.norm(snap[1] - x_o) / np.linalg.norm(x_o)) if same_ids else None)
    flt.enableStageTiming(True)
    # per-kernel profile on extra steps (event pairs around every launch of a class; the graph path is off while profiling)
    flt.enableKernelProfile(True)
    flt.kernelProfile(reset=True)
    prof_stage = dict(propagation=0.0, preprocessing=0.0, correction=0.0)
    nprof = 0
    for k in range(1 + W + 2 * K + K3, 1 + W + 2 * K + K3 + P):
        if flush_buf is not None:
            flush_buf.fill_(1)
            torch.cuda.synchronize()
        seq.step(eb, cam, k)
        for key, v in flt.stageMs().items():
            prof_stage[key] += v
        nprof += 1
    out.update(prof=flt.kernelProfile(reset=True), nprof=nprof, prof_stage=prof_stage, n_meas=len(stream.frames[1 + W].ids),
               n_state=flt.numLandmarks())
    flt.enableKernelProfile(False)
    return out


def batched_leg(eb, args, N, B, Kb, Wb, device, first_instance):
    """B independent NOISY Monte-Carlo sequences (BASELINE configs[4]: 16 per GPU) replayed concurrently through the C ABI with HOST
    buffers, one C++ host thread per sequence (eqvio_replay_batch), wall clock over the whole batch.  Returns
    (wall ms, updates, launches, final sensor states (B, 23))."""
    from eqvio_b200.simulator import DeviceSimulator
    from simdata import SimConfig

    # the noisy instances come from the device VIOSimulator: all B streams (IMU, visibility, pixel / IMU noise) in one launch each
    dsim = DeviceSimulator([SimConfig.benchmark(N, first_instance + b, duration=20.0, inputNoise=True, outputNoise=True) for b in range(B)],
                           device=device)
    streams = dsim.record_streams(1 + Wb + Kb)
    dsim.close()
    cam = eb.Camera(**streams[0].camera)
    filters = []
    for sm_ in streams:
        bf = eb.VIOFilter(eb.Settings(**settings_dict(args.coord)), eb.VIOState(eb.VIOSensorState.fromFlat(sm_.init_sensor), sm_.init_p, sm_.init_ids),
                          0.0, capacity=N + 8, device=device)
        if args.no_graph:
            bf.setTuning(graph=0)
        bf.setTuning(correction=args.batched_correction)
        filters.append(bf)
    eb.replayBatch(filters, [sm_.frames[:1 + Wb] for sm_ in streams], cam)  # t = 0 image + warm-up (graph capture)
    l0 = sum(f_.launchCount() for f_ in filters)
    _, est, wall = eb.replayBatch(filters, [sm_.frames[1 + Wb:] for sm_ in streams], cam)
    assert np.isfinite(est).all()
    launches = int(sum(f_.launchCount() for f_ in filters) - l0)
    for bf in filters:
        bf.close()
    return wall, B * Kb, launches, np.ascontiguousarray(est[:, -1, :])


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
    dev = torch.device("cuda", local_rank)
    host_cores = None
    if world > 1 and os.environ.get("EQVIO_BENCH_PIN", "1") != "0":
        # one disjoint block of host cores per rank: the e2e figure is a max over ranks of HOST wall clock, and ranks whose driver
        # threads share or migrate between cores pay for it (round 1: e2e efficiency 0.80 at 8 GPUs with every rank on cores 0-31)
        try:
            allowed = sorted(os.sched_getaffinity(0))
            per = max(1, len(allowed) // world)
            mine_cores = allowed[local_rank * per:(local_rank + 1) * per] or allowed
            os.sched_setaffinity(0, mine_cores)
            host_cores = [mine_cores[0], mine_cores[-1]]
        except Exception:
            host_cores = None
    flush_buf = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sampler = ClockSampler(local_rank)
    total_instances = R * world
    mine = shard_instances(total_instances, world, rank)
    if dist:  # bring the collective path up (communicator, buffers) before anything is timed
        gather_trajectories({inst: np.zeros((K, 11)) for inst in mine}, total_instances, device=dev)

    collective_ms = 0.0
    if R == 1:
        K3 = K if 1 + W + 3 * K + P <= LAP_UPDATES else 0
        total = 1 + W + 2 * K + K3 + P
        duration = 20.0 if total <= LAP_UPDATES else float((total + 1) // 20 + 2)
        if args.device_sim:  # IMU / vision streams from one device launch
            from eqvio_b200.simulator import DeviceSimulator

            dsim = DeviceSimulator([SimConfig.benchmark(N, mine[0], duration=duration)], device=local_rank)
            stream = dsim.record_streams(total)[0]
            dsim.close()
        else:
            stream = record_stream(SimConfig.benchmark(N, mine[0], duration=duration), total)
        if dist:
            dist.barrier()
        m = measure_single(eb, torch, args, N, K, W, P, local_rank, flush_buf, sampler=sampler, stream=stream)
        dev_ms, py_ms, cpp_ms, real_ms = m["dev_ms"], m["py_ms"], m["cpp_ms"], m["real_ms"]
        traj = {mine[0]: m["traj"]}
    else:
        # several sequences per GPU (--sequences-per-gpu): Python-threaded loop for `value` (host bracket of each step: the
        # sequences' device times overlap), C++ batch for e2e
        from concurrent.futures import ThreadPoolExecutor

        total = 1 + W + 2 * K
        streams = [record_stream(SimConfig.benchmark(N, inst, duration=20.0 if total <= LAP_UPDATES else float((total + 1) // 20 + 2)), total)
                   for inst in mine]
        seqs = [Sequence(eb, sm_, N, local_rank, args) for sm_ in streams]
        cam = eb.Camera(**streams[0].camera)
        pool = ThreadPoolExecutor(max_workers=min(len(seqs), max(1, (os.cpu_count() or 2) - 1)))
        for k in range(0, 1 + W):
            list(pool.map(lambda s: s.step(eb, cam, k), seqs))
        torch.cuda.synchronize()
        launches0 = sum(s.flt.launchCount() for s in seqs)
        dev_ms = 0.0
        traj = {inst: np.zeros((K, 11)) for inst in mine}
        h2d = d2h = 0
        for kk in range(K):
            k = 1 + W + kk
            if flush_buf is not None:
                flush_buf.fill_(kk & 0xFF)
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            ests = list(pool.map(lambda s: s.step(eb, cam, k), seqs))
            dev_ms += 1000.0 * (time.perf_counter() - t0)
            if kk % max(1, K // 8) == 0:
                sampler.sample()
            for inst, est, sm_ in zip(mine, ests, streams):
                traj[inst][kk] = trajectory_row(sm_.frames[k].stamp, est)
                h2d += h2d_bytes(sm_.frames[k])
                d2h += 8 * (23 + 3 * len(est.ids)) + 8 * 3 * N + 4 * (2 + N)
        py_ms = dev_ms
        _, est_b, wall_b = eb.replayBatch([s.flt for s in seqs], [sm_.frames[1 + W + K:1 + W + 2 * K] for sm_ in streams], cam)
        assert np.isfinite(est_b).all()
        cpp_ms, real_ms = float(wall_b), 0.0
        m = dict(launches=sum(s.flt.launchCount() for s in seqs) - launches0, h2d=h2d // K, d2h=d2h // K, K3=0,
                 stage_acc=dict(propagation=0.0, preprocessing=0.0, correction=dev_ms), wall_in=dev_ms / 1e3, prof={}, nprof=0,
                 prof_stage={}, n_meas=len(streams[0].frames[1 + W].ids), n_state=seqs[0].flt.numLandmarks(), seq=seqs[0], stream=streams[0])

    # the single collective of the path: all-gather of the trajectories (closes a simulated lap)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist:
        dist.barrier()  # the slowest rank's steps are already charged through the max over ranks: do not charge the skew twice
        torch.cuda.synchronize()
    g0.record()
    all_traj = gather_trajectories(traj, total_instances, device=dev if dist else None)
    g1.record()
    torch.cuda.synchronize()
    if dist:
        collective_ms = g0.elapsed_time(g1)
    assert all_traj.shape == (total_instances, K, 11) and np.isfinite(all_traj).all()

    # batched leg: BASELINE configs[4], B noisy Monte-Carlo instances per GPU, poses gathered at the end
    batched = None
    B = args.batched_sequences
    if R == 1 and B > 1:
        Kb, Wb = min(K, 40 if N <= 256 else 15), 3
        if dist:
            dist.barrier()
        wall_b, upd_b, launch_b, poses = batched_leg(eb, args, N, B, Kb, Wb, local_rank, first_instance=1000 + rank * B)
        tb = torch.tensor([wall_b], dtype=torch.float64, device="cuda")
        gather_ms = 0.0
        if dist:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            pt = torch.from_numpy(poses).to(dev)
            parts = [torch.empty_like(pt) for _ in range(world)]
            gb0, gb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            gb0.record()
            dist.all_gather(parts, pt)  # "NCCL gather of poses" of configs[4]
            gb1.record()
            torch.cuda.synchronize()
            gather_ms = gb0.elapsed_time(gb1)
            assert bool(torch.isfinite(torch.stack(parts)).all())
        wall_max = float(tb[0])
        batched = dict(sequences_per_gpu=B, sequences=B * world, steps_per_sequence=Kb, noisy=True,
                       correction=("sequential chunks (EQVIO_TUNE_CORRECTION=0)" if args.batched_correction == 0 else "block sweep"),
                       value=world * upd_b / ((wall_max + gather_ms) * 1e-3), unit="updates/s", ms_per_batch_step=wall_max / Kb,
                       pose_gather_ms=gather_ms, gpu_launches=launch_b,
                       timing="max over ranks of the host wall clock over the batch (host buffers in, state estimates out: the e2e "
                       "contract) + the all-gather of the final poses")

    t = torch.tensor([dev_ms, py_ms, cpp_ms, real_ms, collective_ms], dtype=torch.float64, device="cuda")
    per_rank = None
    if dist:
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)  # per-rank figures beside the max (a slow rank is a host-side finding, not a kernel one)
        per_rank = [[float(x) for x in r] for r in allr]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, py_ms_max, cpp_ms_max, real_ms_max, coll_ms_max = (float(x) for x in t)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        f64_tflops, f64_ms = fp64_peak(torch)
        sampler.sample()
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        K3 = m["K3"]
        cnt = alg_counts(m["n_state"], m["n_meas"])
        kern = kernel_rooflines(m["prof"], m["nprof"], cnt, N, hbm_peak, f64_tflops, traffic, tc=args.downdate == "tc",
                                bf16_peak=peaks.get("bf16_tflops", 1590.0))
        roofline = None
        if kern:
            dom = max(kern, key=lambda k_: kern[k_]["ms_per_update"])
            d = kern[dom]
            roofline = dict(kernel=dom, bound=d.get("bound"), achieved=d.get("achieved"), peak=d.get("peak"), unit=d.get("unit"),
                            frac=d.get("frac"), traffic=d.get("traffic"), avg_launch_us=d["avg_launch_us"],
                            traffic_source="profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu "
                            "--set full captures named there (scripts/ncu_traffic.py)",
                            peak_source=(hbm_src if d.get("bound") == "hbm" else
                                         f"cuBLAS DGEMM 4096^3 measured in this run ({f64_tflops:.1f} TFLOP/s, {f64_ms:.2f} ms); "
                                         "MEASURED_PEAKS.json has no fp64 figure"),
                            fp64_peak=dict(tflops=f64_tflops, how="torch.matmul fp64 4096^3, best of 5, CUDA events", ms=f64_ms),
                            hbm_peak=dict(gbs=hbm_peak, source=hbm_src),
                            dominant_by_time=dom, kernels=kern,
                            profile_stage_ms={k_: v / max(m["nprof"], 1) for k_, v in m["prof_stage"].items()},
                            update=dict(flops=cnt["upd_flops"], bytes=cnt["upd_bytes"], chunks=cnt["chunks"],
                                        flops_note=("flops the block sweep executes (blockchol.cuh: diagonal steps + panels + trailing tiles of "
                                                    "[S; W^T] and Sigma, 64-row tiles) + 72 dim^2 of the propagation" if cnt["block_sweep"] else
                                                    "flops the sequential-chunk algorithm executes: m dim^2 + 72 dim^2 + sum over chunks "
                                                    "(r^3 / 3 + r^2 dim)"),
                                        flops_frac=cnt["upd_flops"] * K * R / (dev_ms / 1e3) / 1e12 / f64_tflops,
                                        hbm_frac=cnt["upd_bytes"] * K * R / (dev_ms / 1e3) / 1e9 / hbm_peak))
        value = world * R * K / (dev_ms_max * 1e-3)
        # the all-gather closes a lap of LAP_UPDATES updates: K of them were timed, so K / LAP_UPDATES of it is charged
        coll_share = coll_ms_max * min(1.0, K / float(LAP_UPDATES))
        e2e_ms = cpp_ms_max + coll_share
        line = dict(metric="vision-updates/sec", value=value, unit="updates/s", n_gpus=world, steps=K, warmup=W,
                    ms_per_step=dev_ms_max / K, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=workload_name(N, args.coord), landmarks=N, measured_per_update=m["n_meas"], state_dim=cnt["dim"],
                                sequences_per_gpu=R, l2="not flushed" if args.no_l2_flush else
                                "flushed between steps (256 MiB write) outside the per-step event brackets",
                                value_timing="CUDA events on the filter's stream around the device work of processVisionData "
                                "(frame upload -> status download; + augmentLandmarkStates kernels)" if R == 1 else
                                "host bracket of each step of the GPU's concurrent sequences (their device times overlap)",
                                launch_mode="per-kernel launches" if args.no_graph else "steady frames replayed as a cached CUDA graph",
                                parallelism=f"replicas: {R} sequence(s) per GPU x {world} GPU(s), no data-path collective, "
                                "one all-gather of trajectories per lap",
                                host_cores_rank0=host_cores),
                    e2e=dict(value=world * R * K / (e2e_ms * 1e-3), unit="updates/s", h2d_bytes_per_step=m["h2d"], d2h_bytes_per_step=m["d2h"],
                             ms_per_step=e2e_ms / K, collective_ms=coll_ms_max, collective_charged_ms=coll_share,
                             per_rank_ms_per_step=([dict(value=r[0] / K, python_driver=r[1] / K, e2e=r[2] / K) for r in per_rank]
                                                   if per_rank else None),
                             collective_note=(f"one all-gather of the trajectories per simulated lap ({LAP_UPDATES} updates); "
                                              f"{K} updates timed -> {K}/{LAP_UPDATES} of it is charged to e2e") if world > 1
                             else "single GPU: no collective",
                             driver=("C++ host loop over the C ABI (eqvio_replay), host wall clock per synchronised frame" if R == 1 else
                                     "one C++ host thread per sequence over the C ABI (eqvio_replay_batch), wall clock of the batch"),
                             real_data_flow=(dict(value=world * R * K3 / (real_ms_max * 1e-3), ms_per_step=real_ms_max / K3,
                                                  note="same C++ loop without augmentLandmarkStates: ids lost / added inside "
                                                  "processVisionData (eqvio_opt's flow)") if real_ms_max > 0 else None),
                             python_driver=dict(value=world * R * K / (py_ms_max * 1e-3), ms_per_step=py_ms_max / K,
                                                host_ms_per_step=1000.0 * m["wall_in"] / K,
                                                timing="CUDA events around each synchronised step")),
                    gpu_launches=int(m["launches"]), launches_per_step=m["launches"] / K, graphs=m.get("graphs"),
                    stage_ms={k_: v / K for k_, v in m["stage_acc"].items()},
                    stage_ms_note=("per-stage event brackets of plain launches" if args.no_graph else
                                   "the timed steps replay one CUDA graph per update: a single bracket, booked under correction; the "
                                   "propagation / preprocessing / correction split of the same update with plain launches is "
                                   "roofline.profile_stage_ms"),
                    clocks=sampler.result(), roofline=roofline)
        if batched:
            line["batched"] = batched
        if world == 1 and R == 1 and not args.no_sweep and N == 256:
            sweep = {}
            for Ns, Ks, Ps in ((64, min(K, 30), 6), (1024, min(K, 10), 3)):
                Ws = 6  # enough warm-up updates for the graphs of the recurring frame shapes to be captured before the timed steps
                ms_ = measure_single(eb, torch, args, Ns, Ks, Ws, Ps, local_rank, flush_buf, parity=True)
                cs = alg_counts(ms_["n_state"], ms_["n_meas"])
                ks = kernel_rooflines(ms_["prof"], ms_["nprof"], cs, Ns, hbm_peak, f64_tflops, traffic)
                keep = ("ms_per_update", "avg_launch_us", "bound", "achieved", "peak", "unit", "frac", "traffic")
                entry_ = dict(value=Ks / (ms_["dev_ms"] * 1e-3), e2e=Ks / (ms_["cpp_ms"] * 1e-3), unit="updates/s", steps=Ks, warmup=Ws,
                              ms_per_step=ms_["dev_ms"] / Ks, state_dim=cs["dim"], graphs=ms_["graphs"],
                              stage_ms={k_: v / Ks for k_, v in ms_["stage_acc"].items()},
                              real_data_flow=(ms_["K3"] / (ms_["real_ms"] * 1e-3) if ms_["real_ms"] > 0 else None),
                              update_flops_frac=cs["upd_flops"] * Ks / (ms_["dev_ms"] / 1e3) / 1e12 / f64_tflops,
                              kernels={k_: {kk_: v for kk_, v in e.items() if kk_ in keep} for k_, e in ks.items()},
                              parity_vs_oracle=ms_["parity"])
                if not args.no_cpu_baseline:
                    c_ = cpu_arm(ms_["stream"], settings_dict(args.coord), 2, 3 if Ns >= 1024 else 10, variants=Ns < 1024)
                    entry_["cpu_baseline"] = {k_: c_[k_] for k_ in c_ if k_ != "sample"}
                ms_["seq"].flt.close()
                sweep[str(Ns)] = entry_
            line["sweep"] = sweep
        if not args.no_cpu_baseline and world == 1 and R == 1:  # the CPU baseline is a single-GPU-run figure (rank 0 at N = 1 only)
            line["cpu_baseline"] = cpu_arm(m["stream"], settings_dict(args.coord), min(W, 2), args.cpu_sample or cpu_sample_size(N, K))
        guard.emit(line)
    m["seq"].flt.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global RICCATI
    args = parse()
    RICCATI = args.riccati
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
