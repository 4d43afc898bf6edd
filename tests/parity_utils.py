"""Shared helpers of the parity tests, smoke() and bench.py: record a simulated input stream with
the oracle's VIOSimulator restatement, replay it through the CPU oracle and through the CUDA path
(C ABI), and compare the results.

A *stream* is a dict: settings (oracle Settings), cam (oracle camera), init (oracle VIOState, the
truncated true state at t=0), frames (list of oracle.simulator.Frame).
"""
import numpy as np

from oracle import eqf
from oracle.simulator import SimulationDataServer, benchmarkSettings, benchmarkSim, replayOracle


def make_stream(N=32, frames=4, coord=0, seed=0, settings_overrides=None, sim_overrides=None):
    st = benchmarkSettings(coord, **(settings_overrides or {}))
    sim = benchmarkSim(N, seed, **(sim_overrides or {}))
    server = SimulationDataServer(sim, st)
    init = server.initialCondition()
    cam = server.simulator.cameraPtr
    st.cameraOffset = server.cameraExtrinsics()
    fr = server.record(frames)
    return dict(settings=st, cam=cam, init=init, frames=fr, N=N)


def snapshot_oracle(flt):
    fs = flt.viewEqFState()
    est = flt.stateEstimate()
    return dict(ids=np.array(fs.X.ids, dtype=np.int64), sensor=est.sensor.flat(), p=est.p.copy(), Sigma=fs.Sigma.copy(),
                X_sensor=fs.X.sensorFlat(), Qq=fs.X.Qq.copy(), Qa=fs.X.Qa.copy(), time=flt.getTime())


def run_oracle(stream, dense_lazy=False, on_update=None, structured=False):
    flt = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
    flt.filterState.mirrorLazyEvaluation = dense_lazy
    flt.filterState.structuredEvaluation = structured
    out = []

    def cb(k, f):
        out.append(snapshot_oracle(f))
        if on_update:
            on_update(k, f)

    replayOracle(flt, stream["frames"], stream["cam"], cb)
    return out


def gpu_filter(stream, capacity=None, device=0):
    import eqvio_b200 as eb

    st = eb.Settings.fromObject(stream["settings"])
    init = stream["init"]
    xi0 = eb.VIOState(eb.VIOSensorState.fromFlat(init.sensor.flat()), init.p, init.ids)
    cap = capacity if capacity is not None else max(stream["N"], init.p.shape[0]) + 8
    flt = eb.VIOFilter(st, xi0, 0.0, capacity=cap, device=device)
    cam = eb.Camera.fromPod(stream["cam"].pod())
    return flt, cam


def snapshot_gpu(flt):
    est = flt.stateEstimate()
    fs = flt.viewEqFState()
    return dict(ids=est.ids.copy(), sensor=est.sensor.flat(), p=est.p.copy(), Sigma=fs.Sigma, X_sensor=fs.X_sensor,
                Qq=fs.X_Qq, Qa=fs.X_Qa, time=flt.getTime())


def replay_gpu(flt, cam, frames, on_update=None):
    import eqvio_b200 as eb

    for k, fr in enumerate(frames):
        flt.processIMUArray(fr.imu)
        flt.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        if on_update:
            on_update(k, flt)


def run_gpu(stream, capacity=None, tuning=None):
    flt, cam = gpu_filter(stream, capacity)
    if tuning:
        flt.setTuning(**tuning)
    out = []
    replay_gpu(flt, cam, stream["frames"], lambda k, f: out.append(snapshot_gpu(f)))
    flt.close()
    return out


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def compare_states(g, r):
    """Relative Frobenius errors of the CUDA result g against the oracle result r (BASELINE.json's
    parity metric) and the element-wise landmark-id check."""
    ids_equal = g["ids"].shape == r["ids"].shape and bool(np.all(g["ids"] == r["ids"]))
    if not ids_equal:
        return dict(ids_equal=False, sigma=np.inf, state=np.inf, pose=np.inf)
    sg = np.concatenate([g["sensor"], g["p"].reshape(-1)])
    sr = np.concatenate([r["sensor"], r["p"].reshape(-1)])
    return dict(ids_equal=True, sigma=rel_fro(g["Sigma"], r["Sigma"]), state=rel_fro(sg, sr),
                pose=rel_fro(g["sensor"][6:13], r["sensor"][6:13]), time=abs(g["time"] - r["time"]))
