"""The arithmetic shared by host and device code (eqvio_b200/csrc/{lie,model}.cuh and the sensor-sized
parts of kernels.cuh) compiled for the HOST through tests/csrc/host_hooks.cu and checked against the
oracle -- CPU-only coverage of the code the kernels execute (no GPU needed, no product path involved)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import eqf
from oracle.camera import EquidistantCamera, StandardCamera, createDefaultCamera
from parity_utils import make_stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "csrc", "host_hooks.cu")
LIB = os.path.join(ROOT, "tests", "lib", "libeqvio_host_hooks.so")
PD = C.POINTER(C.c_double)


def pd(a):
    return a.ctypes.data_as(PD)


@pytest.fixture(scope="module")
def hooks():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "eqvio_b200", "csrc", f) for f in ("kernels.cuh", "model.cuh", "lie.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler",
                        "-fPIC", "-shared", "-diag-suppress", "550", "-o", LIB, SRC], check=True)
    return C.CDLL(LIB)


@pytest.mark.parametrize("coord,discrete", [(0, 1), (1, 1), (0, 0)])
def test_riccati_context_and_observer_sensor_part(hooks, coord, discrete):
    stream = make_stream(N=6, frames=3, coord=coord, settings_overrides=dict(useDiscreteVelocityLift=bool(discrete)))
    st, init = stream["settings"], stream["init"]
    o = eqf.VIOFilter(st, init, 0.0)
    # make X non-trivial first: run one full update in the oracle
    fr = stream["frames"][1]
    for row in fr.imu:
        o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
    o.augmentLandmarkStates(list(fr.ids), eqf.VIOState(None, fr.provided_p, fr.ids))
    o.processVisionData(eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, stream["cam"]))
    fs = o.filterState
    imu = stream["frames"][2].imu
    t, newT = fs.currentTime, stream["frames"][2].stamp
    n = len(imu)
    dts = np.array([max((min(imu[i + 1, 0], newT) if i + 1 < n else newT) - max(imu[i, 0], t), 0.0) for i in range(n)])
    accT = dts.sum()
    mean = (imu[:, 1:] * dts[:, None]).sum(0) * (1.0 / accT)
    xi0s = fs.xi0.sensor.flat().copy()
    Xs = fs.X.sensorFlat().copy()
    rows = np.ascontiguousarray(np.concatenate([dts[:, None], imu[:, 1:]], 1))
    ctx = np.zeros(hooks.hook_sizeof_ctx() // 8)
    steps = np.zeros(n * hooks.hook_sizeof_step() // 8)
    qd = np.array([st.velGyrNoise**2, st.velAccNoise**2, st.velGyrBiasWalk**2, st.velAccBiasWalk**2])
    pdg = np.array([st.biasOmegaProcessVariance, st.biasAccelProcessVariance, st.attitudeProcessVariance,
                    st.positionProcessVariance, st.velocityProcessVariance, st.cameraAttitudeProcessVariance,
                    st.cameraPositionProcessVariance, st.pointProcessVariance])
    hooks.hook_sensor_prep(pd(xi0s), pd(Xs), pd(rows), n, pd(mean), C.c_double(accT), discrete, pd(qd), pd(pdg), pd(ctx), pd(steps))
    Fs, Ns = ctx[:441].reshape(21, 21), ctx[441:882].reshape(21, 21)
    u = eqf.IMUVelocity(0.0, mean[0:3], mean[3:6], mean[6:9], mean[9:12])
    A = fs.coordinateSuite.stateMatrixA(fs.X, fs.xi0, u)
    B = fs.coordinateSuite.inputMatrixB(fs.X, fs.xi0)
    Fo = np.eye(21) + accT * A[:21, :21]
    No = accT * (B[:21] @ st.constructInputGainMatrix() @ B[:21].T + st.constructStateGainMatrix(0))
    assert np.abs(Fs - Fo).max() < 1e-14 and np.abs(Ns - No).max() < 1e-16
    # observer integration, sensor part of X
    for i in range(n):
        fs.integrateObserverState(eqf.IMUVelocity(imu[i, 0], imu[i, 1:4], imu[i, 4:7], imu[i, 7:10], imu[i, 10:13]), dts[i], bool(discrete))
    assert np.abs(Xs - fs.X.sensorFlat()).max() < 1e-13


@pytest.mark.parametrize("coord", [0, 1])
@pytest.mark.parametrize("radtan", [0, 1, 2])
def test_output_block_matches_oracle(hooks, coord, radtan):
    rng = np.random.default_rng(5)
    cam = [createDefaultCamera(),
           StandardCamera(752, 480, 458.654, 457.296, 367.215, 248.375, [-0.283, 0.074, 0.0002, 1.8e-05, 0.0]),
           EquidistantCamera(512, 512, 190.978, 190.973, 254.93, 256.9, [-0.0137, 0.0207, -0.0128, 0.0025])][radtan]
    pod = cam.pod()
    camv = np.array([pod["model"], pod["width"], pod["height"], pod["ndist"], pod["fx"], pod["fy"], pod["cx"], pod["cy"]]
                    + list(pod["dist"]) + list(pod["inv_dist"]), dtype=np.float64)
    suite = eqf.getCoordinates(coord)
    for _ in range(10):
        q0 = rng.uniform(-1, 1, 3) * 2 + np.array([0, 0, 6.0])
        Qq = eqf.lg.so3_exp(rng.uniform(-1, 1, 3) * 0.05)
        Qa = 1.0 + 0.2 * rng.uniform()
        qhat = eqf.lg.sot3_apply_inverse(Qq, np.array(Qa), q0)
        y = cam.projectPoint(qhat) + rng.uniform(-2, 2, 2)
        Cs = np.zeros(6)
        hooks.hook_output_block(pd(camv), coord, pd(np.ascontiguousarray(q0)), pd(np.ascontiguousarray(Qq)), C.c_double(Qa), 1,
                                C.c_double(y[0]), C.c_double(y[1]), pd(Cs))
        ref = suite.outputMatrixCiStar(q0[None], Qq[None], np.array([Qa]), cam, y[None])[0]
        assert np.abs(Cs.reshape(2, 3) - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())
        hooks.hook_output_block(pd(camv), coord, pd(np.ascontiguousarray(q0)), pd(np.ascontiguousarray(Qq)), C.c_double(Qa), 0,
                                C.c_double(0), C.c_double(0), pd(Cs))
        ref0 = suite.outputMatrixCi(q0[None], Qq[None], np.array([Qa]), cam)[0]
        assert np.abs(Cs.reshape(2, 3) - ref0).max() < 1e-10 * max(1.0, np.abs(ref0).max())
