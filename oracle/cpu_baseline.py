"""TEST / BASELINE INFRASTRUCTURE (not product code): driver of oracle/cpu_update.c.

The numpy oracle replays a stream and, per vision update, records the operands of the two dense stages -- the propagation
(Sigma, A, B, dt; VIO_eqf.cpp:62-72) and the correction (Sigma, C, ytilde; VIO_eqf.cpp:105-135) -- OUTSIDE any timed region.
The C library then runs exactly those stages, in the reference's dense evaluation order or in the block-structured + Cholesky
form, timing itself with clock_gettime; every C result is compared with the oracle's own Sigma / Gamma for that update, so
the baseline that is timed is also checked.  What is not restated in C (and not timed) is the O(N) glue of the reference:
Lie-group actions, the per-landmark Jacobian blocks, landmark bookkeeping ("preprocessing" in the reference's LoopTimer).

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__ (build) may import this module.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

from . import eqf

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(BUILD_DIR, "libcpu_update.so")
SRC = os.path.join(HERE, "cpu_update.c")
_lib = None


def build(force=False):
    """gcc -O3 -march=native -> oracle/_build/libcpu_update.so (git-ignored, travels to the GPU box with the snapshot)."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC):
        # -march=x86-64-v3 rather than native: the library is built in this container and runs on the GPU box's host CPU;
        # the GEMMs (where the time goes) are OpenBLAS's own run-time dispatched kernels either way
        subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-shared", "-fPIC", "-o", LIB_PATH, SRC, "-ldl"], check=True)
    return LIB_PATH


def blas_path():
    import numpy

    hits = glob.glob(os.path.join(os.path.dirname(numpy.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
    if not hits:
        raise RuntimeError("numpy's bundled ILP64 OpenBLAS (numpy.libs/libscipy_openblas64_*.so) not found")
    return os.path.abspath(hits[0])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.cpu_init.restype = C.c_int
        L.cpu_init.argtypes = [C.c_char_p]
        L.cpu_set_threads.argtypes = [C.c_int]
        L.cpu_get_threads.restype = C.c_int
        L.cpu_dense_propagate.restype = C.c_double
        L.cpu_dense_propagate.argtypes = [C.c_int, pd, pd, pd, pd, pd, C.c_double]
        L.cpu_dense_correct.restype = C.c_double
        L.cpu_dense_correct.argtypes = [C.c_int, C.c_int, pd, pd, C.c_double, pd, pd]
        L.cpu_structured_propagate.restype = C.c_double
        L.cpu_structured_propagate.argtypes = [C.c_int, C.c_int, pd, pd, pd, pd, pd, pd, C.c_double]
        L.cpu_structured_correct.restype = C.c_double
        L.cpu_structured_correct.argtypes = [C.c_int, C.c_int, pd, pd, pi, C.c_double, pd, pd]
        rc = L.cpu_init(blas_path().encode())
        if rc != 0:
            raise RuntimeError(f"cpu_init failed ({rc})")
        _lib = L
    return _lib


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Recorder:
    """Patches VIO_eqf for the lifetime of a `with` block: records the operands and results of the fast Riccati step and of the
    vision update of every processVisionData call."""

    def __init__(self):
        self.updates = []
        self._cur = None

    def __enter__(self):
        rec = self
        self._riccati, self._vision = eqf.VIO_eqf.integrateRiccatiStateFast, eqf.VIO_eqf.performVisionUpdate

        def riccati(fs, imuVelocity, dt, inputGainMatrix, stateGainMatrix):
            A = fs.coordinateSuite.stateMatrixA(fs.X, fs.xi0, imuVelocity)
            B = fs.coordinateSuite.inputMatrixB(fs.X, fs.xi0)
            entry = dict(S0=fs.Sigma.copy(), A=np.ascontiguousarray(A), B=np.ascontiguousarray(B), dt=float(dt),
                         q=np.ascontiguousarray(np.diagonal(inputGainMatrix)), p=np.ascontiguousarray(np.diagonal(stateGainMatrix)))
            rec._riccati(fs, imuVelocity, dt, inputGainMatrix, stateGainMatrix)
            entry["S1"] = fs.Sigma.copy()
            rec._cur = dict(prop=entry)

        def vision(fs, measurement, outputGainMatrix, useEquivariantOutput=True, discreteCorrection=False):
            if measurement.camCoordinates and rec._cur is not None:
                est = eqf.measureSystemState(fs.stateEstimate(), measurement.cameraPtr)
                y = (measurement - est).asVector()
                Cm = fs.coordinateSuite.outputMatrixC(fs.xi0, fs.X, measurement, useEquivariantOutput)
                entry = dict(S0=fs.Sigma.copy(), C=np.ascontiguousarray(Cm), y=np.ascontiguousarray(y), r2=float(outputGainMatrix[0, 0]))
                rec._vision(fs, measurement, outputGainMatrix, useEquivariantOutput, discreteCorrection)
                entry["S1"] = fs.Sigma.copy()
                entry["Gamma"] = np.array(fs.lastGamma, dtype=np.float64)
                rec._cur["corr"] = entry
                rec.updates.append(rec._cur)
                rec._cur = None
            else:
                rec._vision(fs, measurement, outputGainMatrix, useEquivariantOutput, discreteCorrection)

        eqf.VIO_eqf.integrateRiccatiStateFast, eqf.VIO_eqf.performVisionUpdate = riccati, vision
        return self

    def __exit__(self, *exc):
        eqf.VIO_eqf.integrateRiccatiStateFast, eqf.VIO_eqf.performVisionUpdate = self._riccati, self._vision


def record_updates(flt, frames, cam, augment=True):
    """Replays frames through the oracle filter flt (already warmed up by the caller) and returns the recorded updates."""
    with Recorder() as rec:
        for fr in frames:
            for row in fr.imu:
                flt.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
            meas = eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, cam)
            if augment:
                flt.augmentLandmarkStates(meas.getIds(), eqf.VIOState(None, fr.provided_p, fr.ids))
            flt.processVisionData(meas)
    return rec.updates


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_updates(updates, structured=False, threads=None):
    """Runs the recorded updates through the C library.  Returns dict(updates_per_s, stage_ms (mean per update), worst relative
    error of Sigma / Gamma against the oracle, threads)."""
    L = lib()
    if threads:
        L.cpu_set_threads(int(threads))
    t_prop = t_corr = 0.0
    worst = 0.0
    for u in updates:
        pr, co = u["prop"], u["corr"]
        dim = pr["S0"].shape[0]
        S = pr["S0"].copy()
        if structured:
            N = (dim - 21) // 3
            As = np.ascontiguousarray(pr["A"][:, :21])
            ar = np.arange(N)
            D = np.ascontiguousarray(pr["A"][21:, 21:].reshape(N, 3, N, 3)[ar, :, ar, :]) if N else np.zeros((0, 3, 3))
            t = L.cpu_structured_propagate(dim, N, _pd(S), _pd(As), _pd(D), _pd(pr["B"]), _pd(pr["q"]), _pd(pr["p"]), pr["dt"])
        else:
            t = L.cpu_dense_propagate(dim, _pd(S), _pd(pr["A"]), _pd(pr["B"]), _pd(pr["q"]), _pd(pr["p"]), pr["dt"])
        t_prop += t
        worst = max(worst, _rel(S, pr["S1"]))
        dim = co["S0"].shape[0]
        m = co["C"].shape[0]
        S = co["S0"].copy()
        G = np.zeros(dim)
        if structured:
            N, n = (dim - 21) // 3, m // 2
            C4 = co["C"][:, 21:].reshape(n, 2, N, 3)
            lm = np.abs(C4).sum(axis=(1, 3)).argmax(axis=1).astype(np.int32)  # one 2x3 block per row pair (EqFMatrices.cpp:74-76)
            Cb = np.ascontiguousarray(C4[np.arange(n), :, lm, :])
            t = L.cpu_structured_correct(dim, n, _pd(S), _pd(Cb), lm.ctypes.data_as(C.POINTER(C.c_int)), co["r2"], _pd(co["y"]), _pd(G))
        else:
            t = L.cpu_dense_correct(dim, m, _pd(S), _pd(co["C"]), co["r2"], _pd(co["y"]), _pd(G))
        if t < 0:
            raise RuntimeError("cpu_update: factorisation of S failed")
        t_corr += t
        worst = max(worst, _rel(S, co["S1"]), _rel(G, co["Gamma"]))
    n = max(len(updates), 1)
    return dict(updates_per_s=n / (t_prop + t_corr), stage_ms=dict(propagation=1e3 * t_prop / n, preprocessing=0.0, correction=1e3 * t_corr / n),
                worst_rel_error_vs_oracle=worst, threads=L.cpu_get_threads(), updates=len(updates))
