"""Host-side mirror of the reference's ``VIOFilter`` (include/eqvio/VIOFilter.h:36-192) over the
C ABI.  Same member names, argument meaning and silent-return behaviour as the reference class;
all arithmetic happens on the GPU behind ``libeqvio_b200.so``.

Containers are plain numpy-backed records (the reference's are Eigen/LiePP types):

  VIOSensorState  inputBias(6) | pose (q wxyz, x) | velocity(3) | cameraOffset (q wxyz, x)
  VIOState        sensor + landmarks ``p`` (N,3) and ``ids`` (N,) in state order
  IMUVelocity     stamp, gyr, acc, gyrBiasVel, accBiasVel           (IMUVelocity.h:33-84)
  VisionMeasurement  stamp, {id: pixel}, camera                    (VisionMeasurement.h:35-40)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import EqvioError, lib

COORD_EUCLIDEAN, COORD_INVDEPTH, COORD_NORMAL = 0, 1, 2
CAMERA_PINHOLE, CAMERA_RADTAN, CAMERA_EQUIDISTANT = 0, 1, 2


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if n is not None and a.shape[0] != n:
        raise ValueError(f"expected {n} doubles, got {a.shape[0]}")
    return a


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32).reshape(-1)


class ReplayFrame(C.Structure):  # eqvio_replay_frame
    _fields_ = [("stamp", C.c_double), ("n_imu", C.c_int), ("imu_rows", C.c_void_p), ("n", C.c_int), ("ids", C.c_void_p),
                ("y", C.c_void_p), ("provided_p", C.c_void_p)]


@dataclass
class IMUVelocity:
    stamp: float = 0.0
    gyr: np.ndarray = field(default_factory=lambda: np.zeros(3))
    acc: np.ndarray = field(default_factory=lambda: np.zeros(3))
    gyrBiasVel: np.ndarray = field(default_factory=lambda: np.zeros(3))
    accBiasVel: np.ndarray = field(default_factory=lambda: np.zeros(3))


@dataclass
class VIOSensorState:
    inputBias: np.ndarray = field(default_factory=lambda: np.zeros(6))
    pose_q: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0]))
    pose_x: np.ndarray = field(default_factory=lambda: np.zeros(3))
    velocity: np.ndarray = field(default_factory=lambda: np.zeros(3))
    cameraOffset_q: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0]))
    cameraOffset_x: np.ndarray = field(default_factory=lambda: np.zeros(3))

    def flat(self):
        return np.concatenate([self.inputBias, self.pose_q, self.pose_x, self.velocity, self.cameraOffset_q,
                               self.cameraOffset_x]).astype(np.float64)

    @staticmethod
    def fromFlat(f):
        f = np.asarray(f, dtype=np.float64)
        return VIOSensorState(f[0:6].copy(), f[6:10].copy(), f[10:13].copy(), f[13:16].copy(), f[16:20].copy(),
                              f[20:23].copy())


@dataclass
class VIOState:
    sensor: VIOSensorState = field(default_factory=VIOSensorState)
    p: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    ids: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))

    def getIds(self):
        return [int(i) for i in self.ids]

    def Dim(self):
        return 21 + 3 * len(self.ids)


@dataclass
class EqFState:
    """What ``viewEqFState()`` exposes of VIO_eqf (VIO_eqf.h:34-42): xi0, X and Sigma."""

    xi0: VIOState
    X_sensor: np.ndarray  # 23: beta6 | A (q, x) | w3 | B (q, x)
    X_Qq: np.ndarray  # (N,4) wxyz
    X_Qa: np.ndarray  # (N,)
    Sigma: np.ndarray  # (dim, dim)
    currentTime: float


class Camera:
    """Flattened GIFT::GICamera (pinhole / radtan).  ``invDist`` of a radtan camera is recomputed by
    the library (StandardCamera::computeInverseDistortion) unless given."""

    def __init__(self, width, height, fx, fy, cx, cy, dist=(), inv_dist=None, model=None):
        pod = _capi.Camera()
        pod.model = model if model is not None else (CAMERA_RADTAN if len(dist) else CAMERA_PINHOLE)
        pod.width, pod.height, pod.ndist = int(width), int(height), len(dist)
        pod.fx, pod.fy, pod.cx, pod.cy = float(fx), float(fy), float(cx), float(cy)
        for i in range(5):
            pod.dist[i] = float(dist[i]) if i < len(dist) else 0.0
            pod.inv_dist[i] = 0.0
        if len(dist) and pod.model == CAMERA_RADTAN:
            if inv_dist is None:
                rc = lib.eqvio_camera_fit_inverse_distortion(C.byref(pod))
                if rc != 0:
                    raise EqvioError(rc, "inverse distortion fit failed")
            else:
                for i in range(min(5, len(inv_dist))):  # invDist always has five entries (StandardCamera.cpp:117-147)
                    pod.inv_dist[i] = float(inv_dist[i])
        self.pod = pod

    @staticmethod
    def fromPod(d):
        """From the dict produced by the oracle cameras' ``pod()`` (tests) or any mapping with the same keys."""
        nd = d["ndist"]
        return Camera(d["width"], d["height"], d["fx"], d["fy"], d["cx"], d["cy"], d["dist"][:nd],
                      d["inv_dist"][:5] if (nd and d["model"] == CAMERA_RADTAN) else None, model=d["model"])


@dataclass
class VisionMeasurement:
    stamp: float = 0.0
    camCoordinates: dict = field(default_factory=dict)
    cameraPtr: Camera = None

    @staticmethod
    def fromArrays(stamp, ids, y, cameraPtr):
        y = np.asarray(y, dtype=np.float64).reshape(-1, 2)
        return VisionMeasurement(float(stamp), {int(i): y[k] for k, i in enumerate(ids)}, cameraPtr)

    def getIds(self):
        return sorted(self.camCoordinates.keys())

    def arrays(self):
        ids = np.array(self.getIds(), dtype=np.int32)
        y = np.array([self.camCoordinates[int(i)] for i in ids], dtype=np.float64).reshape(-1, 2)
        return ids, y


class Settings:
    """VIOFilter::Settings; attribute names are the reference's (VIOFilterSettings.h:58-99)."""

    def __init__(self, **overrides):
        pod = _capi.Settings()
        lib.eqvio_settings_default(C.byref(pod))
        object.__setattr__(self, "pod", pod)
        for k, v in overrides.items():
            setattr(self, k, v)

    _names = {n for n, _ in _capi.Settings._fields_}

    def __getattr__(self, name):
        if name in Settings._names and name != "cameraOffset":
            return getattr(self.pod, name)
        if name == "cameraOffset":
            return np.array(list(self.pod.cameraOffset))
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name == "cameraOffset":  # (q wxyz, x)
            v = _f64(value, 7)
            for i in range(7):
                self.pod.cameraOffset[i] = v[i]
        elif name in Settings._names:
            setattr(self.pod, name, int(value) if isinstance(getattr(self.pod, name), int) else float(value))
        else:
            raise AttributeError(f"VIOFilter::Settings has no field {name}")

    @staticmethod
    def fromObject(o):
        """Copy the same-named attributes of any settings-like object (e.g. the oracle's)."""
        s = Settings()
        for n in Settings._names:
            if n == "cameraOffset":
                co = getattr(o, n, None)
                if co is not None:
                    s.cameraOffset = np.concatenate([np.asarray(co.q, dtype=np.float64), np.asarray(co.x, dtype=np.float64)]) \
                        if hasattr(co, "q") else co
            elif hasattr(o, n):
                setattr(s, n, getattr(o, n))
        return s


class VIOFilter:
    """VIOFilter (src/VIOFilter.cpp) with the filter state resident on one GPU.

    ``capacity`` is the largest landmark count the filter will hold (Sigma is allocated once,
    capacity-padded, in HBM)."""

    def __init__(self, settings: Settings, xi0: VIOState = None, time: float = 0.0, *, capacity: int = 256,
                 device: int = 0, stream=None):
        self.settings = settings
        self._h = _capi._H()
        st = C.c_void_p(stream) if stream else None
        if xi0 is None:  # VIOFilter(const Settings&), VIOFilter.cpp:31-41
            rc = lib.eqvio_create(C.byref(settings.pod), device, capacity, st, C.byref(self._h))
        else:  # VIOFilter(const VIOState&, const Settings&, const double&), VIOFilter.cpp:43-56
            sensor = _f64(xi0.sensor.flat(), 23)
            ids = _i32(xi0.ids)
            p = _f64(xi0.p)
            rc = lib.eqvio_create_from_state(C.byref(settings.pod), device, capacity, st, _pd(sensor), len(ids), _pi(ids),
                                             _pd(p), float(time), C.byref(self._h))
        if rc != 0:
            self._h = None
            raise EqvioError(rc, (lib.eqvio_last_error(None) or b"").decode())
        self.capacity = capacity

    def close(self):
        if getattr(self, "_h", None):
            lib.eqvio_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EqvioError(rc, (lib.eqvio_last_error(self._h) or b"").decode())

    # -- input ---------------------------------------------------------------------------------------
    def processIMUData(self, imu: IMUVelocity):  # VIOFilter.cpp:58-63
        self._check(lib.eqvio_process_imu(self._h, float(imu.stamp), _pd(_f64(imu.gyr, 3)), _pd(_f64(imu.acc, 3)),
                                          _pd(_f64(imu.gyrBiasVel, 3)), _pd(_f64(imu.accBiasVel, 3))))

    def processIMUArray(self, rows):
        """rows (k,13): stamp, gyr3, acc3, gyrBiasVel3, accBiasVel3 -- k processIMUData calls."""
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 13)
        self._check(lib.eqvio_process_imu_rows(self._h, rows.shape[0], _pd(rows)))

    def initialiseFromIMUData(self, imu: IMUVelocity):  # VIOFilter.cpp:65-78
        self._check(lib.eqvio_initialise_from_imu(self._h, float(imu.stamp), _pd(_f64(imu.gyr, 3)), _pd(_f64(imu.acc, 3))))

    def setState(self, xi: VIOState):  # VIOFilter.cpp:80-92
        ids = _i32(xi.ids)
        self._check(lib.eqvio_set_state(self._h, _pd(_f64(xi.sensor.flat(), 23)), len(ids), _pi(ids), _pd(_f64(xi.p))))

    def setLandmarks(self, p, ids):  # VIOFilter.cpp:94-110
        ids = _i32(ids)
        self._check(lib.eqvio_set_landmarks(self._h, len(ids), _pi(ids), _pd(_f64(p))))

    def augmentLandmarkStates(self, newIds, providedState: VIOState):  # VIOFilter.cpp:112-132
        nid = _i32(newIds)
        pid = _i32(providedState.ids)
        self._check(lib.eqvio_augment_landmark_states(self._h, len(nid), _pi(nid), len(pid), _pi(pid),
                                                      _pd(_f64(providedState.p))))

    def processVisionData(self, measurement: VisionMeasurement):  # VIOFilter.cpp:194-241
        ids, y = measurement.arrays()
        return self.processVisionArrays(measurement.stamp, ids, y, measurement.cameraPtr)

    def processVisionArrays(self, stamp, ids, y, camera: Camera):
        """processVisionData on flat arrays (ids ascending, y (n,2) pixels).  Returns True when the
        EqF correction ran, False on the reference's silent returns."""
        ids = _i32(ids)
        y = _f64(y, 2 * len(ids))
        did = C.c_int(0)
        self._check(lib.eqvio_process_vision(self._h, float(stamp), len(ids), _pi(ids), _pd(y), C.byref(camera.pod),
                                             C.byref(did)))
        return bool(did.value)

    # -- output --------------------------------------------------------------------------------------
    def getTime(self):  # VIOFilter.cpp:256
        return lib.eqvio_get_time(self._h)

    def isInitialised(self):
        return bool(lib.eqvio_is_initialised(self._h))

    def numLandmarks(self):
        return lib.eqvio_num_landmarks(self._h)

    def stateEstimate(self) -> VIOState:  # VIOFilter.cpp:243
        N = self.numLandmarks()
        sensor = np.zeros(23)
        ids = np.zeros(max(N, 1), dtype=np.int32)
        p = np.zeros(3 * max(N, 1))
        n = C.c_int(0)
        self._check(lib.eqvio_get_state_estimate(self._h, _pd(sensor), _pi(ids), _pd(p), C.byref(n)))
        return VIOState(VIOSensorState.fromFlat(sensor), p[:3 * N].reshape(N, 3).copy(), ids[:N].astype(np.int64))

    def viewEqFState(self, withSigma=True) -> EqFState:  # VIOFilter.cpp:245
        N = self.numLandmarks()
        dim = 21 + 3 * N
        xs = np.zeros(23)
        ids = np.zeros(max(N, 1), dtype=np.int32)
        p = np.zeros(3 * max(N, 1))
        X = np.zeros(23)
        XQ = np.zeros(5 * max(N, 1))
        Sig = np.zeros((dim, dim)) if withSigma else None
        self._check(lib.eqvio_get_eqf_state(self._h, _pd(xs), _pi(ids), _pd(p), _pd(X), _pd(XQ),
                                            _pd(Sig) if withSigma else None, dim))
        XQ = XQ[:5 * N].reshape(N, 5)
        xi0 = VIOState(VIOSensorState.fromFlat(xs), p[:3 * N].reshape(N, 3).copy(), ids[:N].astype(np.int64))
        # Sigma arrives column-major; it is symmetric to rounding, transposing restores (row, col) indexing
        return EqFState(xi0, X, XQ[:, 0:4].copy(), XQ[:, 4].copy(), Sig.T.copy() if withSigma else None, self.getTime())

    def landmarkCovBlocks(self):
        """VIO_eqf::getLandmarkCovById for every landmark: (N,3,3)."""
        N = self.numLandmarks()
        b = np.zeros(9 * max(N, 1))
        self._check(lib.eqvio_get_landmark_cov_blocks(self._h, _pd(b)))
        return b[:9 * N].reshape(N, 3, 3).transpose(0, 2, 1).copy()

    def getFeaturePredictions(self, camera: Camera, stamp: float = -1.0) -> VisionMeasurement:  # VIOFilter.cpp:247-252
        N = self.numLandmarks()
        ids = np.zeros(max(N, 1), dtype=np.int32)
        y = np.zeros(2 * max(N, 1))
        n = C.c_int(0)
        self._check(lib.eqvio_get_feature_predictions(self._h, C.byref(camera.pod), float(stamp), _pi(ids), _pd(y), C.byref(n)))
        k = n.value
        return VisionMeasurement(float(stamp), {int(ids[j]): y[2 * j:2 * j + 2].copy() for j in range(k)}, camera)

    def computeNEES(self, trueState: VIOState) -> float:
        """viewEqFState().computeNEES(trueState) of the reference (VIO_eqf.cpp:153-170), evaluated on the device."""
        ids = _i32(trueState.ids)
        out = C.c_double(0.0)
        self._check(lib.eqvio_compute_nees(self._h, _pd(_f64(trueState.sensor.flat(), 23)), len(ids), _pi(ids), _pd(_f64(trueState.p)),
                                           C.byref(out)))
        return out.value

    def lastOutliers(self):
        ids = np.zeros(max(self.capacity, 1), dtype=np.int32)
        n = C.c_int(0)
        self._check(lib.eqvio_get_last_outliers(self._h, _pi(ids), len(ids), C.byref(n)))
        return [int(i) for i in ids[:n.value]]

    # -- measurement hooks -----------------------------------------------------------------------------
    def enableStageTiming(self, on=True):
        self._check(lib.eqvio_enable_stage_timing(self._h, int(on)))

    def stageMs(self):
        ms = np.zeros(3)
        self._check(lib.eqvio_get_stage_ms(self._h, _pd(ms)))
        return dict(propagation=ms[0], preprocessing=ms[1], correction=ms[2])

    def setTuning(self, **knobs):
        """Evaluation-order knobs (eqvio_set_tuning, include/eqvio_b200.h EQVIO_TUNE_*): correction (0 = sequential chunks, 1 = batch
        sweep), chunkLandmarks, speculate, graph, downdate (0 = fp64 DMMA, 1 = tcgen05 split-bf16), lookahead (split
        downdates beside the next factor kernel), fuseObserver, pdl, fuseSmall, speculateNew, stageS (S blocks through one TMA
        tensor copy), lazyDowndate (M = deferred Sigma tiles visited once per M chunks in the look-ahead form, 0 = every chunk).  Results agree across every knob (tests/test_gpu_parity.py)."""
        for name, value in knobs.items():
            if value is None:
                continue
            if name not in _capi.TUNE:
                raise TypeError(f"setTuning: unknown knob {name!r} (known: {', '.join(_capi.TUNE)})")
            self._check(lib.eqvio_set_tuning(self._h, _capi.TUNE[name], int(value)))

    def replay(self, frames, camera: Camera, flushBytes=0):
        """C++ host loop over the C ABI (eqvio_replay): per frame processIMUData x k, augmentLandmarkStates,
        processVisionData, stateEstimate on host buffers.  frames: objects with stamp, imu (k,13), ids, y (n,2),
        provided_p (n,3) or None.  Returns (frame_ms, est_sensor (len, 23))."""
        n = len(frames)
        arr = (ReplayFrame * n)()
        keep = []  # the C side reads these buffers during the call
        for k, fr in enumerate(frames):
            imu = np.ascontiguousarray(fr.imu, dtype=np.float64).reshape(-1, 13)
            ids = _i32(fr.ids)
            y = _f64(fr.y, 2 * len(ids))
            pp = None if getattr(fr, "provided_p", None) is None else _f64(fr.provided_p, 3 * len(ids))
            keep += [imu, ids, y, pp]
            arr[k].stamp = float(fr.stamp)
            arr[k].n_imu = imu.shape[0]
            arr[k].imu_rows = imu.ctypes.data
            arr[k].n = len(ids)
            arr[k].ids = ids.ctypes.data
            arr[k].y = y.ctypes.data
            arr[k].provided_p = None if pp is None else pp.ctypes.data
        ms = np.zeros(max(n, 1))
        est = np.zeros((max(n, 1), 23))
        self._check(lib.eqvio_replay(self._h, n, C.cast(arr, C.c_void_p), C.cast(C.pointer(camera.pod), C.c_void_p), int(flushBytes),
                                     _pd(ms), _pd(est)))
        return ms[:n], est[:n]

    @staticmethod
    def _replay_frames(frames, keep):
        n = len(frames)
        arr = (ReplayFrame * n)()
        for k, fr in enumerate(frames):
            imu = np.ascontiguousarray(fr.imu, dtype=np.float64).reshape(-1, 13)
            ids = _i32(fr.ids)
            y = _f64(fr.y, 2 * len(ids))
            pp = None if getattr(fr, "provided_p", None) is None else _f64(fr.provided_p, 3 * len(ids))
            keep += [imu, ids, y, pp]
            arr[k].stamp = float(fr.stamp)
            arr[k].n_imu = imu.shape[0]
            arr[k].imu_rows = imu.ctypes.data
            arr[k].n = len(ids)
            arr[k].ids = ids.ctypes.data
            arr[k].y = y.ctypes.data
            arr[k].provided_p = None if pp is None else pp.ctypes.data
        return arr

    def hostProfile(self, reset=True):
        """Host-side microseconds per processVisionData call since the last reset (eqvio_get_host_profile)."""
        us = np.zeros(4)
        calls = C.c_longlong(0)
        self._check(lib.eqvio_get_host_profile(self._h, int(reset), _pd(us), C.byref(calls)))
        n = max(int(calls.value), 1)
        return dict(enqueue_us=us[0] / n, decisions_us=us[1] / n, device_wait_us=us[2] / n, finish_us=us[3] / n, calls=int(calls.value))

    def launchCount(self):
        return int(lib.eqvio_get_launch_count(self._h))

    def graphStats(self):
        """(captures, replays) of the cached CUDA graphs of steady frames (eqvio_get_graph_stats)."""
        c, r = C.c_longlong(0), C.c_longlong(0)
        self._check(lib.eqvio_get_graph_stats(self._h, C.byref(c), C.byref(r)))
        return int(c.value), int(r.value)

    def enableKernelProfile(self, on=True):
        self._check(lib.eqvio_enable_kernel_profile(self._h, int(on)))

    def kernelProfile(self, reset=False):
        ms = np.zeros(_capi.PROF_CLASSES)
        ln = np.zeros(_capi.PROF_CLASSES, dtype=np.int64)
        self._check(lib.eqvio_get_kernel_profile(self._h, int(reset), _pd(ms), ln.ctypes.data_as(C.POINTER(C.c_longlong))))
        return {name: dict(ms=float(ms[i]), launches=int(ln[i])) for i, name in enumerate(_capi.PROF_NAMES)}


def replayBatch(filters, frames_per_filter, camera: Camera):
    """eqvio_replay_batch: one C++ host thread per filter, each replaying its own frames through the C ABI.
    Returns (frame_ms (R, K), est_sensor (R, K, 23), wall_ms)."""
    R = len(filters)
    K = len(frames_per_filter[0])
    keep = []
    arrs = [VIOFilter._replay_frames(fr, keep) for fr in frames_per_filter]
    ptrs = (C.c_void_p * R)(*[C.cast(a, C.c_void_p) for a in arrs])
    hs = (C.c_void_p * R)(*[C.cast(f._h, C.c_void_p) for f in filters])
    ms = np.zeros((R, max(K, 1)))
    est = np.zeros((R, max(K, 1), 23))
    wall = C.c_double(0.0)
    rc = lib.eqvio_replay_batch(C.cast(hs, C.c_void_p), R, K, C.cast(ptrs, C.c_void_p), C.cast(C.pointer(camera.pod), C.c_void_p), _pd(ms),
                                _pd(est), C.byref(wall))
    if rc != 0:
        for f in filters:
            f._check(rc)
    return ms[:, :K], est[:, :K], wall.value


def batchProcessVision(filters, stamps, ids_list, y_list, camera: Camera):
    """eqvio_batch_process_vision: the same update for independent filters (Monte-Carlo replicas on one
    GPU); kernels of different filters overlap on their streams.  Returns did_update per filter."""
    k = len(filters)
    H = (_capi._H * k)(*[f._h for f in filters])
    st = np.ascontiguousarray(stamps, dtype=np.float64)
    ids = [_i32(i) for i in ids_list]
    ys = [_f64(y) for y in y_list]
    n = np.array([len(i) for i in ids], dtype=np.int32)
    IP = (C.POINTER(C.c_int) * k)(*[_pi(i) for i in ids])
    YP = (C.POINTER(C.c_double) * k)(*[_pd(y) for y in ys])
    did = np.zeros(k, dtype=np.int32)
    rc = lib.eqvio_batch_process_vision(H, k, _pd(st), _pi(n), IP, YP, C.byref(camera.pod), _pi(did))
    if rc != 0:
        msgs = [(lib.eqvio_last_error(f._h) or b"").decode() for f in filters]
        raise EqvioError(rc, "; ".join(m for m in msgs if m))
    return did.astype(bool)
