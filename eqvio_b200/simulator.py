"""Device VIOSimulator (include/eqvio_b200_sim.h): IMU and vision streams of the reference's simulator for many Monte-Carlo
instances at once.  The world points of each instance come from the host generator (``simdata``, seeded
numpy draws -- SURVEY 8c: the reference's libc ``rand()`` stream is not reproducible across platforms); trajectory, IMU,
visibility, selection, sorting and the true states run on the GPU, and so does the reference's input / output noise
(VIOSimulator.cpp:163-167, 258-262): counter-based Philox draws keyed by the instance's noise seed (simdata/philox.py states the
same function on the host), i.e. the noisy Monte-Carlo instances of BASELINE configs[4] come out of the same launches."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import lib

_H = C.c_void_p
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)
lib.eqvio_sim_create.restype = _H
lib.eqvio_sim_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _PD, C.c_int, C.c_int, _PD, _PI]
lib.eqvio_sim_destroy.restype = None
lib.eqvio_sim_destroy.argtypes = [_H]
lib.eqvio_sim_imu.restype = C.c_int
lib.eqvio_sim_imu.argtypes = [_H, C.c_int, _PD, _PD]
lib.eqvio_sim_vision.restype = C.c_int
lib.eqvio_sim_vision.argtypes = [_H, C.c_int, _PD, _PI, _PI, _PD, _PD, _PD, C.POINTER(C.c_float)]
lib.eqvio_sim_imu_instances.restype = C.c_int
lib.eqvio_sim_imu_instances.argtypes = [_H, C.c_int, _PD, _PD]
lib.eqvio_sim_set_noise.restype = C.c_int
lib.eqvio_sim_set_noise.argtypes = [_H, C.c_int, C.c_int, _PD, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_ulonglong)]
lib.eqvio_sim_last_error.restype = C.c_char_p


def _pd(a):
    return a.ctypes.data_as(_PD)


def _pi(a):
    return a.ctypes.data_as(_PI)


class DeviceSimulator:
    """``configs``: simdata.SimConfig objects that share everything but the seed (one per Monte-Carlo instance)."""

    def __init__(self, configs, device=0):
        from simdata.vio_simulator import _Simulator  # host side: world points only

        c0 = configs[0]
        for c in configs:
            if (c.inputNoise, c.outputNoise) != (c0.inputNoise, c0.outputNoise):
                raise ValueError("instances must share the noise switches")
            if (c.numPoints, c.maxFeatures, c.duration, c.numWalls, c.wallDistance) != (c0.numPoints, c0.maxFeatures, c0.duration, c0.numWalls,
                                                                                          c0.wallDistance):
                raise ValueError("instances must share the scene geometry (only the seed differs)")
        self.configs = list(configs)
        self.cfg = c0
        sims = [_Simulator(c) for c in configs]
        self._points = np.ascontiguousarray(np.stack([s.pointsP for s in sims]), dtype=np.float64)
        self._ids = np.ascontiguousarray(np.stack([s.pointsId for s in sims]), dtype=np.int32)
        intr = np.array([c0.fx, c0.fy, c0.cx, c0.cy], dtype=np.float64)
        self._h = lib.eqvio_sim_create(int(device), len(configs), int(c0.numPoints), int(c0.maxFeatures), float(c0.duration), _pd(intr),
                                       int(c0.width), int(c0.height), _pd(self._points), _pi(self._ids))
        if not self._h:
            raise RuntimeError((lib.eqvio_sim_last_error() or b"").decode())
        self.last_vision_ms = 0.0
        self.noise_seeds = np.array([(c.randomSeed if c.noiseSeed is None else c.noiseSeed) for c in configs], dtype=np.uint64)
        if c0.inputNoise or c0.outputNoise:
            # standard deviations of VIOSimulator.cpp:163-167 (constructInputGainMatrix() x sampling frequency) and :258-262
            self.imu_sigma = np.sqrt(np.repeat(np.array([c0.velGyrNoise, c0.velAccNoise, c0.velGyrBiasWalk, c0.velAccBiasWalk]) ** 2, 3)
                                     * max(c0.imuFreq, 0.0))
            rc = lib.eqvio_sim_set_noise(self._h, int(c0.inputNoise), int(c0.outputNoise), _pd(self.imu_sigma), float(c0.measurementNoise),
                                         float(c0.imuFreq), float(c0.imageFreq), self.noise_seeds.ctypes.data_as(C.POINTER(C.c_ulonglong)))
            if rc != 0:
                raise RuntimeError("eqvio_sim_set_noise failed")

    def close(self):
        if getattr(self, "_h", None):
            lib.eqvio_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def imu(self, stamps):
        """(k,) stamps -> (k, 13) rows: stamp, gyr3, acc3, gyrBiasVel3, accBiasVel3 (identical for every instance)."""
        t = np.ascontiguousarray(stamps, dtype=np.float64).reshape(-1)
        rows = np.zeros((t.shape[0], 13))
        if lib.eqvio_sim_imu(self._h, t.shape[0], _pd(t), _pd(rows)) != 0:
            raise RuntimeError((lib.eqvio_sim_last_error() or b"").decode())
        return rows

    def imu_instances(self, stamps):
        """(k,) stamps -> (instances, k, 13): the IMU rows of every instance (they differ when input noise is on)."""
        t = np.ascontiguousarray(stamps, dtype=np.float64).reshape(-1)
        rows = np.zeros((len(self.configs), t.shape[0], 13))
        if lib.eqvio_sim_imu_instances(self._h, t.shape[0], _pd(t), _pd(rows)) != 0:
            raise RuntimeError((lib.eqvio_sim_last_error() or b"").decode())
        return rows

    def vision(self, stamps):
        """(k,) stamps -> n (inst, k), ids (inst, k, maxFeatures), y (.., 2), provided_p (.., 3), true_sensor (k, 23)."""
        t = np.ascontiguousarray(stamps, dtype=np.float64).reshape(-1)
        I, k, F = len(self.configs), t.shape[0], self.cfg.maxFeatures
        n = np.zeros((I, k), dtype=np.int32)
        ids = np.zeros((I, k, F), dtype=np.int32)
        y = np.zeros((I, k, F, 2))
        p = np.zeros((I, k, F, 3))
        sensor = np.zeros((k, 23))
        ms = C.c_float(0.0)
        if lib.eqvio_sim_vision(self._h, k, _pd(t), _pi(n), _pi(ids), _pd(y), _pd(p), _pd(sensor), C.byref(ms)) != 0:
            raise RuntimeError((lib.eqvio_sim_last_error() or b"").decode())
        self.last_vision_ms = float(ms.value)
        return n, ids, y, p, sensor

    def record_streams(self, num_frames):
        """The eqvio_sim event loop (image wins ties, SimulationDataServer.cpp:173-187) for every instance: a list of
        simdata.SimStream, interchangeable with simdata.record_stream(cfg, num_frames)."""
        from simdata.vio_simulator import Frame, SimStream

        c = self.cfg
        img_t, imu_t, imu_of = [], [], []
        n_img = n_imu = 0
        while len(img_t) < num_frames:
            t_img, t_imu = n_img / c.imageFreq, n_imu / c.imuFreq
            if min(t_img, t_imu) >= c.duration:
                break
            if t_img <= t_imu:
                img_t.append(t_img)
                n_img += 1
            else:
                imu_t.append(t_imu)
                imu_of.append(len(img_t))  # belongs to the next image
                n_imu += 1
        if imu_t and c.inputNoise:
            rows_all = self.imu_instances(np.array(imu_t))
        else:
            rows_all = np.broadcast_to(self.imu(np.array(imu_t)) if imu_t else np.zeros((0, 13)), (len(self.configs), len(imu_t), 13))
        imu_of = np.array(imu_of, dtype=np.int64)
        n, ids, y, p, sensor = self.vision(np.array(img_t))
        cam = dict(width=c.width, height=c.height, fx=c.fx, fy=c.fy, cx=c.cx, cy=c.cy)
        streams = []
        for i, cfg in enumerate(self.configs):
            frames = []
            rows = rows_all[i]
            for k, t in enumerate(img_t):
                m = int(n[i, k])
                frames.append(Frame(t, ids[i, k, :m].astype(np.int64), y[i, k, :m].copy(), p[i, k, :m].copy(), rows[imu_of == k].copy(),
                                    sensor[k].copy()))
            # initial condition: the true state at t = 0 truncated to the first frame's visible ids, in shuffled world-point order
            f0 = frames[0]
            order = {int(pid): j for j, pid in enumerate(self._ids[i])}
            sel = sorted(range(len(f0.ids)), key=lambda j: order[int(f0.ids[j])])
            streams.append(SimStream(cfg, f0.true_sensor.copy(), f0.provided_p[sel].copy(), f0.ids[sel].copy(), frames, cam))
        return streams
