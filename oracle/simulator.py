"""VIOSimulator restatement -- the synthetic input generator (TEST
INFRASTRUCTURE, see oracle/__init__.py).

Follows src/VIOSimulator.cpp:36-310 and src/dataserver/SimulationDataServer.cpp:22-237.
Differences, all deliberate and listed in SURVEY.md section 8(c):
  * the reference seeds its noise generator from std::random_device
    (src/mathematical/Geometry.cpp:22-23) and its world points from libc
    rand(); this restatement owns a seeded numpy Generator so that a stream can
    be *recorded* once and replayed by both the CPU oracle and the CUDA path;
  * the filter's initial condition is the true state truncated to the first
    frame's visible ids in the simulator's shuffled world-point order (the
    reference would allocate a (21+3*numPoints)^2 covariance and erase the rest
    on the first frame -- same result, see SURVEY.md "scale trap").
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import liegroups as lg
from .camera import simulationCamera
from .eqf import GRAVITY_CONSTANT, IMUVelocity, Settings, VIOState, VisionMeasurement, getCoordinates
from .liegroups import SE3


def generateWaveTrajectory(endTime, frequency, initialTime):
    """SimulationDataServer.cpp:46-65.  Returns (t, q, x) arrays."""
    numPoses = int(np.floor(endTime * frequency))
    circleTime = 20.0
    i = np.arange(numPoses)
    t0 = i / frequency + initialTime
    angle = 3.14 * 2 * t0 / circleTime
    w = np.stack([np.zeros_like(angle), np.zeros_like(angle), angle], -1)
    q = lg.so3_exp(w)
    x = np.stack([np.cos(angle), np.sin(angle), 0.2 * np.sin(10 * angle)], -1)
    return t0 - initialTime, q, x


def generateSinTrajectory(endTime, frequency, initialTime):
    """SimulationDataServer.cpp:114-138."""
    numPoses = int(np.floor(endTime * frequency))
    sinTime = 20.0
    i = np.arange(numPoses)
    t0 = i / frequency + initialTime
    x = np.stack([0.5 * np.cos(2 * t0 / sinTime * 2 * 3.14), 0.5 * np.cos(t0 / sinTime * 2 * 3.14),
                  0.5 * np.cos(1.5 * t0 / sinTime * 2 * 3.14)], -1)
    att = np.stack([np.cos(5 * t0 / sinTime) * 3.14 / 4, np.cos(-6 * t0 / sinTime) * 3.14 / 4,
                    np.cos(4 * t0 / sinTime) * 3.14 / 4], -1)
    return t0 - initialTime, lg.so3_exp(att), x


def generateLineTrajectory(endTime, frequency, initialTime):
    """SimulationDataServer.cpp:22-44."""
    numPoses = int(np.floor(endTime * frequency))
    sinTime = 10.0
    i = np.arange(numPoses)
    t0 = i / frequency + initialTime
    newCoord = 5 * (2 * (t0 + np.sin(t0 * 3.14 * 2 / sinTime)) / endTime - 1)
    x = np.stack([np.zeros_like(t0), newCoord, np.zeros_like(t0)], -1)
    q = np.tile(lg.QUAT_IDENTITY, (numPoses, 1))
    return t0 - initialTime, q, x


@dataclass
class SimSettings:
    """YAML keys of the `sim` node (VIOSimulator.cpp:47-60, SimulationDataServer.cpp:224-232)."""

    numPoints: int = 1000
    wallDistance: float = 2.0
    randomSeed: int = 0
    numWalls: int = 1
    maxFeatures: int = 30
    initialNoise: bool = False
    inputNoise: bool = False
    outputNoise: bool = False
    duration: float = 100.0
    trajectory: str = "wave"
    imuFreq: float = 200.0
    imageFreq: float = 20.0


class VIOSimulator:
    def __init__(self, poses, cameraPtr, sim: SimSettings, filterSettings: Settings, noiseSeed=None):
        self.t, self.q, self.x = poses
        self.cameraPtr = cameraPtr
        self.sim = sim
        self.filterSettings = filterSettings
        self.cameraOffset = SE3()
        self.rng_world = np.random.default_rng(sim.randomSeed)
        self.rng_noise = np.random.default_rng(sim.randomSeed if noiseSeed is None else noiseSeed)
        self.pointsP, self.pointsId = self.generateWorldPoints(sim.numPoints, sim.wallDistance, sim.numWalls)
        self.maxFeatures = sim.maxFeatures

    # VIOSimulator.cpp:63-126
    def generateWorldPoints(self, num, distance, numWalls):
        tmin = self.x.min(0)
        tmax = self.x.max(0)
        temp = 0.8 * np.array([numWalls > 0, numWalls > 1, numWalls > 3], dtype=np.float64) + 0.2 * np.ones(3)
        scaling = tmax - tmin + 2 * distance * temp
        offset = tmin - distance * temp
        p = 0.5 * (self.rng_world.uniform(-1.0, 1.0, (num, 3)) + 1.0)
        p = p * scaling + offset
        wall = (numWalls * np.arange(num)) // num
        for i in range(num):
            w = wall[i]
            if w == 0:
                p[i, 0] = offset[0] + scaling[0]
            elif w == 1:
                p[i, 1] = offset[1] + scaling[1]
            elif w == 2:
                p[i, 1] = offset[1]
            elif w == 3:
                p[i, 0] = offset[0]
            elif w == 4:
                p[i, 2] = offset[2]
            elif w == 5:
                p[i, 2] = offset[2] + scaling[2]
            else:
                p[i, 2] = offset[2]
        ids = np.arange(num, dtype=np.int64)  # ids assigned before the shuffle
        perm = self.rng_world.permutation(num)
        return p[perm], ids[perm]

    def _timeIndex(self, t):  # :36-40 lower_bound
        return int(np.searchsorted(self.t, t, side="left"))

    def _sampleGaussianDiag(self, var):
        """sampleGaussianDistribution (Geometry.cpp:38-52) for a diagonal covariance."""
        return np.sqrt(var) * self.rng_noise.standard_normal(var.shape[0])

    # :172-214
    def getInertialStates(self, it, ct):
        idx = [it - 2, it - 1, it, it + 1]
        tau = self.t[idx] - ct
        positionMat = self.x[idx].T  # 3x4
        timeMat = np.stack([np.ones(4), tau, tau * tau / 2.0, tau * tau * tau / 6.0], 0)  # 4x4, cols per pose
        AMat = positionMat @ timeMat.T @ np.linalg.inv(timeMat @ timeMat.T)
        return AMat[:, 0:3]

    def _clampIt(self, it):
        M = self.t.shape[0]
        while it + 1 >= M:
            it -= 1
        while it - 2 <= 0:
            it += 1
        return it

    # :128-170
    def getIMU(self, currentTime, samplingFrequency=-1.0):
        imu = IMUVelocity(currentTime)
        it = self._timeIndex(currentTime)
        M = self.t.shape[0]
        if it == M:
            imu.gyr = np.zeros(3)
            imu.acc = lg.quat_rotate(lg.quat_inv(self.q[-1]), np.array([0.0, 0.0, GRAVITY_CONSTANT]))
            return imu
        it = self._clampIt(it)
        q1, q2 = self.q[it - 1], self.q[it]
        t1, t2 = self.t[it - 1], self.t[it]
        imu.gyr = lg.so3_log(lg.quat_mul(lg.quat_inv(q1), q2)) / (t2 - t1)
        imuAtt = lg.quat_mul(q1, lg.so3_exp((currentTime - t1) * imu.gyr))
        inertialAccel = self.getInertialStates(it, currentTime)[:, 2]
        imu.acc = lg.quat_rotate(lg.quat_inv(imuAtt), inertialAccel - np.array([0.0, 0.0, -GRAVITY_CONSTANT]))
        if self.sim.inputNoise:
            var = np.diag(self.filterSettings.constructInputGainMatrix()) * max(samplingFrequency, 0.0)
            imu = imu + self._sampleGaussianDiag(var)
            imu.stamp = currentTime
        return imu

    # :216-265
    def getVision(self, currentTime):
        meas = VisionMeasurement(currentTime, {}, self.cameraPtr)
        it = self._timeIndex(currentTime)
        if it == self.t.shape[0]:
            return meas
        while it - 1 < 0:
            it += 1
        pose0 = SE3(self.q[it - 1], self.x[it - 1])
        pose1 = SE3(self.q[it], self.x[it])
        vel = lg.se3_log(pose0.inverse() * pose1) / (self.t[it] - self.t[it - 1])
        currentPose = pose0 * lg.se3_exp(vel * (currentTime - self.t[it - 1]))
        camPoseInv = (currentPose * self.cameraOffset).inverse()
        pc = lg.quat_rotate(camPoseInv.q, self.pointsP) + camPoseInv.x
        vis = np.nonzero(self.cameraPtr.isInDomain(pc))[0]
        if vis.shape[0] > self.maxFeatures:
            vis = vis[: self.maxFeatures]
        px = self.cameraPtr.projectPoint(pc[vis])
        meas.camCoordinates = {int(self.pointsId[i]): px[k] for k, i in enumerate(vis)}
        if self.sim.outputNoise:
            n = len(meas.camCoordinates)
            var = np.full(2 * n, self.filterSettings.measurementNoise ** 2)
            meas = meas.plusVector(self._sampleGaussianDiag(var))
            meas.stamp = currentTime
        return meas

    # :269-310
    def getFullState(self, time, allowNoise=False):
        it = self._clampIt(self._timeIndex(time))
        q0, q1 = self.q[it - 1], self.q[it]
        t0, t1 = self.t[it - 1], self.t[it]
        angularVel = lg.so3_log(lg.quat_mul(lg.quat_inv(q0), q1)) / (t1 - t0)
        xi = VIOState()
        xi.sensor.inputBias = np.zeros(6)
        xi.sensor.pose.q = lg.quat_mul(q0, lg.so3_exp(angularVel * (time - t0)))
        inertialStates = self.getInertialStates(it, time)
        xi.sensor.pose.x = inertialStates[:, 0].copy()
        xi.sensor.velocity = lg.quat_rotate(lg.quat_inv(xi.sensor.pose.q), inertialStates[:, 1])
        xi.sensor.cameraOffset = self.cameraOffset.copy()
        camPoseInv = (xi.sensor.pose * self.cameraOffset).inverse()
        xi.p = lg.quat_rotate(camPoseInv.q, self.pointsP) + camPoseInv.x
        xi.ids = self.pointsId.copy()
        if allowNoise and self.sim.initialNoise:
            var = np.diag(self.filterSettings.constructInitialStateCovariance(xi.N))
            eps = self._sampleGaussianDiag(var)
            xi = getCoordinates(self.filterSettings.coordinateChoice).stateChart.chartInv(eps, xi)
        return xi


@dataclass
class Frame:
    """One recorded vision event and the IMU samples that preceded it."""

    stamp: float
    ids: np.ndarray  # (n,) ascending measured ids
    y: np.ndarray  # (n,2) pixels
    provided_p: np.ndarray  # (n,3) true camera-frame positions of the measured ids (augmentLandmarkStates input)
    imu: np.ndarray  # (k,13): stamp, gyr3, acc3, gyrBiasVel3, accBiasVel3 -- samples since the previous frame
    true_sensor: np.ndarray = field(default_factory=lambda: np.zeros(23))


class SimulationDataServer:
    """Event generator of eqvio_sim (SimulationDataServer.cpp:173-237, main_sim.cpp:128-184)."""

    def __init__(self, sim: SimSettings, filterSettings: Settings, noiseSeed=None):
        self.sim = sim
        self.filterSettings = filterSettings
        self.imuFreq, self.imageFreq = sim.imuFreq, sim.imageFreq
        self.maxSimulationTime = sim.duration
        # generateTrajectory runs before imuFreq/imageFreq are read from YAML, so the
        # pose rate always uses the 200/20 Hz defaults (:223-232); initialTime = 0.5/imuFreq
        desiredFreq = 10 * max(200.0, 20.0)
        initialTime = 0.5 / 200.0
        gen = {"wave": generateWaveTrajectory, "sine": generateSinTrajectory, "line": generateLineTrajectory}.get(
            sim.trajectory, generateWaveTrajectory)
        poses = gen(self.maxSimulationTime, desiredFreq, initialTime)
        self.simulator = VIOSimulator(poses, simulationCamera(), sim, filterSettings, noiseSeed)
        R = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
        self.simulator.cameraOffset.q = lg.matrix_to_quat(R)
        self.imuMeasCount = 0
        self.imageMeasCount = 0

    def cameraExtrinsics(self):
        return self.simulator.cameraOffset.copy()

    def nextImageTime(self):
        return self.imageMeasCount / self.imageFreq

    def nextIMUTime(self):
        return self.imuMeasCount / self.imuFreq

    def nextMeasurementType(self):  # :177-187 (image wins ties)
        if min(self.nextImageTime(), self.nextIMUTime()) >= self.maxSimulationTime:
            return None
        return "image" if self.nextImageTime() <= self.nextIMUTime() else "imu"

    def getSimVision(self):
        m = self.simulator.getVision(self.nextImageTime())
        self.imageMeasCount += 1
        return m

    def getSimIMU(self):
        m = self.simulator.getIMU(self.nextIMUTime(), self.imuFreq)
        self.imuMeasCount += 1
        return m

    def getTrueState(self, stamp, withNoise=False):
        return self.simulator.getFullState(stamp, withNoise)

    def initialCondition(self):
        """True state at t=0 truncated to the first frame's visible ids, in the
        simulator's shuffled world-point order (see module docstring)."""
        full = self.simulator.getFullState(0.0, True)
        vis = self.simulator.getVision(0.0)
        keep = np.isin(full.ids, np.array(vis.getIds(), dtype=np.int64))
        return VIOState(full.sensor, full.p[keep], full.ids[keep])

    def record(self, numFrames):
        """Run the eqvio_sim event loop and record `numFrames` vision events
        (the first one is the t=0 image that only augments landmarks)."""
        frames = []
        imuBuf = []
        while len(frames) < numFrames:
            kind = self.nextMeasurementType()
            if kind is None:
                break
            if kind == "image":
                meas = self.getSimVision()
                true = self.getTrueState(meas.stamp, True)
                ids, y = meas.arrays()
                order = {int(i): k for k, i in enumerate(true.ids)}
                prov = np.array([true.p[order[int(i)]] for i in ids]).reshape(-1, 3)
                imu = np.array([np.concatenate([[u.stamp], u.asVector12()]) for u in imuBuf]).reshape(-1, 13)
                frames.append(Frame(meas.stamp, ids, y, prov, imu, true.sensor.flat()))
                imuBuf = []
            else:
                imuBuf.append(self.getSimIMU())
        return frames


def benchmarkSettings(coordinateChoice=0, **overrides):
    """Filter settings of the benchmark configs: struct
    defaults (VIOFilterSettings.h:58-99) with fastRiccati on, as both shipped
    dataset configs do (configs/EQVIO_config_EuRoC_stationary.yaml:45)."""
    st = Settings()
    st.fastRiccati = True
    st.coordinateChoice = coordinateChoice
    for k, v in overrides.items():
        setattr(st, k, v)
    return st


def benchmarkSim(N, seed=0, **overrides):
    """Simulator settings of the benchmark configs: wave trajectory, 200/20 Hz,
    4 walls at 2 m, 20*N world points so that >= N stay in view."""
    sim = SimSettings(numPoints=20 * N, wallDistance=2.0, randomSeed=seed, numWalls=4, maxFeatures=N,
                      duration=20.0, trajectory="wave")
    for k, v in overrides.items():
        setattr(sim, k, v)
    return sim


def replayOracle(filter_, frames, cam, onUpdate=None):
    """Drive an oracle VIOFilter with a recorded stream the way main_sim.cpp:128-184 does."""
    from .eqf import VIOState as _VS

    for k, fr in enumerate(frames):
        for row in fr.imu:
            filter_.processIMUData(IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
        meas = VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, cam)
        provided = _VS(None, fr.provided_p, fr.ids)
        filter_.augmentLandmarkStates(meas.getIds(), provided)
        filter_.processVisionData(meas)
        if onUpdate is not None:
            onUpdate(k, filter_)
