"""VIOSimulator input generator (plain numpy, no dependencies inside this repo).

What the reference does (path:line relative to the reference checkout):
  trajectory "wave"       src/dataserver/SimulationDataServer.cpp:46-65  (circle r=1 m, period 20 s,
                          yaw = angle, z = 0.2 sin(10 angle); pose rate 10*max(200,20) Hz,
                          initialTime 0.5/200 -- :223-232, :145)
  world points            src/VIOSimulator.cpp:63-126   (uniform in the inflated trajectory bounding
                          box, one coordinate snapped to a wall, ids assigned before the shuffle)
  IMU                     src/VIOSimulator.cpp:128-214  (gyro from consecutive poses, acceleration
                          from a cubic fit through four poses)
  vision                  src/VIOSimulator.cpp:216-265  (SE(3) geodesic pose interpolation, keep the
                          first maxFeatures visible points in shuffled order, id-sorted pixels)
  true state              src/VIOSimulator.cpp:269-310
  event order             src/dataserver/SimulationDataServer.cpp:173-187 (image wins ties)
  camera / extrinsics     src/dataserver/SimulationDataServer.cpp:156-176, 234-236

Deliberate differences: the reference draws world points from libc rand() and noise
from std::random_device; here both come from seeded numpy Generators so that a stream can be recorded
once and replayed by every implementation.  The initial condition is the true state at t=0 truncated
to the first frame's visible ids, in shuffled world-point order.
"""
from dataclasses import dataclass, field

import numpy as np

GRAVITY_CONSTANT = 9.80665
E3 = np.array([0.0, 0.0, 1.0])


# ---- quaternion (w,x,y,z) / SE(3) helpers, Eigen + LiePP conventions -------------------------------
def _skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def _qmul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx], -1)


def _qinv(q):
    n2 = np.sum(q * q, -1, keepdims=True)
    return np.concatenate([q[..., 0:1], -q[..., 1:4]], -1) / n2


def _qrot(q, v):
    u = q[..., 1:4]
    uv = np.cross(u, v)
    uv = uv + uv
    return v + q[..., 0:1] * uv + np.cross(u, uv)


def _qmat(q):
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy], [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def _mat2quat(m):
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1] = (m[2, 1] - m[1, 2]) * t
        q[2] = (m[0, 2] - m[2, 0]) * t
        q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def _so3_exp(w):
    w = np.asarray(w, dtype=np.float64)
    n = np.sqrt(np.sum(w * w, -1))
    theta = n / 2.0
    with np.errstate(divide="ignore", invalid="ignore"):
        wn = np.where(n[..., None] > 0, w / n[..., None], w)
    qb = np.concatenate([np.cos(theta)[..., None], np.sin(theta)[..., None] * wn], -1)
    qs = np.concatenate([np.ones_like(theta)[..., None], w / 2.0], -1)
    return np.where((theta > 1e-6)[..., None], qb, qs)


def _so3_log(q):
    R = _qmat(q)
    theta = np.arccos((R[0, 0] + R[1, 1] + R[2, 2] - 1.0) / 2.0)
    coef = theta / (2.0 * np.sin(theta)) if abs(theta) > 1e-6 else 0.5
    O = coef * (R - R.T)
    return np.array([O[2, 1], O[0, 2], O[1, 0]])


def _se3_mul(a, b):
    return (_qmul(a[0], b[0]), a[1] + _qrot(a[0], b[1]))


def _se3_inv(a):
    qi = _qinv(a[0])
    return (qi, -_qrot(qi, a[1]))


def _se3_exp(u):
    w, v = u[0:3], u[3:6]
    th = np.sqrt(w @ w)
    if abs(th) > 1e-12:
        A = np.sin(th) / th
        B = (1.0 - np.cos(th)) / th**2
        Cc = (1.0 - A) / th**2
    else:
        A, B, Cc = 1.0, 0.5, 1.0 / 6.0
    wx = _skew(w)
    R = np.eye(3) + A * wx + B * (wx @ wx)
    V = np.eye(3) + B * wx + Cc * (wx @ wx)
    return (_mat2quat(R), V @ v)


def _se3_log(P):
    om = _so3_log(P[0])
    O = _skew(om)
    theta = np.sqrt(np.sum(om**2))
    coef = 1.0 / 12.0
    if abs(theta) > 1e-6:
        coef = 1.0 / (theta * theta) * (1.0 - (theta * np.sin(theta)) / (2.0 * (1.0 - np.cos(theta))))
    VInv = np.eye(3) - 0.5 * O + coef * (O @ O)
    return np.concatenate([om, VInv @ P[1]])


# ---- configuration ------------------------------------------------------------------------------------
@dataclass
class SimConfig:
    """`sim` YAML node (VIOSimulator.cpp:47-60, SimulationDataServer.cpp:224-232) + the noise
    magnitudes the simulator takes from VIOFilter::Settings."""

    numPoints: int = 1000
    wallDistance: float = 2.0
    randomSeed: int = 0
    numWalls: int = 1
    maxFeatures: int = 30
    inputNoise: bool = False
    outputNoise: bool = False
    duration: float = 100.0
    imuFreq: float = 200.0
    imageFreq: float = 20.0
    noiseSeed: int = None
    # VIOFilterSettings.h defaults
    velGyrNoise: float = 1e-4
    velAccNoise: float = 1e-3
    velGyrBiasWalk: float = 1e-5
    velAccBiasWalk: float = 1e-3
    measurementNoise: float = 2.0
    # pinhole camera of generatePinholeCameraSquare (SimulationDataServer.cpp:156-171)
    width: int = 752
    height: int = 480
    fx: float = 458.654
    fy: float = 457.296
    cx: float = 367.215
    cy: float = 248.375

    @staticmethod
    def benchmark(N, seed=0, **kw):
        """Benchmark configs of SURVEY.md 8(d): 4 walls at 2 m, 20 N world points so that >= N stay in
        view over the 20 s lap, maxFeatures = N."""
        c = SimConfig(numPoints=20 * N, wallDistance=2.0, randomSeed=seed, numWalls=4, maxFeatures=N, duration=20.0)
        for k, v in kw.items():
            setattr(c, k, v)
        return c


@dataclass
class Frame:
    """One vision event and the IMU samples that preceded it (what main_sim.cpp:128-184 feeds the filter)."""

    stamp: float
    ids: np.ndarray  # (n,) ascending measured ids
    y: np.ndarray  # (n,2) pixels
    provided_p: np.ndarray  # (n,3) true camera-frame positions of the measured ids
    imu: np.ndarray  # (k,13): stamp, gyr3, acc3, gyrBiasVel3, accBiasVel3
    true_sensor: np.ndarray = field(default_factory=lambda: np.zeros(23))


@dataclass
class SimStream:
    config: SimConfig
    init_sensor: np.ndarray  # 23: bias6 | pose q(wxyz) x | vel | cameraOffset q x
    init_p: np.ndarray  # (N0,3)
    init_ids: np.ndarray  # (N0,)
    frames: list
    camera: dict  # width, height, fx, fy, cx, cy (pinhole)


class _Simulator:
    def __init__(self, cfg: SimConfig):
        self.cfg = cfg
        frequency = 10 * max(200.0, 20.0)
        initialTime = 0.5 / 200.0
        numPoses = int(np.floor(cfg.duration * frequency))
        t0 = np.arange(numPoses) / frequency + initialTime
        angle = 3.14 * 2 * t0 / 20.0
        self.q = _so3_exp(np.stack([np.zeros_like(angle), np.zeros_like(angle), angle], -1))
        self.x = np.stack([np.cos(angle), np.sin(angle), 0.2 * np.sin(10 * angle)], -1)
        self.t = t0 - initialTime
        R = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
        self.camOffset = (_mat2quat(R), np.zeros(3))
        self.rng_world = np.random.default_rng(cfg.randomSeed)
        self.rng_noise = np.random.default_rng(cfg.randomSeed if cfg.noiseSeed is None else cfg.noiseSeed)
        self._world_points()

    def _world_points(self):
        c = self.cfg
        num, numWalls, distance = c.numPoints, c.numWalls, c.wallDistance
        tmin, tmax = self.x.min(0), self.x.max(0)
        temp = 0.8 * np.array([numWalls > 0, numWalls > 1, numWalls > 3], dtype=np.float64) + 0.2 * np.ones(3)
        scaling = tmax - tmin + 2 * distance * temp
        offset = tmin - distance * temp
        p = 0.5 * (self.rng_world.uniform(-1.0, 1.0, (num, 3)) + 1.0)
        p = p * scaling + offset
        wall = (numWalls * np.arange(num)) // num
        for w, (axis, hi) in {0: (0, True), 1: (1, True), 2: (1, False), 3: (0, False), 4: (2, False), 5: (2, True)}.items():
            sel = wall == w
            p[sel, axis] = offset[axis] + (scaling[axis] if hi else 0.0)
        p[wall > 5, 2] = offset[2]
        ids = np.arange(num, dtype=np.int64)
        perm = self.rng_world.permutation(num)
        self.pointsP, self.pointsId = p[perm], ids[perm]

    def _project(self, pc):
        c = self.cfg
        return np.stack([c.fx * pc[..., 0] / pc[..., 2] + c.cx, c.fy * pc[..., 1] / pc[..., 2] + c.cy], -1)

    def _in_domain(self, pc):
        c = self.cfg
        with np.errstate(divide="ignore", invalid="ignore"):
            px = self._project(pc)
        ok = (px[..., 0] >= 0) & (px[..., 1] >= 0) & (px[..., 0] < c.width) & (px[..., 1] < c.height)
        return ok & (pc[..., 2] > 0)

    def _time_index(self, t):
        return int(np.searchsorted(self.t, t, side="left"))

    def _clamp(self, it):
        M = self.t.shape[0]
        while it + 1 >= M:
            it -= 1
        while it - 2 <= 0:
            it += 1
        return it

    def _inertial_states(self, it, ct):
        idx = [it - 2, it - 1, it, it + 1]
        tau = self.t[idx] - ct
        positionMat = self.x[idx].T
        timeMat = np.stack([np.ones(4), tau, tau * tau / 2.0, tau * tau * tau / 6.0], 0)
        return (positionMat @ timeMat.T @ np.linalg.inv(timeMat @ timeMat.T))[:, 0:3]

    def imu(self, t):
        c = self.cfg
        it = self._time_index(t)
        if it == self.t.shape[0]:
            gyr = np.zeros(3)
            acc = _qrot(_qinv(self.q[-1]), np.array([0.0, 0.0, GRAVITY_CONSTANT]))
            return np.concatenate([[t], gyr, acc, np.zeros(6)])
        it = self._clamp(it)
        q1, q2, t1, t2 = self.q[it - 1], self.q[it], self.t[it - 1], self.t[it]
        gyr = _so3_log(_qmul(_qinv(q1), q2)) / (t2 - t1)
        att = _qmul(q1, _so3_exp((t - t1) * gyr))
        a = self._inertial_states(it, t)[:, 2]
        acc = _qrot(_qinv(att), a - np.array([0.0, 0.0, -GRAVITY_CONSTANT]))
        v = np.concatenate([gyr, acc, np.zeros(6)])
        if c.inputNoise:
            var = np.repeat(np.array([c.velGyrNoise, c.velAccNoise, c.velGyrBiasWalk, c.velAccBiasWalk]) ** 2, 3) * max(c.imuFreq, 0.0)
            v = v + np.sqrt(var) * self.rng_noise.standard_normal(12)
        return np.concatenate([[t], v])

    def vision(self, t):
        c = self.cfg
        it = self._time_index(t)
        if it == self.t.shape[0]:
            return np.zeros(0, dtype=np.int64), np.zeros((0, 2))
        while it - 1 < 0:
            it += 1
        pose0 = (self.q[it - 1], self.x[it - 1])
        pose1 = (self.q[it], self.x[it])
        vel = _se3_log(_se3_mul(_se3_inv(pose0), pose1)) / (self.t[it] - self.t[it - 1])
        cur = _se3_mul(pose0, _se3_exp(vel * (t - self.t[it - 1])))
        ci = _se3_inv(_se3_mul(cur, self.camOffset))
        pc = _qrot(ci[0], self.pointsP) + ci[1]
        vis = np.nonzero(self._in_domain(pc))[0]
        if vis.shape[0] > c.maxFeatures:
            vis = vis[:c.maxFeatures]
        px = self._project(pc[vis])
        ids = self.pointsId[vis]
        order = np.argsort(ids, kind="stable")
        ids, px = ids[order], px[order]
        if c.outputNoise:
            px = px + (c.measurementNoise * self.rng_noise.standard_normal(2 * len(ids))).reshape(-1, 2)
        return ids, px

    def full_state(self, t):
        it = self._clamp(self._time_index(t))
        q0, q1, t0, t1 = self.q[it - 1], self.q[it], self.t[it - 1], self.t[it]
        w = _so3_log(_qmul(_qinv(q0), q1)) / (t1 - t0)
        pq = _qmul(q0, _so3_exp(w * (t - t0)))
        st = self._inertial_states(it, t)
        px = st[:, 0].copy()
        vel = _qrot(_qinv(pq), st[:, 1])
        sensor = np.concatenate([np.zeros(6), pq, px, vel, self.camOffset[0], self.camOffset[1]])
        ci = _se3_inv(_se3_mul((pq, px), self.camOffset))
        p = _qrot(ci[0], self.pointsP) + ci[1]
        return sensor, p, self.pointsId


def record_stream(cfg: SimConfig, num_frames: int) -> SimStream:
    """Run the eqvio_sim event loop (image wins ties) and record `num_frames` vision events; the first
    one is the t=0 image, which only augments landmarks (integration refuses newTime <= currentTime)."""
    sim = _Simulator(cfg)
    sensor0, p0, ids0 = sim.full_state(0.0)
    vis_ids0, _ = sim.vision(0.0)  # the visibility probe consumes noise draws like any other image
    keep = np.isin(ids0, vis_ids0)
    frames, imu_buf = [], []
    n_img = n_imu = 0
    while len(frames) < num_frames:
        t_img, t_imu = n_img / cfg.imageFreq, n_imu / cfg.imuFreq
        if min(t_img, t_imu) >= cfg.duration:
            break
        if t_img <= t_imu:
            ids, y = sim.vision(t_img)
            n_img += 1
            sensor, p, pid = sim.full_state(t_img)
            lookup = np.full(cfg.numPoints, -1, dtype=np.int64)
            lookup[pid] = np.arange(pid.shape[0])
            prov = p[lookup[ids]].reshape(-1, 3)
            imu = np.array(imu_buf, dtype=np.float64).reshape(-1, 13)
            frames.append(Frame(t_img, ids.astype(np.int64), y, prov, imu, sensor))
            imu_buf = []
        else:
            imu_buf.append(sim.imu(t_imu))
            n_imu += 1
    cam = dict(width=cfg.width, height=cfg.height, fx=cfg.fx, fy=cfg.fy, cx=cfg.cx, cy=cfg.cy)
    return SimStream(cfg, sensor0, p0[keep], ids0[keep].astype(np.int64), frames, cam)

