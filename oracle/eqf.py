"""EqF mathematics + filter driver of the oracle (TEST INFRASTRUCTURE, see
oracle/__init__.py).

A numpy fp64 restatement that follows the reference's *dense* evaluation order
(zero-filled A / B / C, LU inverse of S, Sigma - K C Sigma), so that it doubles
as the CPU baseline.  Names mirror the reference so tests read like the
reference's own.  Landmark-sized work is vectorised over the landmark axis.

Reference files followed (path:line relative to the reference checkout):
  include/eqvio/mathematical/IMUVelocity.h:26-84, src/mathematical/IMUVelocity.cpp:20-77
  src/mathematical/VIOGroup.cpp:25-290          (actions, product, lifts, exp)
  src/mathematical/VIOState.cpp:27-401          (system function, charts, differentials)
  src/mathematical/EqFMatrices.cpp:24-89        (A discrete, C assembly, C_i)
  src/mathematical/coordinateSuite/euclid.cpp:36-233
  src/mathematical/coordinateSuite/invdepth.cpp:36-266
  src/mathematical/coordinateSuite/normal.cpp:37-65
  src/mathematical/VIO_eqf.cpp:27-245           (Riccati, update, NEES, bookkeeping)
  src/mathematical/VisionMeasurement.cpp:24-89
  src/mathematical/Geometry.cpp:25-36           (numericalDifferential)
  include/eqvio/VIOFilterSettings.h:58-229      (settings and gain matrices)
  src/VIOFilter.cpp:31-380                      (filter driver)
"""

from __future__ import annotations

import copy
from dataclasses import dataclass, field

import numpy as np

from . import liegroups as lg
from .liegroups import SE3, skew

GRAVITY_CONSTANT = 9.80665  # IMUVelocity.h:26
SENSOR_DIM = 21  # VIOSensorState::CompDim

E3 = np.array([0.0, 0.0, 1.0])


# ----------------------------------------------------------------------------
# IMUVelocity
# ----------------------------------------------------------------------------


class IMUVelocity:
    __slots__ = ("stamp", "gyr", "acc", "gyrBiasVel", "accBiasVel")

    def __init__(self, stamp=0.0, gyr=None, acc=None, gyrBiasVel=None, accBiasVel=None):
        self.stamp = float(stamp)
        self.gyr = np.zeros(3) if gyr is None else np.array(gyr, dtype=np.float64)
        self.acc = np.zeros(3) if acc is None else np.array(acc, dtype=np.float64)
        self.gyrBiasVel = np.zeros(3) if gyrBiasVel is None else np.array(gyrBiasVel, dtype=np.float64)
        self.accBiasVel = np.zeros(3) if accBiasVel is None else np.array(accBiasVel, dtype=np.float64)

    @staticmethod
    def Zero():
        return IMUVelocity()

    @staticmethod
    def fromVector(vec):
        vec = np.asarray(vec, dtype=np.float64)
        if vec.shape[0] == 6:
            return IMUVelocity(0.0, vec[0:3], vec[3:6])
        return IMUVelocity(0.0, vec[0:3], vec[3:6], vec[6:9], vec[9:12])

    def __add__(self, other):  # IMUVelocity.cpp:42-50
        if not isinstance(other, IMUVelocity):
            other = IMUVelocity.fromVector(other)
        st = self.stamp if self.stamp > 0 else other.stamp
        return IMUVelocity(st, self.gyr + other.gyr, self.acc + other.acc, self.gyrBiasVel + other.gyrBiasVel,
                           self.accBiasVel + other.accBiasVel)

    def minusBias(self, vec6):
        """operator-(Matrix<6,1>) (IMUVelocity.cpp:52-58): bias velocities reset to zero."""
        return IMUVelocity(self.stamp, self.gyr - vec6[0:3], self.acc - vec6[3:6])

    def __mul__(self, c):  # IMUVelocity.cpp:69-77
        return IMUVelocity(self.stamp, self.gyr * c, self.acc * c, self.gyrBiasVel * c, self.accBiasVel * c)

    def asVector12(self):
        return np.concatenate([self.gyr, self.acc, self.gyrBiasVel, self.accBiasVel])


# ----------------------------------------------------------------------------
# States, group, algebra
# ----------------------------------------------------------------------------


class VIOSensorState:
    __slots__ = ("inputBias", "pose", "velocity", "cameraOffset")

    def __init__(self):
        self.inputBias = np.zeros(6)
        self.pose = SE3()
        self.velocity = np.zeros(3)
        self.cameraOffset = SE3()

    def copy(self):
        s = VIOSensorState()
        s.inputBias = self.inputBias.copy()
        s.pose = self.pose.copy()
        s.velocity = self.velocity.copy()
        s.cameraOffset = self.cameraOffset.copy()
        return s

    def gravityDir(self):  # VIOState.cpp:94
        return lg.quat_rotate(lg.quat_inv(self.pose.q), E3)

    def flat(self):
        """bias6 | pose q(w,x,y,z) x3 | vel3 | camOffset q4 x3 -- the C-ABI's sensor[23]."""
        return np.concatenate([self.inputBias, self.pose.q, self.pose.x, self.velocity, self.cameraOffset.q,
                               self.cameraOffset.x])

    @staticmethod
    def fromFlat(f):
        s = VIOSensorState()
        f = np.asarray(f, dtype=np.float64)
        s.inputBias = f[0:6].copy()
        s.pose = SE3(f[6:10], f[10:13])
        s.velocity = f[13:16].copy()
        s.cameraOffset = SE3(f[16:20], f[20:23])
        return s


class VIOState:
    """sensor + camera-frame landmarks (VIOState.h:41-90).  Landmarks are SoA:
    ``p`` (N,3) and ``ids`` (N,), in *state order*."""

    __slots__ = ("sensor", "p", "ids")

    def __init__(self, sensor=None, p=None, ids=None):
        self.sensor = VIOSensorState() if sensor is None else sensor
        self.p = np.zeros((0, 3)) if p is None else np.array(p, dtype=np.float64).reshape(-1, 3)
        self.ids = np.zeros(0, dtype=np.int64) if ids is None else np.array(ids, dtype=np.int64).reshape(-1)

    def copy(self):
        return VIOState(self.sensor.copy(), self.p.copy(), self.ids.copy())

    def getIds(self):
        return [int(i) for i in self.ids]

    def Dim(self):  # VIOState.cpp:102
        return SENSOR_DIM + 3 * self.p.shape[0]

    @property
    def N(self):
        return self.p.shape[0]


class VIOGroup:
    """(beta, A, w, B, Q_1..Q_N) (VIOGroup.h:32-38); Q as arrays Qq (N,4), Qa (N,)."""

    __slots__ = ("beta", "A", "w", "B", "Qq", "Qa", "ids")

    def __init__(self):
        self.beta = np.zeros(6)
        self.A = SE3()
        self.w = np.zeros(3)
        self.B = SE3()
        self.Qq = np.zeros((0, 4))
        self.Qa = np.zeros(0)
        self.ids = np.zeros(0, dtype=np.int64)

    @staticmethod
    def Identity(ids=()):  # VIOGroup.cpp:94-106
        X = VIOGroup()
        n = len(ids)
        X.ids = np.array(ids, dtype=np.int64).reshape(-1)
        X.Qq = np.tile(lg.QUAT_IDENTITY, (n, 1))
        X.Qa = np.ones(n)
        return X

    def copy(self):
        X = VIOGroup()
        X.beta = self.beta.copy()
        X.A = self.A.copy()
        X.w = self.w.copy()
        X.B = self.B.copy()
        X.Qq = self.Qq.copy()
        X.Qa = self.Qa.copy()
        X.ids = self.ids.copy()
        return X

    def __mul__(self, other):  # VIOGroup.cpp:71-92
        r = VIOGroup()
        r.beta = self.beta + other.beta
        r.A = self.A * other.A
        r.B = self.B * other.B
        r.w = self.w + lg.quat_rotate(self.A.q, other.w)
        assert np.array_equal(self.ids, other.ids)
        r.Qq = lg.quat_mul(self.Qq, other.Qq)
        r.Qa = self.Qa * other.Qa
        r.ids = self.ids.copy()
        return r

    def inverse(self):  # VIOGroup.cpp:108-121
        r = VIOGroup()
        r.beta = -self.beta
        r.A = self.A.inverse()
        r.B = self.B.inverse()
        r.w = -lg.quat_rotate(lg.quat_inv(self.A.q), self.w)
        r.Qq = lg.quat_inv(self.Qq) if self.Qq.shape[0] else self.Qq.copy()
        r.Qa = 1.0 / self.Qa
        r.ids = self.ids.copy()
        return r

    def hasNaN(self):
        return bool(
            np.isnan(self.beta).any() or np.isnan(self.A.q).any() or np.isnan(self.A.x).any()
            or np.isnan(self.B.q).any() or np.isnan(self.B.x).any() or np.isnan(self.w).any()
            or np.isnan(self.Qq).any() or np.isnan(self.Qa).any())

    def sensorFlat(self):
        """beta6 | A q4 x3 | w3 | B q4 x3 -- the C-ABI's group sensor[23]."""
        return np.concatenate([self.beta, self.A.q, self.A.x, self.w, self.B.q, self.B.x])


class VIOAlgebra:
    __slots__ = ("u_beta", "U_A", "U_B", "u_w", "W", "ids")

    def __init__(self):
        self.u_beta = np.zeros(6)
        self.U_A = np.zeros(6)
        self.U_B = np.zeros(6)
        self.u_w = np.zeros(3)
        self.W = np.zeros((0, 4))
        self.ids = np.zeros(0, dtype=np.int64)

    def __mul__(self, c):  # VIOGroup.cpp:145-156
        r = VIOAlgebra()
        r.u_beta, r.U_A, r.U_B, r.u_w, r.W, r.ids = (self.u_beta * c, self.U_A * c, self.U_B * c, self.u_w * c,
                                                     self.W * c, self.ids.copy())
        return r

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1.0

    def __add__(self, o):  # VIOGroup.cpp:171-191
        r = VIOAlgebra()
        assert np.array_equal(self.ids, o.ids)
        r.u_beta, r.U_A, r.U_B, r.u_w, r.W, r.ids = (self.u_beta + o.u_beta, self.U_A + o.U_A, self.U_B + o.U_B,
                                                     self.u_w + o.u_w, self.W + o.W, self.ids.copy())
        return r

    def __sub__(self, o):
        return self + (-o)


class VisionMeasurement:
    """stamp + id->pixel map + camera (VisionMeasurement.h:35-40).  The map is a
    dict; every consumer iterates it in ascending id like std::map."""

    __slots__ = ("stamp", "camCoordinates", "cameraPtr")

    def __init__(self, stamp=0.0, camCoordinates=None, cameraPtr=None):
        self.stamp = float(stamp)
        self.camCoordinates = {} if camCoordinates is None else dict(camCoordinates)
        self.cameraPtr = cameraPtr

    @staticmethod
    def fromArrays(stamp, ids, y, cameraPtr):
        y = np.asarray(y, dtype=np.float64).reshape(-1, 2)
        return VisionMeasurement(stamp, {int(i): y[k].copy() for k, i in enumerate(ids)}, cameraPtr)

    def getIds(self):  # VisionMeasurement.cpp:24-28
        return sorted(self.camCoordinates.keys())

    def arrays(self):
        ids = np.array(self.getIds(), dtype=np.int64)
        y = np.array([self.camCoordinates[int(i)] for i in ids], dtype=np.float64).reshape(-1, 2)
        return ids, y

    def copy(self):
        return VisionMeasurement(self.stamp, {k: v.copy() for k, v in self.camCoordinates.items()}, self.cameraPtr)

    def __sub__(self, other):  # VisionMeasurement.cpp:60-71
        d = {}
        for k in self.getIds():
            if k in other.camCoordinates:
                d[k] = self.camCoordinates[k] - other.camCoordinates[k]
        return VisionMeasurement(0.0, d, self.cameraPtr)

    def asVector(self):  # VisionMeasurement.cpp:72-79
        ids = self.getIds()
        if not ids:
            return np.zeros(0)
        return np.concatenate([self.camCoordinates[i] for i in ids])

    def plusVector(self, eta):  # VisionMeasurement.cpp:81-89
        r = self.copy()
        for k, i in enumerate(self.getIds()):
            r.camCoordinates[i] = r.camCoordinates[i] + eta[2 * k:2 * k + 2]
        return r


# ----------------------------------------------------------------------------
# Group actions
# ----------------------------------------------------------------------------


def sensorStateGroupAction(X, sensor):  # VIOGroup.cpp:25-32
    r = VIOSensorState()
    r.inputBias = sensor.inputBias + X.beta
    r.pose = sensor.pose * X.A
    r.velocity = lg.quat_rotate(lg.quat_inv(X.A.q), sensor.velocity - X.w)
    r.cameraOffset = X.A.inverse() * sensor.cameraOffset * X.B
    return r


def stateGroupAction(X, state):  # VIOGroup.cpp:34-55
    assert np.array_equal(X.ids, state.ids)
    p = lg.sot3_apply_inverse(X.Qq, X.Qa, state.p) if state.N else state.p.copy()
    return VIOState(sensorStateGroupAction(X, state.sensor), p, state.ids.copy())


def outputGroupAction(X, measurement):  # VIOGroup.cpp:57-69
    cam = measurement.cameraPtr
    out = {}
    for i, idn in enumerate(X.ids):
        idn = int(idn)
        if idn in measurement.camCoordinates:
            bearing = cam.undistortPoint(measurement.camCoordinates[idn])
            out[idn] = cam.projectPoint(lg.quat_rotate(lg.quat_inv(X.Qq[i]), bearing))
    return VisionMeasurement(0.0, out, cam)


def measureSystemState(state, cameraPtr):  # VIOState.cpp:70-78
    px = cameraPtr.projectPoint(state.p) if state.N else np.zeros((0, 2))
    return VisionMeasurement(0.0, {int(i): px[k] for k, i in enumerate(state.ids)}, cameraPtr)


# ----------------------------------------------------------------------------
# Lifts and exponential
# ----------------------------------------------------------------------------


def liftVelocity(state, velocity):  # VIOGroup.cpp:190-227
    lift = VIOAlgebra()
    sensor = state.sensor
    v_est = velocity.minusBias(sensor.inputBias)
    lift.u_beta = np.concatenate([velocity.gyrBiasVel, velocity.accBiasVel])
    lift.U_A = np.concatenate([v_est.gyr, sensor.velocity])
    lift.U_B = sensor.cameraOffset.inverse().Adjoint() @ lift.U_A
    lift.u_w = -v_est.acc + sensor.gravityDir() * GRAVITY_CONSTANT
    U_C = sensor.cameraOffset.inverse().Adjoint() @ lift.U_A
    omega_C, v_C = U_C[0:3], U_C[3:6]
    p = state.p
    n2 = np.sum(p * p, -1)
    W = np.zeros((state.N, 4))
    if state.N:
        W[:, 0:3] = omega_C + np.cross(p, v_C) / n2[:, None]
        W[:, 3] = (p @ v_C) / n2
    lift.W = W
    lift.ids = state.ids.copy()
    return lift


def liftVelocityDiscrete(state, velocity, dt):  # VIOGroup.cpp:229-271
    lift = VIOGroup()
    sensor = state.sensor
    v_est = velocity.minusBias(sensor.inputBias)
    lift.beta = dt * np.concatenate([velocity.gyrBiasVel, velocity.accBiasVel])
    lift.A.q = lg.so3_exp(dt * v_est.gyr)
    Rq = sensor.pose.q
    x = dt * lg.quat_rotate(Rq, sensor.velocity) + 0.5 * dt * dt * (
        lg.quat_rotate(Rq, v_est.acc) + np.array([0.0, 0.0, -GRAVITY_CONSTANT]))
    lift.A.x = lg.quat_rotate(lg.quat_inv(Rq), x)
    lift.B = sensor.cameraOffset.inverse() * lift.A * sensor.cameraOffset
    bodyVelocityDiff = v_est.acc - sensor.gravityDir() * GRAVITY_CONSTANT
    lift.w = sensor.velocity - (sensor.velocity + dt * bodyVelocityDiff)
    camChangeInv = sensor.cameraOffset.inverse() * lift.A.inverse() * sensor.cameraOffset
    p0 = state.p
    if state.N:
        p1 = lg.quat_rotate(camChangeInv.q, p0) + camChangeInv.x
        lift.Qq = lg.quat_from_two_vectors(lg.normalized(p1), lg.normalized(p0))
        lift.Qa = lg.norm(p0) / lg.norm(p1)
    else:
        lift.Qq = np.zeros((0, 4))
        lift.Qa = np.zeros(0)
    lift.ids = state.ids.copy()
    return lift


def VIOExp(lam):  # VIOGroup.cpp:273-290
    q, x0, x1 = lg.se23_exp(np.concatenate([lam.U_A, lam.u_w]))
    r = VIOGroup()
    r.beta = lam.u_beta.copy()
    r.A = SE3(q, x0)
    r.w = x1
    r.B = lg.se3_exp(lam.U_B)
    r.ids = lam.ids.copy()
    if lam.W.shape[0]:
        r.Qq, r.Qa = lg.sot3_exp(lam.W)
    else:
        r.Qq, r.Qa = np.zeros((0, 4)), np.zeros(0)
    return r


def integrateSystemFunction(state, velocity, dt):  # VIOState.cpp:27-68
    new = VIOState()
    sensor = state.sensor
    v_est = velocity.minusBias(sensor.inputBias)
    new.sensor.inputBias = sensor.inputBias + dt * np.concatenate([velocity.gyrBiasVel, velocity.accBiasVel])
    poseChange = SE3()
    poseChange.q = lg.so3_exp(dt * v_est.gyr)
    Rq = sensor.pose.q
    x = dt * lg.quat_rotate(Rq, sensor.velocity) + 0.5 * dt * dt * (
        lg.quat_rotate(Rq, v_est.acc) + np.array([0.0, 0.0, -GRAVITY_CONSTANT]))
    poseChange.x = lg.quat_rotate(lg.quat_inv(Rq), x)
    new.sensor.pose = sensor.pose * poseChange
    inertialVelocityDiff = sensor.pose.R @ v_est.acc + np.array([0.0, 0.0, -GRAVITY_CONSTANT])
    new.sensor.velocity = lg.quat_rotate(
        lg.quat_inv(new.sensor.pose.q), lg.quat_rotate(Rq, sensor.velocity) + dt * inertialVelocityDiff)
    camChangeInv = sensor.cameraOffset.inverse() * poseChange.inverse() * sensor.cameraOffset
    new.p = lg.quat_rotate(camChangeInv.q, state.p) + camChangeInv.x if state.N else state.p.copy()
    new.ids = state.ids.copy()
    new.sensor.cameraOffset = sensor.cameraOffset.copy()
    return new


# ----------------------------------------------------------------------------
# Sphere charts (VIOState.cpp:246-353) -- vectorised over the pole axis
# ----------------------------------------------------------------------------


def e3ProjectSphere(eta):  # :246-251
    return (eta[..., 0:2] - E3[0:2]) / (1.0 - eta[..., 2:3])


def e3ProjectSphereInv(y):  # :253-258
    yBar = np.concatenate([y, np.zeros(y.shape[:-1] + (1,))], -1)
    return E3 + 2.0 / (np.sum(yBar * yBar, -1, keepdims=True) + 1.0) * (yBar - E3)


def e3ProjectSphereDiff(eta):  # :260-267
    I3 = np.eye(3)
    M = I3 * (1.0 - eta[..., 2])[..., None, None] + (eta - E3)[..., :, None] * E3[None, :]
    D = M[..., 0:2, :]
    return (1.0 - eta[..., 2])[..., None, None] ** -2.0 * D


def e3ProjectSphereInvDiff(y):  # :269-275
    n2 = np.sum(y * y, -1)
    D = np.zeros(y.shape[:-1] + (3, 2))
    D[..., 0:2, 0:2] = np.eye(2) * (n2 + 1.0)[..., None, None] - 2.0 * y[..., :, None] * y[..., None, :]
    D[..., 2, :] = 2.0 * y
    return 2.0 * ((n2 + 1.0) ** -2.0)[..., None, None] * D


def sphereChart_stereo(eta, pole):  # :286-290
    rot = lg.quat_from_two_vectors(-pole, E3)
    return e3ProjectSphere(lg.quat_rotate(rot, eta))


def sphereChart_stereo_inv(y, pole):  # :292-296
    etaRot = e3ProjectSphereInv(y)
    rot = lg.quat_from_two_vectors(-pole, E3)
    return lg.quat_rotate(lg.quat_inv(rot), etaRot)


def sphereChart_stereo_diff0(pole):  # :298-302
    rot = lg.quat_from_two_vectors(-pole, E3)
    etaRot = lg.quat_rotate(rot, pole)
    return e3ProjectSphereDiff(etaRot) @ lg.quat_to_matrix(rot)


def sphereChart_stereo_inv_diff0(pole):  # :304-307
    rot = lg.quat_from_two_vectors(-pole, E3)
    D0 = e3ProjectSphereInvDiff(np.zeros(pole.shape[:-1] + (2,)))
    return lg.quat_to_matrix(lg.quat_inv(rot)) @ D0


def sphereChart_normal(eta, pole):  # :310-327
    rot = lg.quat_from_two_vectors(pole, E3)
    y = lg.quat_rotate(rot, eta)
    ye3 = np.cross(y, E3)
    sin_th = lg.norm(ye3)
    cos_th = y[..., 2]
    th = np.arctan2(sin_th, cos_th)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(np.abs(th) < 1e-8, 1.0, th / sin_th)
    omega = ye3 * scale[..., None]
    return omega[..., 0:2]


def sphereChart_normal_inv(eps, pole):  # :328-337
    omega = np.concatenate([eps, np.zeros(eps.shape[:-1] + (1,))], -1)
    y = lg.quat_rotate(lg.so3_exp(-omega), E3)
    rot = lg.quat_from_two_vectors(pole, E3)
    return lg.quat_rotate(lg.quat_inv(rot), y)


def sphereChart_normal_diff0(pole):  # :338-345
    rot = lg.quat_from_two_vectors(pole, E3)
    d = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0]])
    return d @ lg.quat_to_matrix(rot)


def sphereChart_normal_inv_diff0(pole):  # :346-353
    rot = lg.quat_from_two_vectors(pole, E3)
    d = np.array([[0.0, -1.0], [1.0, 0.0], [0.0, 0.0]])
    return lg.quat_to_matrix(lg.quat_inv(rot)) @ d


# ----------------------------------------------------------------------------
# Coordinate charts (VIOState.cpp:104-244)
# ----------------------------------------------------------------------------


def sensorChart_std(Xi, Xi0):  # :104-113
    eps = np.zeros(SENSOR_DIM)
    eps[0:6] = Xi.inputBias - Xi0.inputBias
    eps[6:12] = lg.se3_log(Xi0.pose.inverse() * Xi.pose)
    eps[12:15] = Xi.velocity - Xi0.velocity
    eps[15:21] = lg.se3_log(Xi0.cameraOffset.inverse() * Xi.cameraOffset)
    return eps


def sensorChart_std_inv(eps, Xi0):  # :114-121
    Xi = VIOSensorState()
    Xi.inputBias = Xi0.inputBias + eps[0:6]
    Xi.pose = Xi0.pose * lg.se3_exp(eps[6:12])
    Xi.velocity = Xi0.velocity + eps[12:15]
    Xi.cameraOffset = Xi0.cameraOffset * lg.se3_exp(eps[15:21])
    return Xi


def sensorChart_normal(Xi, Xi0):  # :123-137
    A = Xi0.pose.inverse() * Xi.pose
    v_xi0 = lg.quat_rotate(Xi0.pose.q, Xi0.velocity)
    v_xi = lg.quat_rotate(Xi.pose.q, Xi.velocity)
    v_A = lg.quat_rotate(lg.quat_inv(Xi0.pose.q), v_xi - v_xi0)
    B = Xi0.cameraOffset.inverse() * A * Xi.cameraOffset
    eps = np.zeros(SENSOR_DIM)
    eps[0:6] = Xi.inputBias - Xi0.inputBias
    eps[6:15] = lg.se23_log(A.q, A.x, v_A)
    eps[15:21] = lg.se3_log(B)
    return eps


def sensorChart_normal_inv(eps, Xi0):  # :138-152
    q, x0, x1 = lg.se23_exp(eps[6:15])
    B = lg.se3_exp(eps[15:21])
    A = SE3(q, x0)
    v_A = x1
    Xi = VIOSensorState()
    Xi.inputBias = Xi0.inputBias + eps[0:6]
    Xi.pose = Xi0.pose * A
    v_xi0 = lg.quat_rotate(Xi0.pose.q, Xi0.velocity)
    Xi.velocity = lg.quat_rotate(lg.quat_inv(Xi.pose.q), v_xi0 + lg.quat_rotate(Xi0.pose.q, v_A))
    Xi.cameraOffset = A.inverse() * Xi0.cameraOffset * B
    return Xi


def pointChart_euclid(p, p0):  # :154-158
    return p - p0


def pointChart_euclid_inv(eps, p0):
    return p0 + eps


def pointChart_invdepth(p, p0):  # :160-173
    rho = 1.0 / lg.norm(p)
    rho0 = 1.0 / lg.norm(p0)
    y = p * rho[..., None]
    y0 = p0 * rho0[..., None]
    return np.concatenate([sphereChart_stereo(y, y0), (rho - rho0)[..., None]], -1)


def pointChart_invdepth_inv(eps, p0):  # :174-188
    rho0 = 1.0 / lg.norm(p0)
    y0 = p0 * rho0[..., None]
    y = sphereChart_stereo_inv(eps[..., 0:2], y0)
    rho = eps[..., 2] + rho0
    rho = np.where(rho <= 0.0, 1e-6, rho)
    return y / rho[..., None]


def pointChart_normal(p, p0):  # :190-203
    rho = 1.0 / lg.norm(p)
    rho0 = 1.0 / lg.norm(p0)
    y = p * rho[..., None]
    y0 = p0 * rho0[..., None]
    return np.concatenate([sphereChart_normal(y, y0), np.log(rho / rho0)[..., None]], -1)


def pointChart_normal_inv(eps, p0):  # :204-215
    rho0 = 1.0 / lg.norm(p0)
    y0 = p0 * rho0[..., None]
    y = sphereChart_normal_inv(eps[..., 0:2], y0)
    rho = rho0 * np.exp(eps[..., 2])
    return y / rho[..., None]


class CoordinateChart:
    """constructVIOChart (VIOState.cpp:217-244)."""

    def __init__(self, sensorChart, sensorChartInv, pointChart, pointChartInv):
        self._s, self._si, self._p, self._pi = sensorChart, sensorChartInv, pointChart, pointChartInv

    def __call__(self, Xi, Xi0):
        eps = np.zeros(SENSOR_DIM + 3 * Xi.N)
        eps[0:SENSOR_DIM] = self._s(Xi.sensor, Xi0.sensor)
        if Xi.N:
            eps[SENSOR_DIM:] = self._p(Xi.p, Xi0.p).reshape(-1)
        return eps

    def inv(self, eps, Xi0):
        sensor = self._si(eps[0:SENSOR_DIM], Xi0.sensor)
        p = self._pi(eps[SENSOR_DIM:].reshape(-1, 3), Xi0.p) if Xi0.N else Xi0.p.copy()
        return VIOState(sensor, p, Xi0.ids.copy())

    chartInv = inv


VIOChart_euclid = CoordinateChart(sensorChart_std, sensorChart_std_inv, pointChart_euclid, pointChart_euclid_inv)
VIOChart_invdepth = CoordinateChart(sensorChart_std, sensorChart_std_inv, pointChart_invdepth,
                                    pointChart_invdepth_inv)
VIOChart_normal = CoordinateChart(sensorChart_normal, sensorChart_normal_inv, pointChart_normal,
                                  pointChart_normal_inv)


def numericalDifferential(f, x, h=-1.0):  # Geometry.cpp:25-36
    if h < 0:
        h = np.cbrt(np.finfo(np.float64).eps)
    x = np.asarray(x, dtype=np.float64)
    f0 = np.asarray(f(x))
    Df = np.zeros((f0.shape[0], x.shape[0]))
    for j in range(x.shape[0]):
        ej = np.zeros(x.shape[0])
        ej[j] = 1.0
        Df[:, j] = (np.asarray(f(x + h * ej)) - np.asarray(f(x - h * ej))) / (2 * h)
    return Df


def _conv_euc2ind(q0):
    """invdepth.cpp:65-73 (also VIOState.cpp:376-385). q0 (N,3) -> (N,3,3)."""
    rho0 = 1.0 / lg.norm(q0)
    y0 = q0 * rho0[:, None]
    M = np.zeros((q0.shape[0], 3, 3))
    proj = np.eye(3) - y0[:, :, None] * y0[:, None, :]
    M[:, 0:2, :] = rho0[:, None, None] * (sphereChart_stereo_diff0(y0) @ proj)
    M[:, 2, :] = -(rho0 * rho0)[:, None] * y0
    return M


def _conv_ind2euc(q0):
    """invdepth.cpp:74-81."""
    rho0 = 1.0 / lg.norm(q0)
    y0 = q0 * rho0[:, None]
    M = np.zeros((q0.shape[0], 3, 3))
    M[:, :, 0:2] = sphereChart_stereo_inv_diff0(y0) / rho0[:, None, None]
    M[:, :, 2] = -y0 / (rho0 * rho0)[:, None]
    return M


def _ind2euc_lift(q0):
    """invdepth.cpp:203-209 / :259-263: [r0 * DPhi^-1(y0), -r0 * q0]."""
    r0 = lg.norm(q0)
    y0 = q0 / r0[:, None]
    M = np.zeros((q0.shape[0], 3, 3))
    M[:, :, 0:2] = r0[:, None, None] * sphereChart_stereo_inv_diff0(y0)
    M[:, :, 2] = -r0[:, None] * q0
    return M


def coordinateDifferential_invdepth_euclid(Xi0):  # VIOState.cpp:355-389
    M = np.eye(Xi0.Dim())
    if Xi0.N:
        Mi = _conv_euc2ind(Xi0.p)
        for i in range(Xi0.N):
            s = SENSOR_DIM + 3 * i
            M[s:s + 3, s:s + 3] = Mi[i]
    return M


def coordinateDifferential_normal_euclid(Xi0):  # VIOState.cpp:391-401
    def coordChange(eps):
        return VIOChart_normal(VIOChart_euclid.inv(eps, Xi0), Xi0)

    return numericalDifferential(coordChange, np.zeros(Xi0.Dim()))


# ----------------------------------------------------------------------------
# Coordinate suites
# ----------------------------------------------------------------------------


def _QhatMatrices(X):
    R_Q = lg.quat_to_matrix(X.Qq)
    return R_Q, R_Q * X.Qa[:, None, None]


def EqFInputMatrixB_euclid(X, xi0, conv=None):  # euclid.cpp:186-233 (conv: invdepth.cpp:123-181)
    N = xi0.N
    Bt = np.zeros((xi0.Dim(), 12))
    xi_hat = stateGroupAction(X, xi0)
    Bt[0:6, 6:12] = np.eye(6)
    R_A = X.A.R
    Bt[6:9, 0:3] = R_A
    Bt[9:12, 0:3] = skew(X.A.x) @ R_A
    Bt[12:15, 0:3] = R_A @ skew(xi_hat.sensor.velocity)
    Bt[12:15, 3:6] = R_A
    if N:
        RT_IC = lg.quat_to_matrix(lg.quat_inv(xi_hat.sensor.cameraOffset.q))
        x_IC = xi_hat.sensor.cameraOffset.x
        _, Qhat = _QhatMatrices(X)
        blocks = Qhat @ (skew(xi_hat.p) @ RT_IC + RT_IC @ skew(x_IC))
        if conv is not None:
            blocks = conv @ blocks
        Bt[SENSOR_DIM:, 0:3] = blocks.reshape(3 * N, 3)
    return Bt


def EqFStateMatrixA_euclid(X, xi0, imuVel, conv=None, convInv=None):  # euclid.cpp:99-160 (invdepth.cpp:36-121)
    N = xi0.N
    dim = xi0.Dim()
    A0t = np.zeros((dim, dim))
    A0t[:, 0:6] = -EqFInputMatrixB_euclid(X, xi0, conv)[:, 0:6]
    A0t[9:12, 12:15] = np.eye(3)
    A0t[12:15, 6:9] = -GRAVITY_CONSTANT * skew(xi0.sensor.gravityDir())
    xi_hat = stateGroupAction(X, xi0)
    v_est = imuVel.minusBias(xi_hat.sensor.inputBias)
    U_I = np.concatenate([v_est.gyr, xi_hat.sensor.velocity])
    adTerm = lg.se3_adjoint(xi0.sensor.cameraOffset.inverse().Adjoint() @ X.A.Adjoint() @ U_I)
    A0t[15:21, 15:21] = adTerm
    if N:
        R_IC = xi_hat.sensor.cameraOffset.R
        R_Ahat = X.A.R
        R_Q, Qhat = _QhatMatrices(X)
        velBlocks = -Qhat @ R_IC.T @ R_Ahat.T
        commonTerm = X.B.inverse().Adjoint() @ adTerm
        temp = np.concatenate([skew(xi0.p) @ R_Q, -X.Qa[:, None, None] * R_Q], -1)
        camBlocks = temp @ commonTerm
        U_C = xi_hat.sensor.cameraOffset.inverse().Adjoint() @ U_I
        v_C = U_C[3:6]
        qhat = xi_hat.p
        inner = (skew(qhat) @ skew(v_C) - 2.0 * v_C[None, :, None] * qhat[:, None, :]
                 + qhat[:, :, None] * v_C[None, None, :])
        A_q = -Qhat @ inner @ np.linalg.inv(Qhat) * (1.0 / np.sum(qhat * qhat, -1))[:, None, None]
        if conv is not None:
            velBlocks = conv @ velBlocks
            camBlocks = conv @ camBlocks
            A_q = conv @ A_q @ convInv
        A0t[SENSOR_DIM:, 12:15] = velBlocks.reshape(3 * N, 3)
        A0t[SENSOR_DIM:, 15:21] = camBlocks.reshape(3 * N, 6)
        for i in range(N):
            s = SENSOR_DIM + 3 * i
            A0t[s:s + 3, s:s + 3] = A_q[i]
    return A0t


def EqFoutputMatrixCiStar_euclid(q0, Qq, Qa, cam, y):  # euclid.cpp:162-184, vectorised over landmarks
    q0 = np.asarray(q0, dtype=np.float64)
    qHat = lg.sot3_apply_inverse(Qq, Qa, q0)
    yHat = lg.normalized(qHat)
    n2 = np.sum(q0 * q0, -1)
    m2g = np.concatenate([-skew(q0), -q0[..., None, :]], -2) / n2[..., None, None]  # (...,4,3)

    def DRho(v):
        DRhoVec = np.concatenate([skew(v), np.zeros(v.shape[:-1] + (3, 1))], -1)  # (...,3,4)
        return cam.projectionJacobian(v) @ DRhoVec  # (...,2,4)

    yTru = cam.undistortPoint(y)
    Rinv = lg.quat_to_matrix(lg.quat_inv(Qq))
    Ad = np.zeros(q0.shape[:-1] + (4, 4))
    Ad[..., 0:3, 0:3] = Rinv
    Ad[..., 3, 3] = 1.0
    return 0.5 * (DRho(yTru) + DRho(yHat)) @ Ad @ m2g


def liftInnovation_euclid(totalInnovation, xi0, ind2euc=None):  # euclid.cpp:36-69 (invdepth.cpp:183-223)
    g = np.asarray(totalInnovation, dtype=np.float64)
    assert g.shape[0] == xi0.Dim()
    D = VIOAlgebra()
    D.u_beta = g[0:6].copy()
    D.U_A = g[6:12].copy()
    D.u_w = -g[12:15] - skew(D.U_A[0:3]) @ xi0.sensor.velocity
    D.U_B = g[15:21] + xi0.sensor.cameraOffset.inverse().Adjoint() @ D.U_A
    N = xi0.N
    W = np.zeros((N, 4))
    if N:
        gq = g[SENSOR_DIM:].reshape(N, 3)
        if ind2euc is not None:
            gq = np.einsum("nij,nj->ni", ind2euc, gq)
        q0 = xi0.p
        n2 = np.sum(q0 * q0, -1)
        W[:, 0:3] = -np.cross(q0, gq) / n2[:, None]
        W[:, 3] = -np.sum(q0 * gq, -1) / n2
    D.W = W
    D.ids = xi0.ids.copy()
    return D


def _liftInnovationDiscrete_common(g, xi0, q1):
    lift = VIOGroup()
    lift.beta = g[0:6].copy()
    lift.A = lg.se3_exp(g[6:12])
    v0 = xi0.sensor.velocity
    lift.w = v0 - lg.quat_rotate(lift.A.q, v0 + g[12:15])
    T0 = xi0.sensor.cameraOffset
    lift.B = T0.inverse() * lift.A * T0 * lg.se3_exp(g[15:21])
    N = xi0.N
    if N:
        q0 = xi0.p
        lift.Qq = lg.quat_from_two_vectors(lg.normalized(q1), lg.normalized(q0))
        lift.Qa = lg.norm(q0) / lg.norm(q1)
    else:
        lift.Qq, lift.Qa = np.zeros((0, 4)), np.zeros(0)
    lift.ids = xi0.ids.copy()
    return lift


def liftInnovationDiscrete_euclid(totalInnovation, xi0):  # euclid.cpp:71-97
    g = np.asarray(totalInnovation, dtype=np.float64)
    q1 = xi0.p + g[SENSOR_DIM:].reshape(-1, 3)
    return _liftInnovationDiscrete_common(g, xi0, q1)


def liftInnovationDiscrete_invdepth(totalInnovation, xi0):  # invdepth.cpp:225-253
    g = np.asarray(totalInnovation, dtype=np.float64)
    q1 = pointChart_invdepth_inv(g[SENSOR_DIM:].reshape(-1, 3), xi0.p) if xi0.N else xi0.p
    return _liftInnovationDiscrete_common(g, xi0, q1)


class EqFCoordinateSuite:
    """EqFMatrices.h:35-67."""

    def __init__(self, name, stateChart):
        self.name = name
        self.stateChart = stateChart

    # -- chart specific pieces, overridden below --
    def stateMatrixA(self, X, xi0, imuVel):
        raise NotImplementedError

    def inputMatrixB(self, X, xi0):
        raise NotImplementedError

    def outputMatrixCiStar(self, q0, Qq, Qa, cam, y):
        raise NotImplementedError

    def liftInnovation(self, g, xi0):
        raise NotImplementedError

    def liftInnovationDiscrete(self, g, xi0):
        raise NotImplementedError

    # -- shared (EqFMatrices.cpp) --
    def outputMatrixCi(self, q0, Qq, Qa, cam):  # EqFMatrices.cpp:84-89
        qHat = lg.sot3_apply_inverse(Qq, Qa, q0)
        yHat = cam.projectPoint(qHat)
        return self.outputMatrixCiStar(q0, Qq, Qa, cam, yHat)

    def outputMatrixC(self, xi0, X, y, useEquivariance=True):  # EqFMatrices.cpp:43-82
        M = xi0.N
        ids, ypx = y.arrays()
        Nmeas = ids.shape[0]
        C = np.zeros((2 * Nmeas, SENSOR_DIM + 3 * M))
        if M == 0 or Nmeas == 0:
            return C
        # column triple i <- state landmark i; row pair j <- j-th measured id (ascending)
        pos = np.searchsorted(ids, xi0.ids)
        pos_c = np.minimum(pos, Nmeas - 1)
        measured = ids[pos_c] == xi0.ids
        idx = np.nonzero(measured)[0]
        if idx.size == 0:
            return C
        rows = pos_c[idx]
        if useEquivariance:
            blocks = self.outputMatrixCiStar(xi0.p[idx], X.Qq[idx], X.Qa[idx], y.cameraPtr, ypx[rows])
        else:
            blocks = self.outputMatrixCi(xi0.p[idx], X.Qq[idx], X.Qa[idx], y.cameraPtr)
        for k, i in enumerate(idx):
            j = rows[k]
            C[2 * j:2 * j + 2, SENSOR_DIM + 3 * i:SENSOR_DIM + 3 * i + 3] = blocks[k]
        return C

    def stateMatrixADiscrete(self, X, xi0, imuVel, dt):  # EqFMatrices.cpp:24-41
        def a0Discrete(eps):
            xi_e = self.stateChart.inv(eps, xi0)
            xi_hat = stateGroupAction(X, xi0)
            xi = stateGroupAction(X, xi_e)
            LambdaTilde = liftVelocityDiscrete(xi, imuVel, dt) * liftVelocityDiscrete(xi_hat, imuVel, dt).inverse()
            xi_e1 = stateGroupAction(X * LambdaTilde * X.inverse(), xi_e)
            return self.stateChart(xi_e1, xi0)

        return numericalDifferential(a0Discrete, np.zeros(xi0.Dim()))


class _SuiteEuclid(EqFCoordinateSuite):
    def stateMatrixA(self, X, xi0, imuVel):
        return EqFStateMatrixA_euclid(X, xi0, imuVel)

    def inputMatrixB(self, X, xi0):
        return EqFInputMatrixB_euclid(X, xi0)

    def outputMatrixCiStar(self, q0, Qq, Qa, cam, y):
        return EqFoutputMatrixCiStar_euclid(q0, Qq, Qa, cam, y)

    def liftInnovation(self, g, xi0):
        return liftInnovation_euclid(g, xi0)

    def liftInnovationDiscrete(self, g, xi0):
        return liftInnovationDiscrete_euclid(g, xi0)


class _SuiteInvDepth(EqFCoordinateSuite):
    def stateMatrixA(self, X, xi0, imuVel):  # invdepth.cpp:36-121
        if xi0.N:
            return EqFStateMatrixA_euclid(X, xi0, imuVel, _conv_euc2ind(xi0.p), _conv_ind2euc(xi0.p))
        return EqFStateMatrixA_euclid(X, xi0, imuVel)

    def inputMatrixB(self, X, xi0):  # invdepth.cpp:123-181
        return EqFInputMatrixB_euclid(X, xi0, _conv_euc2ind(xi0.p) if xi0.N else None)

    def outputMatrixCiStar(self, q0, Qq, Qa, cam, y):  # invdepth.cpp:255-266
        q0 = np.asarray(q0, dtype=np.float64)
        single = q0.ndim == 1
        q0b = q0.reshape(-1, 3)
        C = EqFoutputMatrixCiStar_euclid(q0b, np.reshape(Qq, (-1, 4)), np.reshape(Qa, (-1,)), cam,
                                         np.reshape(y, (-1, 2))) @ _ind2euc_lift(q0b)
        return C[0] if single else C

    def liftInnovation(self, g, xi0):  # invdepth.cpp:183-223
        return liftInnovation_euclid(g, xi0, _ind2euc_lift(xi0.p) if xi0.N else None)

    def liftInnovationDiscrete(self, g, xi0):
        return liftInnovationDiscrete_invdepth(g, xi0)


class _SuiteNormal(EqFCoordinateSuite):
    def stateMatrixA(self, X, xi0, imuVel):  # normal.cpp:37-40
        M = coordinateDifferential_normal_euclid(xi0)
        return M @ EqFStateMatrixA_euclid(X, xi0, imuVel) @ np.linalg.inv(M)

    def inputMatrixB(self, X, xi0):  # normal.cpp:42-45
        return coordinateDifferential_normal_euclid(xi0) @ EqFInputMatrixB_euclid(X, xi0)

    def outputMatrixCiStar(self, q0, Qq, Qa, cam, y):  # normal.cpp:57-65
        q0 = np.asarray(q0, dtype=np.float64)
        single = q0.ndim == 1
        q0b = q0.reshape(-1, 3)
        Qqb = np.reshape(Qq, (-1, 4))
        y0 = lg.normalized(q0b)
        yHat = lg.quat_rotate(lg.quat_inv(Qqb), y0)
        C = np.zeros((q0b.shape[0], 2, 3))
        # NB the reference passes q0 (not y0) as the pole of chartInvDiff0 (normal.cpp:63)
        C[:, :, 0:2] = (cam.projectionJacobian(yHat) @ np.swapaxes(lg.quat_to_matrix(Qqb), -1, -2)
                        @ sphereChart_normal_inv_diff0(q0b))
        return C[0] if single else C

    def liftInnovation(self, g, xi0):  # normal.cpp:47-50
        M = coordinateDifferential_normal_euclid(xi0)
        return liftInnovation_euclid(np.linalg.inv(M) @ g, xi0)

    def liftInnovationDiscrete(self, g, xi0):  # normal.cpp:52-55
        return liftInnovationDiscrete_euclid(VIOChart_euclid(VIOChart_normal.inv(g, xi0), xi0), xi0)


EqFCoordinateSuite_euclid = _SuiteEuclid("Euclidean", VIOChart_euclid)
EqFCoordinateSuite_invdepth = _SuiteInvDepth("InvDepth", VIOChart_invdepth)
EqFCoordinateSuite_normal = _SuiteNormal("Normal", VIOChart_normal)

COORD_EUCLIDEAN, COORD_INVDEPTH, COORD_NORMAL = 0, 1, 2


def getCoordinates(choice):  # EqFMatrices.h:81-90
    return {COORD_EUCLIDEAN: EqFCoordinateSuite_euclid, COORD_INVDEPTH: EqFCoordinateSuite_invdepth,
            COORD_NORMAL: EqFCoordinateSuite_normal, "Euclidean": EqFCoordinateSuite_euclid,
            "InvDepth": EqFCoordinateSuite_invdepth, "Normal": EqFCoordinateSuite_normal}[choice]


# ----------------------------------------------------------------------------
# Settings (VIOFilterSettings.h)
# ----------------------------------------------------------------------------


@dataclass
class Settings:
    biasOmegaProcessVariance: float = 0.001
    biasAccelProcessVariance: float = 0.001
    attitudeProcessVariance: float = 0.001
    positionProcessVariance: float = 0.001
    velocityProcessVariance: float = 0.001
    cameraAttitudeProcessVariance: float = 0.001
    cameraPositionProcessVariance: float = 0.001
    pointProcessVariance: float = 0.001
    velGyrNoise: float = 1e-4
    velAccNoise: float = 1e-3
    velGyrBiasWalk: float = 1e-5
    velAccBiasWalk: float = 1e-3
    measurementNoise: float = 2.0
    outlierThresholdAbs: float = 1e8
    outlierThresholdProb: float = 1e8
    featureRetention: float = 0.3
    initialAttitudeVariance: float = 1.0e-4
    initialPositionVariance: float = 1.0e-4
    initialVelocityVariance: float = 1.0e-2
    initialCameraAttitudeVariance: float = 1.0e-5
    initialCameraPositionVariance: float = 1.0e-4
    initialPointVariance: float = 1.0
    initialPointDepthVariance: float = -1.0
    initialBiasOmegaVariance: float = 0.1
    initialBiasAccelVariance: float = 0.1
    initialSceneDepth: float = 1.0
    useDiscreteInnovationLift: bool = True
    useDiscreteVelocityLift: bool = True
    useDiscreteStateMatrix: bool = False
    fastRiccati: bool = False
    useMedianDepth: bool = True
    useFeaturePredictions: bool = False
    useEquivariantOutput: bool = True
    removeLostLandmarks: bool = True
    coordinateChoice: int = COORD_EUCLIDEAN
    cameraOffset: SE3 = field(default_factory=SE3)

    def constructStateGainMatrix(self, numLandmarks):  # :176-190
        d = np.ones(SENSOR_DIM + 3 * numLandmarks)
        d[0:3] *= self.biasOmegaProcessVariance
        d[3:6] *= self.biasAccelProcessVariance
        d[6:9] *= self.attitudeProcessVariance
        d[9:12] *= self.positionProcessVariance
        d[12:15] *= self.velocityProcessVariance
        d[15:18] *= self.cameraAttitudeProcessVariance
        d[18:21] *= self.cameraPositionProcessVariance
        d[SENSOR_DIM:] *= self.pointProcessVariance
        return np.diag(d)

    def constructInputGainMatrix(self):  # :192-201
        d = np.ones(12)
        d[0:3] *= self.velGyrNoise * self.velGyrNoise
        d[3:6] *= self.velAccNoise * self.velAccNoise
        d[6:9] *= self.velGyrBiasWalk * self.velGyrBiasWalk
        d[9:12] *= self.velAccBiasWalk * self.velAccBiasWalk
        return np.diag(d)

    def constructOutputGainMatrix(self, numLandmarks):  # :203-206
        return self.measurementNoise * self.measurementNoise * np.eye(2 * numLandmarks)

    def constructInitialStateCovariance(self, numLandmarks=0):  # :208-229
        d = np.ones(SENSOR_DIM + 3 * numLandmarks)
        d[0:3] *= self.initialBiasOmegaVariance
        d[3:6] *= self.initialBiasAccelVariance
        d[6:9] *= self.initialAttitudeVariance
        d[9:12] *= self.initialPositionVariance
        d[12:15] *= self.initialVelocityVariance
        d[15:18] *= self.initialCameraAttitudeVariance
        d[18:21] *= self.initialCameraPositionVariance
        d[SENSOR_DIM:] *= self.initialPointVariance
        if self.initialPointDepthVariance > 0:
            d[SENSOR_DIM + 2::3] = self.initialPointDepthVariance
        return np.diag(d)


# ----------------------------------------------------------------------------
# VIO_eqf (src/mathematical/VIO_eqf.cpp)
# ----------------------------------------------------------------------------


class VIO_eqf:
    def __init__(self, coordinateSuite=None, xi0=None, X=None, Sigma=None):
        self.coordinateSuite = EqFCoordinateSuite_euclid if coordinateSuite is None else coordinateSuite
        self.xi0 = VIOState() if xi0 is None else xi0
        self.X = VIOGroup.Identity() if X is None else X
        self.Sigma = np.eye(SENSOR_DIM) if Sigma is None else np.array(Sigma, dtype=np.float64)
        self.currentTime = -1.0
        # when True, S^-1 and K are evaluated twice like the reference's lazy
        # Eigen expressions (VIO_eqf.cpp:116-131); results are identical, only
        # the cost changes.  Used by the CPU-baseline timing.
        self.mirrorLazyEvaluation = False
        # when True, propagation keeps A sparse and the correction uses sparse C + one Cholesky of S (minimal-flop
        # symmetric form, F_alg of SURVEY 8d).  Same result to rounding; the "algorithmic" CPU baseline of bench.py.
        self.structuredEvaluation = False

    def stateEstimate(self):  # :137
        return stateGroupAction(self.X, self.xi0)

    def integrateObserverState(self, imuVelocity, dt, discreteLift=True):  # :47-60
        if discreteLift:
            lifted = liftVelocityDiscrete(self.stateEstimate(), imuVelocity, dt)
        else:
            lifted = VIOExp(liftVelocity(self.stateEstimate(), imuVelocity) * dt)
        self.X = self.X * lifted

    def integrateRiccatiStateFast(self, imuVelocity, dt, inputGainMatrix, stateGainMatrix):  # :62-72
        A0t = self.coordinateSuite.stateMatrixA(self.X, self.xi0, imuVelocity)
        Bt = self.coordinateSuite.inputMatrixB(self.X, self.xi0)
        if self.structuredEvaluation:
            # CPU algorithmic baseline (bench.py): the same step using the block structure of A: 21 dense
            # sensor columns + one 3x3 diagonal block per landmark -- O(dim^2) instead of two dense GEMMs
            N = self.xi0.N
            dim = self.xi0.Dim()
            As = A0t[:, :SENSOR_DIM]
            ar = np.arange(N)
            D = A0t[SENSOR_DIM:, SENSOR_DIM:].reshape(N, 3, N, 3)[ar, :, ar, :]  # (N, 3, 3) diagonal blocks
            S0 = self.Sigma
            FS = S0 + dt * (As @ S0[:SENSOR_DIM, :])
            if N:
                FS[SENSOR_DIM:, :] += np.matmul(dt * D, S0[SENSOR_DIM:, :].reshape(N, 3, dim)).reshape(3 * N, dim)
            out = FS.T + dt * (As @ FS[:, :SENSOR_DIM].T)  # transposed result: (FS F^T)^T = F FS^T
            if N:
                out[SENSOR_DIM:, :] += np.matmul(dt * D, FS.T[SENSOR_DIM:, :].reshape(N, 3, dim)).reshape(3 * N, dim)
                out = out.T
            self.Sigma = out + dt * (Bt @ inputGainMatrix @ Bt.T + stateGainMatrix)
            return
        A0tExp = np.eye(self.xi0.Dim()) + dt * A0t
        self.Sigma = A0tExp @ self.Sigma @ A0tExp.T + dt * (Bt @ inputGainMatrix @ Bt.T + stateGainMatrix)

    def integrateRiccatiStateAccurate(self, imuVelocity, dt, inputGainMatrix, stateGainMatrix):  # :74-91
        from scipy.linalg import expm

        A0t = self.coordinateSuite.stateMatrixA(self.X, self.xi0, imuVelocity)
        Bt = self.coordinateSuite.inputMatrixB(self.X, self.xi0)
        n, m = A0t.shape[0], Bt.shape[1]
        AB = np.zeros((n + m, n + m))
        AB[0:n, 0:n] = A0t
        AB[0:n, n:n + m] = Bt
        ABExp = expm(dt * AB)
        A0tExp = ABExp[0:n, 0:n]
        BtExp = ABExp[0:n, n:n + m]
        self.Sigma = A0tExp @ self.Sigma @ A0tExp.T + BtExp @ (inputGainMatrix / dt) @ BtExp.T + dt * stateGainMatrix

    def integrateRiccatiStateDiscrete(self, imuVelocity, dt, inputGainMatrix, stateGainMatrix):  # :93-103
        Bt = self.coordinateSuite.inputMatrixB(self.X, self.xi0)
        Ad = self.coordinateSuite.stateMatrixADiscrete(self.X, self.xi0, imuVelocity, dt)
        self.Sigma = Ad @ self.Sigma @ Ad.T + dt * (Bt @ inputGainMatrix @ Bt.T + stateGainMatrix)

    def performVisionUpdate(self, measurement, outputGainMatrix, useEquivariantOutput=True,
                            discreteCorrection=False):  # :105-135
        if not measurement.camCoordinates:
            return
        estimated = measureSystemState(self.stateEstimate(), measurement.cameraPtr)
        yTilde = (measurement - estimated).asVector()
        Ct = self.coordinateSuite.outputMatrixC(self.xi0, self.X, measurement, useEquivariantOutput)
        Sigma = self.Sigma
        if self.structuredEvaluation:
            # CPU algorithmic baseline (bench.py): sparse C, Cholesky of S, Sigma -= Y^T Y -- the minimal-flop symmetric form
            # (F_alg of SURVEY 8d) instead of the reference's dense, doubly evaluated gain.  Same result to rounding.
            import scipy.linalg as sl

            N = self.xi0.N
            dim = self.xi0.Dim()
            m = Ct.shape[0]
            n = m // 2
            # one 2x3 block per row pair, at the column triple of its landmark (EqFMatrices.cpp:74-76)
            C4 = Ct[:, SENSOR_DIM:].reshape(n, 2, N, 3)
            lm = np.abs(C4).sum(axis=(1, 3)).argmax(axis=1)
            Cb = C4[np.arange(n), :, lm, :]  # (n, 2, 3)
            W = np.matmul(Cb, Sigma[SENSOR_DIM:, :].reshape(N, 3, dim)[lm]).reshape(m, dim)
            Wl = W[:, SENSOR_DIM:].reshape(m, N, 3)[:, lm, :].transpose(1, 2, 0)  # (n, 3, m)
            S = np.matmul(Cb, Wl).reshape(m, m).T + outputGainMatrix
            L = sl.cholesky(S, lower=True, check_finite=False)
            Y = sl.solve_triangular(L, W, lower=True, check_finite=False)
            z = sl.solve_triangular(L, yTilde, lower=True, check_finite=False)
            Gamma = Y.T @ z
            if discreteCorrection:
                Delta = self.coordinateSuite.liftInnovationDiscrete(Gamma, self.xi0)
            else:
                Delta = VIOExp(self.coordinateSuite.liftInnovation(Gamma, self.xi0))
            self.X = Delta * self.X
            T = sl.blas.dsyrk(1.0, Y, trans=1, lower=1)  # lower triangle of Y^T Y (half the flops), zeros above the diagonal
            out = Sigma - T
            out -= T.T
            out[np.diag_indices_from(out)] += np.diagonal(T)
            self.Sigma = out
            self.lastGamma = Gamma
            return

        def SInv():
            return np.linalg.inv(Ct @ Sigma @ Ct.T + outputGainMatrix)

        def K():
            return Sigma @ Ct.T @ SInv()

        Gamma = K() @ yTilde
        if discreteCorrection:
            Delta = self.coordinateSuite.liftInnovationDiscrete(Gamma, self.xi0)
        else:
            Delta = VIOExp(self.coordinateSuite.liftInnovation(Gamma, self.xi0))
        self.X = Delta * self.X
        if self.mirrorLazyEvaluation:
            self.Sigma = Sigma - (K() @ Ct) @ Sigma
        else:
            # same value: K = Sigma C^T S^-1 evaluated once
            Kc = K()
            self.Sigma = Sigma - (Kc @ Ct) @ Sigma
        self.lastGamma = Gamma

    def predictState(self, stamp, imuVelocities):  # :139-151
        pred = self.stateEstimate()
        n = len(imuVelocities)
        for i in range(n):
            t0 = max(imuVelocities[i].stamp, self.currentTime)
            t1 = min(imuVelocities[i + 1].stamp, stamp) if i + 1 < n else stamp
            dt = max(t1 - t0, 0.0)
            pred = integrateSystemFunction(pred, imuVelocities[i], dt)
        return pred

    def computeNEES(self, trueState):  # :153-170
        idx = [int(np.nonzero(trueState.ids == i)[0][0]) for i in self.X.ids]
        trunc = VIOState(trueState.sensor.copy(), trueState.p[idx], trueState.ids[idx])
        stateError = stateGroupAction(self.X.inverse(), trunc)
        eps = self.coordinateSuite.stateChart(stateError, self.xi0)
        info = np.linalg.inv(self.Sigma)
        return float(eps @ info @ eps) / trunc.Dim()

    def removeLandmarkByIndex(self, idx):  # :172-178 (+ removeRows/Cols :27-45)
        self.xi0.p = np.delete(self.xi0.p, idx, 0)
        self.xi0.ids = np.delete(self.xi0.ids, idx, 0)
        self.X.ids = np.delete(self.X.ids, idx, 0)
        self.X.Qq = np.delete(self.X.Qq, idx, 0)
        self.X.Qa = np.delete(self.X.Qa, idx, 0)
        s = SENSOR_DIM + 3 * idx
        self.Sigma = np.delete(np.delete(self.Sigma, [s, s + 1, s + 2], 0), [s, s + 1, s + 2], 1)

    def removeLandmarksByIndices(self, indices):
        """Batch form of repeated removeLandmarkByIndex (order preserving erase)."""
        indices = sorted(set(int(i) for i in indices))
        if not indices:
            return
        self.xi0.p = np.delete(self.xi0.p, indices, 0)
        self.xi0.ids = np.delete(self.xi0.ids, indices, 0)
        self.X.ids = np.delete(self.X.ids, indices, 0)
        self.X.Qq = np.delete(self.X.Qq, indices, 0)
        self.X.Qa = np.delete(self.X.Qa, indices, 0)
        rows = np.concatenate([SENSOR_DIM + 3 * np.array(indices)[:, None] + np.arange(3)[None, :]]).reshape(-1)
        self.Sigma = np.delete(np.delete(self.Sigma, rows, 0), rows, 1)

    def removeLandmarkById(self, idn):  # :180-186
        idx = int(np.nonzero(self.xi0.ids == idn)[0][0])
        self.removeLandmarkByIndex(idx)

    def getLandmarkCovById(self, idn):  # :188-194
        i = int(np.nonzero(self.xi0.ids == idn)[0][0])
        s = SENSOR_DIM + 3 * i
        return self.Sigma[s:s + 3, s:s + 3]

    def getOutputCovById(self, idn, y, camPtr):  # :196-211 (y unused, non-star C_i)
        i = int(np.nonzero(self.xi0.ids == idn)[0][0])
        lmCov = self.getLandmarkCovById(idn)
        C0i = self.coordinateSuite.outputMatrixCi(self.xi0.p[i], self.X.Qq[i], self.X.Qa[i], camPtr)
        return C0i @ lmCov @ C0i.T

    def removeInvalidLandmarks(self):  # :213-223
        bad = np.nonzero((self.X.Qa <= 1e-8) | (self.X.Qa > 1e8))[0]
        self.removeLandmarksByIndices(bad)

    def addNewLandmarks(self, newP, newIds, newLandmarkCov):  # :225-245
        newP = np.asarray(newP, dtype=np.float64).reshape(-1, 3)
        newIds = np.asarray(newIds, dtype=np.int64).reshape(-1)
        n = newIds.shape[0]
        self.xi0.p = np.concatenate([self.xi0.p, newP], 0)
        self.xi0.ids = np.concatenate([self.xi0.ids, newIds])
        self.X.ids = np.concatenate([self.X.ids, newIds])
        self.X.Qq = np.concatenate([self.X.Qq.reshape(-1, 4), np.tile(lg.QUAT_IDENTITY, (n, 1))], 0)
        self.X.Qa = np.concatenate([self.X.Qa, np.ones(n)])
        og = self.Sigma.shape[0]
        S = np.zeros((og + 3 * n, og + 3 * n))
        S[0:og, 0:og] = self.Sigma
        S[og:, og:] = newLandmarkCov
        self.Sigma = S


# ----------------------------------------------------------------------------
# VIOFilter (src/VIOFilter.cpp)
# ----------------------------------------------------------------------------


class VIOFilter:
    def __init__(self, settings, xi0=None, time=0.0):
        self.settings = copy.deepcopy(settings)
        self.filterState = VIO_eqf()
        self.velocityBuffer = []
        self.initialisedFlag = False
        fs = self.filterState
        if xi0 is None:  # ctor #1, VIOFilter.cpp:31-41
            fs.Sigma = self.settings.constructInitialStateCovariance()
            fs.xi0.sensor.cameraOffset = self.settings.cameraOffset.copy()
        else:  # ctor #2, VIOFilter.cpp:43-56
            fs.Sigma = self.settings.constructInitialStateCovariance(xi0.N)
            fs.xi0 = xi0.copy()
            fs.X = VIOGroup.Identity(xi0.ids)
            fs.currentTime = float(time)
            self.initialisedFlag = True
        fs.coordinateSuite = getCoordinates(self.settings.coordinateChoice)
        self.timing = {"propagation": 0.0, "preprocessing": 0.0, "correction": 0.0}

    # -- input ---------------------------------------------------------------
    def processIMUData(self, imuVelocity):  # :58-63
        if not self.initialisedFlag:
            self.initialiseFromIMUData(imuVelocity)
        self.velocityBuffer.append(imuVelocity)

    def initialiseFromIMUData(self, imuVelocity):  # :65-78
        s = self.filterState.xi0.sensor
        s.inputBias = np.zeros(6)
        s.pose = SE3()
        s.velocity = np.zeros(3)
        self.initialisedFlag = True
        s.pose.q = lg.quat_from_two_vectors(lg.normalized(imuVelocity.acc), E3)
        self.filterState.currentTime = imuVelocity.stamp

    def setState(self, xi):  # :80-92
        fs = self.filterState
        fs.xi0 = xi.copy()
        fs.X = VIOGroup.Identity(xi.ids)
        N = xi.N
        Sig = np.eye(SENSOR_DIM + 3 * N)
        Sig[0:SENSOR_DIM, 0:SENSOR_DIM] = self.settings.constructInitialStateCovariance()
        Sig[SENSOR_DIM:, SENSOR_DIM:] *= self.settings.initialPointVariance
        fs.Sigma = Sig
        self.initialisedFlag = True

    def setLandmarks(self, p, ids):  # :94-110
        fs = self.filterState
        p = np.asarray(p, dtype=np.float64).reshape(-1, 3)
        n = p.shape[0]
        Full = self.settings.constructInitialStateCovariance(n)
        fs.Sigma[SENSOR_DIM:SENSOR_DIM + 3 * n, SENSOR_DIM:SENSOR_DIM + 3 * n] = Full[SENSOR_DIM:, SENSOR_DIM:]
        fs.xi0.p = p.copy()
        fs.xi0.ids = np.array(ids, dtype=np.int64)
        fs.X.ids = fs.xi0.ids.copy()
        fs.X.Qq = np.tile(lg.QUAT_IDENTITY, (n, 1))
        fs.X.Qa = np.ones(n)

    def augmentLandmarkStates(self, newIds, providedState):  # :112-132
        self.removeOldLandmarks(newIds)
        have = set(int(i) for i in self.filterState.X.ids)
        prov = {int(i): k for k, i in enumerate(providedState.ids)}
        add = [int(i) for i in newIds if int(i) not in have]
        newP = np.array([providedState.p[prov[i]] for i in add]).reshape(-1, 3)
        cov = np.eye(3 * len(add)) * self.settings.initialPointVariance
        self.filterState.addNewLandmarks(newP, add, cov)

    # -- propagation -----------------------------------------------------------
    def integrateUpToTime(self, newTime):  # :134-192
        fs = self.filterState
        st = self.settings
        buf = self.velocityBuffer
        if newTime <= fs.currentTime or fs.currentTime < 0 or not buf:
            return False
        n = len(buf)

        def seg(i):
            t0 = max(buf[i].stamp, fs.currentTime)
            t1 = min(buf[i + 1].stamp, newTime) if i + 1 < n else newTime
            return max(t1 - t0, 0.0)

        if st.fastRiccati:
            accT = 0.0
            accV = IMUVelocity.Zero()
            for i in range(n):
                dt = seg(i)
                accT += dt
                accV = accV + buf[i] * dt
            accV = accV * (1.0 / accT)
            fs.integrateRiccatiStateFast(accV, accT, st.constructInputGainMatrix(),
                                         st.constructStateGainMatrix(fs.xi0.N))
        for i in range(n):
            dt = seg(i)
            if not st.fastRiccati and dt > 0:
                if st.useDiscreteStateMatrix:
                    fs.integrateRiccatiStateDiscrete(buf[i], dt, st.constructInputGainMatrix(),
                                                     st.constructStateGainMatrix(fs.xi0.N))
                else:
                    fs.integrateRiccatiStateAccurate(buf[i], dt, st.constructInputGainMatrix(),
                                                     st.constructStateGainMatrix(fs.xi0.N))
            fs.integrateObserverState(buf[i], dt, st.useDiscreteVelocityLift)
        fs.currentTime = newTime
        # prune, keeping the last sample with stamp < currentTime (:183-189)
        k = next((j for j, u in enumerate(buf) if u.stamp >= fs.currentTime), n)
        if k != 0:
            del buf[0:k - 1]
        return True

    # -- vision ----------------------------------------------------------------
    def processVisionData(self, measurement):  # :194-241
        import time as _t

        t0 = _t.perf_counter()
        ok = self.integrateUpToTime(measurement.stamp)
        if not ok or not self.initialisedFlag:
            return
        t1 = _t.perf_counter()
        self.timing["propagation"] += t1 - t0
        if self.settings.removeLostLandmarks:
            self.removeOldLandmarks(measurement.getIds())
        matched = measurement.copy()
        self.removeOutliers(matched)
        self.addNewLandmarks(matched)
        t2 = _t.perf_counter()
        self.timing["preprocessing"] += t2 - t1
        if not matched.camCoordinates:
            return
        self.filterState.performVisionUpdate(
            matched, self.settings.constructOutputGainMatrix(len(matched.camCoordinates)),
            self.settings.useEquivariantOutput, self.settings.useDiscreteInnovationLift)
        self.filterState.removeInvalidLandmarks()
        self.timing["correction"] += _t.perf_counter() - t2

    def stateEstimate(self):  # :243
        return self.filterState.stateEstimate()

    def viewEqFState(self):  # :245
        return self.filterState

    def getFeaturePredictions(self, camPtr, stamp=-1.0):  # :247-252
        if self.settings.useFeaturePredictions:
            return measureSystemState(self.filterState.predictState(stamp, self.velocityBuffer), camPtr)
        return VisionMeasurement()

    def getTime(self):  # :256
        return self.filterState.currentTime

    def isInitialised(self):
        return self.initialisedFlag

    def addNewLandmarks(self, measurement):  # :258-278
        fs = self.filterState
        have = set(int(i) for i in fs.X.ids)
        newIds = [i for i in measurement.getIds() if i not in have]
        if not newIds:
            return
        ypx = np.array([measurement.camCoordinates[i] for i in newIds])
        bearings = measurement.cameraPtr.undistortPoint(ypx)
        depth = self.getMedianSceneDepth() if self.settings.useMedianDepth else self.settings.initialSceneDepth
        newP = bearings * depth
        cov = np.eye(3 * len(newIds)) * self.settings.initialPointVariance
        fs.addNewLandmarks(newP, newIds, cov)

    def removeOldLandmarks(self, measurementIds):  # :280-302
        fs = self.filterState
        if fs.X.ids.shape[0] == 0:
            return
        keep = np.isin(fs.X.ids, np.array(list(measurementIds), dtype=np.int64))
        fs.removeLandmarksByIndices(np.nonzero(~keep)[0])

    def removeOutliers(self, measurement):  # :304-364
        st = self.settings
        fs = self.filterState
        maxOutliers = int((1.0 - st.featureRetention) * len(measurement.camCoordinates))
        xiHat = self.stateEstimate()
        yHat = measureSystemState(xiHat, measurement.cameraPtr)
        proposed = []
        absoluteOutliers = {}
        for lmId in yHat.getIds():
            if lmId not in measurement.camCoordinates:
                continue
            d = measurement.camCoordinates[lmId] - yHat.camCoordinates[lmId]
            err = float(np.sqrt(d @ d))
            if err > st.outlierThresholdAbs:
                absoluteOutliers[lmId] = err
                proposed.append(lmId)
        probabilisticOutliers = {}
        residual = measurement - yHat
        cand = [lmId for lmId in residual.getIds() if lmId not in absoluteOutliers]
        if cand:
            # getOutputCovById for every candidate at once (VIO_eqf.cpp:196-211): C0i Sigma_ii C0i^T with the
            # non-star C0i; same arithmetic as the per-id calls, vectorised over landmarks
            order = {int(i): k for k, i in enumerate(fs.xi0.ids)}
            idx = np.array([order[lmId] for lmId in cand], dtype=np.int64)
            C0 = fs.coordinateSuite.outputMatrixCi(fs.xi0.p[idx], fs.X.Qq[idx], fs.X.Qa[idx], measurement.cameraPtr)
            s0 = SENSOR_DIM + 3 * idx
            rows = s0[:, None] + np.arange(3)[None, :]
            lmCov = fs.Sigma[rows[:, :, None], rows[:, None, :]]
            covs = C0 @ lmCov @ np.swapaxes(C0, -1, -2)
            for k, lmId in enumerate(cand):
                yT = residual.camCoordinates[lmId]
                cov = covs[k]
                det = cov[0, 0] * cov[1, 1] - cov[0, 1] * cov[1, 0]
                inv = np.array([[cov[1, 1], -cov[0, 1]], [-cov[1, 0], cov[0, 0]]]) / det
                err = float(yT @ inv @ yT)
                if err > st.outlierThresholdProb:
                    probabilisticOutliers[lmId] = err
                    proposed.append(lmId)

        # absolute outliers outrank probabilistic ones; larger error first (:338-358)
        def key(lmId):
            if lmId in absoluteOutliers:
                return (1, absoluteOutliers[lmId])
            return (0, probabilisticOutliers[lmId])

        proposed.sort(key=key, reverse=True)
        if len(proposed) > maxOutliers:
            proposed = proposed[:maxOutliers]
        for lmId in proposed:
            fs.removeLandmarkById(lmId)
            del measurement.camCoordinates[lmId]
        self.lastOutliers = list(proposed)

    def getMedianSceneDepth(self):  # :366-380
        p = self.stateEstimate().p
        if p.shape[0] == 0:
            return self.settings.initialSceneDepth
        d2 = np.sum(p * p, -1)
        mid = d2.shape[0] // 2
        return float(np.sqrt(np.partition(d2, mid)[mid]))
