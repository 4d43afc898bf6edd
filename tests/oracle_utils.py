"""Random-input helpers mirroring test/testing_utilities.cpp of the reference
(reasonableStateElement :24-44, randomStateElement :46-65, randomVelocityElement
:82-90, randomGroupElement :92-108, reasonableGroupElement :110-124, logNorm
:126-135, stateDistance :137-150, randomVisionMeasurement :152-165,
measurementDistance :167-173, assertMatrixEquality :186-213)."""

import numpy as np

from oracle import eqf, liegroups as lg
from oracle.camera import createDefaultCamera
from oracle.liegroups import SE3

TEST_REPS = 25
NEAR_ZERO = 1e-12


def _rand(rng, *shape):
    return rng.uniform(-1.0, 1.0, shape)


def _unit_random_quat(rng):
    q = rng.standard_normal(4)
    return q / np.linalg.norm(q)


def randomStateElement(rng, ids, reasonable=False):
    xi = eqf.VIOState()
    xi.sensor.inputBias = _rand(rng, 6)
    xi.sensor.pose = SE3(_unit_random_quat(rng), _rand(rng, 3))
    xi.sensor.cameraOffset = SE3(_unit_random_quat(rng), _rand(rng, 3))
    xi.sensor.velocity = _rand(rng, 3)
    n = len(ids)
    xi.p = _rand(rng, n, 3) * 10
    if reasonable:
        xi.p[:, 2] += 20.0
    xi.ids = np.array(ids, dtype=np.int64)
    return xi


def reasonableStateElement(rng, ids):
    return randomStateElement(rng, ids, True)


def randomVelocityElement(rng):
    return eqf.IMUVelocity(0.0, _rand(rng, 3), _rand(rng, 3), _rand(rng, 3), _rand(rng, 3))


def randomGroupElement(rng, ids):
    X = eqf.VIOGroup()
    X.beta = _rand(rng, 6)
    X.A = SE3(_unit_random_quat(rng), _rand(rng, 3))
    X.B = SE3(_unit_random_quat(rng), _rand(rng, 3))
    X.w = _rand(rng, 3)
    X.ids = np.array(ids, dtype=np.int64)
    X.Qq = np.array([_unit_random_quat(rng) for _ in ids]).reshape(-1, 4)
    X.Qa = 2.0 * rng.uniform(0.0, 1.0, len(ids)) + 1.0
    return X


def reasonableGroupElement(rng, ids):
    X = eqf.VIOGroup()
    X.beta = _rand(rng, 6) * 0.1
    X.A = lg.se3_exp(_rand(rng, 6) * 0.1)
    X.B = lg.se3_exp(_rand(rng, 6) * 0.1)
    X.w = _rand(rng, 3) * 0.1
    X.ids = np.array(ids, dtype=np.int64)
    X.Qq = lg.so3_exp(_rand(rng, len(ids), 3) * 0.02)
    X.Qa = 2.0 * rng.uniform(0.0, 1.0, len(ids)) + 1.0
    return X


def logNorm(X):
    r = np.linalg.norm(lg.se3_log(X.A)) + np.linalg.norm(lg.se3_log(X.B)) + np.linalg.norm(X.w)
    if X.Qq.shape[0]:
        r += np.sum(np.linalg.norm(lg.sot3_log(X.Qq, X.Qa), axis=-1))
    return r


def stateDistance(xi1, xi2):
    d = np.linalg.norm(xi1.sensor.inputBias - xi2.sensor.inputBias)
    d += np.linalg.norm(lg.se3_log(xi1.sensor.pose.inverse() * xi2.sensor.pose))
    d += np.linalg.norm(lg.se3_log(xi1.sensor.cameraOffset.inverse() * xi2.sensor.cameraOffset))
    d += np.linalg.norm(xi1.sensor.velocity - xi2.sensor.velocity)
    assert np.array_equal(xi1.ids, xi2.ids)
    d += np.sum(np.linalg.norm(xi1.p - xi2.p, axis=-1))
    return d


def randomVisionMeasurement(rng, ids):
    cam = createDefaultCamera()
    out = {}
    for i in ids:
        while True:
            p = _rand(rng, 3)
            p = p / np.linalg.norm(p)
            if p[2] >= 1e-1:
                break
        out[int(i)] = cam.projectPoint(p)
    return eqf.VisionMeasurement(0.0, out, cam)


def measurementDistance(y1, y2):
    scale = max(np.linalg.norm(y1.asVector()), np.linalg.norm(y2.asVector()))
    return np.linalg.norm((y1 - y2).asVector()) / scale


def assertMatrixEquality(M1, M2, h=-1.0):
    if h < 0:
        h = np.cbrt(np.finfo(np.float64).eps)
    M1 = np.atleast_2d(M1)
    M2 = np.atleast_2d(M2)
    assert M1.shape == M2.shape
    assert not np.isnan(M1).any() and not np.isnan(M2).any()
    tol = np.maximum(h, h * 1e1 * np.abs(M1))
    bad = np.abs(M1 - M2) > tol
    assert not bad.any(), f"{bad.sum()} entries differ; max err {np.abs(M1 - M2).max()}"


def testDifferential(f, x, Df, h=-1.0):
    """test/testing_utilities.h:45-53."""
    assertMatrixEquality(Df, eqf.numericalDifferential(f, x, h), h)


testDifferential.__test__ = False
