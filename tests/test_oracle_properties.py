"""Pins the oracle with the reference's own property tests, re-expressed in
pytest.  Each test names the gtest it ports.  These are the only tests the
reference holds for the hot path: none of them pins a
numeric golden value, all are self-consistency properties."""

import numpy as np
import pytest

from oracle import eqf, liegroups as lg
from oracle.camera import EquidistantCamera, PinholeCamera, StandardCamera, createDefaultCamera

from oracle_utils import (NEAR_ZERO, TEST_REPS, assertMatrixEquality, logNorm, measurementDistance,
                          randomGroupElement, randomStateElement, randomVelocityElement, randomVisionMeasurement,
                          reasonableGroupElement, reasonableStateElement, stateDistance, testDifferential)

SUITES = [eqf.EqFCoordinateSuite_euclid, eqf.EqFCoordinateSuite_invdepth, eqf.EqFCoordinateSuite_normal]
IDS5 = [0, 1, 2, 3, 4]


# --------------------------------------------------------------- LiePP (external/LiePP/test/test_groups.cpp:80-216)
def test_liepp_exp_log_roundtrip():
    rng = np.random.default_rng(0)
    for _ in range(TEST_REPS):
        w = rng.uniform(-1, 1, 3)
        assert np.linalg.norm(lg.so3_log(lg.so3_exp(w)) - w) < 1e-8
        u = rng.uniform(-1, 1, 6)
        assert np.linalg.norm(lg.se3_log(lg.se3_exp(u)) - u) < 1e-8
        u9 = rng.uniform(-1, 1, 9)
        assert np.linalg.norm(lg.se23_log(*lg.se23_exp(u9)) - u9) < 1e-8
        W = rng.uniform(-1, 1, 4)
        assert np.linalg.norm(lg.sot3_log(*lg.sot3_exp(W)) - W) < 1e-8


def test_liepp_exp_matches_matrix_exponential():
    from scipy.linalg import expm

    rng = np.random.default_rng(1)
    for _ in range(TEST_REPS):
        u = rng.uniform(-1, 1, 6)
        U = np.zeros((4, 4))
        U[0:3, 0:3] = lg.skew(u[0:3])
        U[0:3, 3] = u[3:6]
        assert np.abs(lg.se3_exp(u).asMatrix() - expm(U)).max() < 1e-8


def test_liepp_adjoint_identities():
    rng = np.random.default_rng(2)
    for _ in range(TEST_REPS):
        P = lg.se3_exp(rng.uniform(-1, 1, 6))
        u = rng.uniform(-1, 1, 6)
        # Ad_P u = vee(P u^ P^-1)
        U = np.zeros((4, 4))
        U[0:3, 0:3] = lg.skew(u[0:3])
        U[0:3, 3] = u[3:6]
        M = P.asMatrix() @ U @ np.linalg.inv(P.asMatrix())
        v = np.concatenate([lg.vex(M[0:3, 0:3]), M[0:3, 3]])
        assert np.abs(P.Adjoint() @ u - v).max() < 1e-8
        # ad_u v = [u, v]
        w = rng.uniform(-1, 1, 6)
        W = np.zeros((4, 4))
        W[0:3, 0:3] = lg.skew(w[0:3])
        W[0:3, 3] = w[3:6]
        C = U @ W - W @ U
        assert np.abs(lg.se3_adjoint(u) @ w - np.concatenate([lg.vex(C[0:3, 0:3]), C[0:3, 3]])).max() < 1e-8


def test_quat_from_two_vectors_rotates_origin_to_dest():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((TEST_REPS, 3))
    b = rng.standard_normal((TEST_REPS, 3))
    q = lg.quat_from_two_vectors(a, b)
    assert np.abs(lg.quat_rotate(q, lg.normalized(a)) - lg.normalized(b)).max() < 1e-12
    # antiparallel branch (c < -1 + 1e-12) still yields a valid half-turn
    q = lg.quat_from_two_vectors(np.array([[0.0, 0.0, 1.0]]), np.array([[0.0, 0.0, -1.0]]))
    assert np.abs(lg.quat_rotate(q, np.array([[0.0, 0.0, 1.0]])) - np.array([[0.0, 0.0, -1.0]])).max() < 1e-7


def test_matrix_quaternion_roundtrip_all_branches():
    rng = np.random.default_rng(4)
    for _ in range(4 * TEST_REPS):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        R = lg.quat_to_matrix(q)
        q2 = lg.matrix_to_quat(R)
        assert min(np.abs(q - q2).max(), np.abs(q + q2).max()) < 1e-12


# --------------------------------------------------------------- GIFT (external/GIFT/test/test_Camera.cpp:62-113)
def test_camera_pinhole_project_grid():
    cam = PinholeCamera(752, 480, 458.654, 457.296, 367.215, 248.375)
    for x in range(0, 752, 30):
        for y in range(0, 480, 30):
            b = cam.undistortPoint(np.array([x, y], dtype=np.float64))
            est = b[0:2] / b[2]
            assert np.linalg.norm(est - np.array([(x - cam.cx) / cam.fx, (y - cam.cy) / cam.fy])) <= 1e-4


def test_camera_standard_projection_jacobian():
    cam = StandardCamera(752, 480, 458.654, 457.296, 367.215, 248.375,
                         [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 1e-4])
    for x in range(0, 752, 30):
        for y in range(0, 480, 30):
            sp = cam.undistortPoint(np.array([x, y], dtype=np.float64))
            testDifferential(cam.projectPoint, sp, cam.projectionJacobian(sp))


def test_camera_standard_reprojection():
    # the LSQ inverse-distortion fit reprojects grid pixels to within a pixel
    cam = StandardCamera(752, 480, 458.654, 457.296, 367.215, 248.375,
                         [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
    for x in range(60, 700, 60):
        for y in range(60, 420, 60):
            px = np.array([x, y], dtype=np.float64)
            assert np.linalg.norm(cam.projectPoint(cam.undistortPoint(px)) - px) < 1.0


# external/GIFT/GIFT/test/test_Camera.cpp (Equidistant cases): Jacobian vs finite differences and
# undistort -> project round trip; the Gauss-Newton inverse stops at a 0.1 px residual by design
EQUI_DIST = [-0.013721808247486035, 0.020727425669427896, -0.012786476702685545, 0.0025242267320687625]


def test_camera_equidistant_projection_jacobian():
    cam = EquidistantCamera(512, 512, 190.978, 190.973, 254.93, 256.9, EQUI_DIST)
    for x in range(20, 512, 41):
        for y in range(20, 512, 41):
            if np.hypot((x - cam.cx) / cam.fx, (y - cam.cy) / cam.fy) > 1.3:
                continue
            sp = cam.undistortPoint(np.array([x, y], dtype=np.float64))
            testDifferential(cam.projectPoint, sp, cam.projectionJacobian(sp))
    # the r <= 1e-6 branch: optical axis
    J = cam.projectionJacobian(np.array([0.0, 0.0, 2.0]))
    assert np.allclose(J, np.array([[cam.fx / 2, 0, 0], [0, cam.fy / 2, 0]]))


def test_camera_equidistant_reprojection():
    cam = EquidistantCamera(512, 512, 190.978, 190.973, 254.93, 256.9, EQUI_DIST)
    for x in range(20, 512, 41):
        for y in range(20, 512, 41):
            px = np.array([x, y], dtype=np.float64)
            if np.hypot((x - cam.cx) / cam.fx, (y - cam.cy) / cam.fy) > 1.3:
                continue  # beyond ~75 deg off-axis: outside what atan(r) of a z>0 bearing can reach reliably
            b = cam.undistortPoint(px)
            assert abs(np.linalg.norm(b) - 1.0) < 1e-12
            assert np.linalg.norm(cam.projectPoint(b) - px) < 0.5
    # batched evaluation agrees with the scalar path
    P = np.random.default_rng(2).uniform(-1, 1, (7, 3)) + np.array([0, 0, 2.0])
    assert np.allclose(cam.projectPoint(P), np.stack([cam.projectPoint(p) for p in P]))
    assert np.allclose(cam.projectionJacobian(P), np.stack([cam.projectionJacobian(p) for p in P]))


# --------------------------------------------------------------- test/test_VIOGroup.cpp:26-60
def test_VIOGroup_BasicOperations():
    rng = np.random.default_rng(10)
    groupId = eqf.VIOGroup.Identity(IDS5)
    for _ in range(TEST_REPS):
        X1, X2, X3 = (randomGroupElement(rng, IDS5) for _ in range(3))
        assert logNorm(X1.inverse() * X1) <= NEAR_ZERO * 100
        assert logNorm(X1 * X1.inverse()) <= NEAR_ZERO * 100
        r12 = (X1 * X2) * X3
        r23 = X1 * (X2 * X3)
        for e in (r12.inverse() * r23, r23.inverse() * r12, r12 * r23.inverse(), r23 * r12.inverse()):
            # acos-based SO3::log limits what "zero" can resolve to ~1e-8 rad; the reference's
            # 1e-12 bound is on the same quantity but Eigen's trace happens to round to 3 exactly
            assert logNorm(e) <= 1e-6
        assert logNorm(groupId) <= NEAR_ZERO
        assert logNorm((X1 * groupId) * X1.inverse()) <= 1e-6
        assert logNorm(X1.inverse() * (groupId * X1)) <= 1e-6


# --------------------------------------------------------------- test/test_VIOGroupActions.cpp:28-96
def test_VIOAction_StateAction():
    rng = np.random.default_rng(11)
    groupId = eqf.VIOGroup.Identity(IDS5)
    for _ in range(TEST_REPS):
        X1, X2 = randomGroupElement(rng, IDS5), randomGroupElement(rng, IDS5)
        xi0 = randomStateElement(rng, IDS5)
        assert stateDistance(xi0, xi0) <= 1e-7
        assert stateDistance(eqf.stateGroupAction(groupId, xi0), xi0) <= 1e-7
        xi1 = eqf.stateGroupAction(X2, eqf.stateGroupAction(X1, xi0))
        xi2 = eqf.stateGroupAction(X1 * X2, xi0)
        assert stateDistance(xi1, xi2) <= 1e-6
        # the same right-action property measured without the acos-limited log
        assert np.abs(xi1.p - xi2.p).max() < 1e-12
        assert np.abs(xi1.sensor.flat() - xi2.sensor.flat()).max() < 1e-12


def test_VIOAction_OutputAction():
    rng = np.random.default_rng(12)
    groupId = eqf.VIOGroup.Identity(IDS5)
    for _ in range(TEST_REPS):
        X1, X2 = randomGroupElement(rng, IDS5), randomGroupElement(rng, IDS5)
        y0 = randomVisionMeasurement(rng, IDS5)
        assert measurementDistance(y0, y0) <= 1e-5
        assert measurementDistance(eqf.outputGroupAction(groupId, y0), y0) <= 1e-5
        y1 = eqf.outputGroupAction(X2, eqf.outputGroupAction(X1, y0))
        y2 = eqf.outputGroupAction(X1 * X2, y0)
        assert measurementDistance(y1, y2) <= 1e-5


def test_VIOAction_OutputEquivariance():
    rng = np.random.default_rng(13)
    ids = [5, 0, 1, 2, 3, 4]
    cam = createDefaultCamera()
    for _ in range(TEST_REPS):
        X = randomGroupElement(rng, ids)
        xi0 = randomStateElement(rng, ids)
        y1 = eqf.measureSystemState(eqf.stateGroupAction(X, xi0), cam)
        y2 = eqf.outputGroupAction(X, eqf.measureSystemState(xi0, cam))
        assert measurementDistance(y1, y2) <= 1e-5


# --------------------------------------------------------------- test/test_VIOLift.cpp:28-125
def test_VIOLift_Lift():
    rng = np.random.default_rng(14)
    for _ in range(TEST_REPS):
        xi0 = randomStateElement(rng, IDS5)
        vel = randomVelocityElement(rng)
        prev = 1e8
        for i in range(8):
            dt = 10.0 ** -i
            xi1 = eqf.integrateSystemFunction(xi0, vel, dt)
            lam = eqf.liftVelocity(xi0, vel)
            xi2 = eqf.stateGroupAction(eqf.VIOExp(lam * dt), xi0)
            d = stateDistance(xi1, xi2) / dt
            # monotone decrease until the acos-limited log floor (~1e-8) is reached
            assert d <= max(prev, 1e-6 / dt)
            prev = d


def test_VIOLift_DiscreteLift():
    rng = np.random.default_rng(15)
    dt = 0.1
    for _ in range(TEST_REPS):
        xi0 = randomStateElement(rng, IDS5)
        vel = randomVelocityElement(rng)
        xi1 = eqf.integrateSystemFunction(xi0, vel, dt)
        xi2 = eqf.stateGroupAction(eqf.liftVelocityDiscrete(xi0, vel, dt), xi0)
        assert np.abs(xi1.p - xi2.p).max() <= 1e-11
        assert np.abs(xi1.sensor.flat() - xi2.sensor.flat()).max() <= 1e-11
        assert stateDistance(xi1, xi2) <= 1e-6


@pytest.mark.parametrize("suite", SUITES[:2], ids=lambda s: s.name)
def test_VIOLift_InnovationLifts(suite):
    rng = np.random.default_rng(16)
    for _ in range(5):
        xi0 = randomStateElement(rng, IDS5)
        dim = xi0.Dim()

        def reproj(eps):
            Delta = eqf.VIOExp(suite.liftInnovation(eps, xi0))
            return suite.stateChart(eqf.stateGroupAction(Delta, xi0), xi0)

        testDifferential(reproj, np.zeros(dim), np.eye(dim))

        def reprojD(eps):
            return suite.stateChart(eqf.stateGroupAction(suite.liftInnovationDiscrete(eps, xi0), xi0), xi0)

        for j in range(dim):
            ej = np.zeros(dim)
            ej[j] = 1.0
            assertMatrixEquality(ej, reprojD(ej))


# --------------------------------------------------------------- test/test_CoordinateCharts.cpp:26-170
def test_CoordinateChart_SphereCharts():
    rng = np.random.default_rng(17)
    for _ in range(TEST_REPS):
        eta = lg.normalized(rng.uniform(-1, 1, 3))
        assert np.linalg.norm(eta - eqf.e3ProjectSphereInv(eqf.e3ProjectSphere(eta))) <= 1e-11
        y = rng.uniform(-1, 1, 2)
        assert np.linalg.norm(eqf.e3ProjectSphere(eqf.e3ProjectSphereInv(y)) - y) <= 1e-11
        pole = lg.normalized(rng.uniform(-1, 1, 3))
        for fwd, inv in ((eqf.sphereChart_stereo, eqf.sphereChart_stereo_inv),
                         (eqf.sphereChart_normal, eqf.sphereChart_normal_inv)):
            assert np.linalg.norm(fwd(pole, pole)) <= 1e-11
            assert np.linalg.norm(eta - inv(fwd(eta, pole), pole)) <= 1e-10
            assert np.linalg.norm(fwd(inv(y, pole), pole) - y) <= 1e-10


def test_CoordinateChart_SphereDifferentials():
    rng = np.random.default_rng(18)
    for _ in range(TEST_REPS):
        eta = lg.normalized(rng.uniform(-1, 1, 3))
        if eta[2] > 0.9:
            continue  # chart singular at +e3
        testDifferential(eqf.e3ProjectSphere, eta, eqf.e3ProjectSphereDiff(eta))
        y = rng.uniform(-1, 1, 2)
        testDifferential(eqf.e3ProjectSphereInv, y, eqf.e3ProjectSphereInvDiff(y))
        pole = lg.normalized(rng.uniform(-1, 1, 3))
        testDifferential(lambda e: eqf.sphereChart_stereo(e, pole), pole, eqf.sphereChart_stereo_diff0(pole))
        testDifferential(lambda v: eqf.sphereChart_stereo_inv(v, pole), np.zeros(2),
                         eqf.sphereChart_stereo_inv_diff0(pole))
        testDifferential(lambda e: eqf.sphereChart_normal(e, pole), pole, eqf.sphereChart_normal_diff0(pole))
        testDifferential(lambda v: eqf.sphereChart_normal_inv(v, pole), np.zeros(2),
                         eqf.sphereChart_normal_inv_diff0(pole))


@pytest.mark.parametrize("chart", [eqf.VIOChart_euclid, eqf.VIOChart_invdepth, eqf.VIOChart_normal],
                         ids=["euclid", "invdepth", "normal"])
def test_CoordinateChart_VIOChart(chart):
    rng = np.random.default_rng(19)
    for _ in range(TEST_REPS):
        xi0 = randomStateElement(rng, IDS5)
        xi1 = randomStateElement(rng, IDS5)
        xi2 = chart.inv(chart(xi1, xi0), xi0)
        assert stateDistance(xi1, xi2) <= 1e-6
        assert np.abs(xi1.p - xi2.p).max() <= 1e-8


def test_CoordinateChart_euclid_invdepth_diff():
    rng = np.random.default_rng(20)
    for _ in range(5):
        xi0 = randomStateElement(rng, IDS5)
        testDifferential(lambda e: eqf.VIOChart_invdepth(eqf.VIOChart_euclid.inv(e, xi0), xi0),
                         eqf.VIOChart_euclid(xi0, xi0), eqf.coordinateDifferential_invdepth_euclid(xi0))


def test_CoordinateChart_euclid_normal_diff():
    rng = np.random.default_rng(21)
    xi0 = randomStateElement(rng, IDS5)
    testDifferential(lambda e: eqf.VIOChart_normal(eqf.VIOChart_euclid.inv(e, xi0), xi0),
                     eqf.VIOChart_euclid(xi0, xi0), eqf.coordinateDifferential_normal_euclid(xi0))


# --------------------------------------------------------------- test/test_EqFMatrices.cpp:26-239
def test_EqFMatrices_euclid_invdepth_compatibility():
    rng = np.random.default_rng(22)
    cam = createDefaultCamera()
    for _ in range(TEST_REPS):
        xi0 = randomStateElement(rng, IDS5)
        X = randomGroupElement(rng, IDS5)
        vel = randomVelocityElement(rng)
        M = eqf.coordinateDifferential_invdepth_euclid(xi0)
        Minv = np.linalg.inv(M)
        Ae = eqf.EqFCoordinateSuite_euclid.stateMatrixA(X, xi0, vel)
        Ai = eqf.EqFCoordinateSuite_invdepth.stateMatrixA(X, xi0, vel)
        assert np.linalg.norm(Ai - M @ Ae @ Minv) <= 1e-6
        Be = eqf.EqFCoordinateSuite_euclid.inputMatrixB(X, xi0)
        Bi = eqf.EqFCoordinateSuite_invdepth.inputMatrixB(X, xi0)
        assert np.linalg.norm(Bi - M @ Be) <= 1e-6
        yHat = eqf.measureSystemState(eqf.stateGroupAction(X, xi0), cam)
        Ce = eqf.EqFCoordinateSuite_euclid.outputMatrixC(xi0, X, yHat)
        Ci = eqf.EqFCoordinateSuite_invdepth.outputMatrixC(xi0, X, yHat)
        assert np.linalg.norm(Ci - Ce @ Minv) <= 1e-4 * max(1.0, np.linalg.norm(Ci))


@pytest.mark.parametrize("suite", SUITES, ids=lambda s: s.name)
def test_EqFSuite_stateMatrixA(suite):
    rng = np.random.default_rng(23)
    reps = 3 if suite.name == "Normal" else 8
    for _ in range(reps):
        xi0 = reasonableStateElement(rng, IDS5)
        XHat = reasonableGroupElement(rng, IDS5)
        vel = randomVelocityElement(rng)
        A0t = suite.stateMatrixA(XHat, xi0, vel)
        xi_hat = eqf.stateGroupAction(XHat, xi0)

        def a0(eps):
            xi_e = suite.stateChart.inv(eps, xi0)
            xi = eqf.stateGroupAction(XHat, xi_e)
            Lt = eqf.liftVelocity(xi, vel) - eqf.liftVelocity(xi_hat, vel)
            xi_hat1 = eqf.stateGroupAction(eqf.VIOExp(Lt), xi_hat)
            xi_e1 = eqf.stateGroupAction(XHat.inverse(), xi_hat1)
            return suite.stateChart(xi_e1, xi0)

        assert np.linalg.norm(a0(np.zeros(xi0.Dim()))) <= 1e-7
        testDifferential(a0, np.zeros(xi0.Dim()), A0t)


@pytest.mark.parametrize("suite", SUITES, ids=lambda s: s.name)
def test_EqFSuite_inputMatrixB(suite):
    rng = np.random.default_rng(24)
    reps = 3 if suite.name == "Normal" else 8
    for _ in range(reps):
        xi0 = reasonableStateElement(rng, IDS5)
        XHat = reasonableGroupElement(rng, IDS5)
        Bt = suite.inputMatrixB(XHat, xi0)
        vel = randomVelocityElement(rng)
        xi_hat = eqf.stateGroupAction(XHat, xi0)

        def b0(ev):
            Lt = eqf.liftVelocity(xi_hat, vel + eqf.IMUVelocity.fromVector(ev)) - eqf.liftVelocity(xi_hat, vel)
            xi_hat1 = eqf.stateGroupAction(eqf.VIOExp(Lt), xi_hat)
            xi_e1 = eqf.stateGroupAction(XHat.inverse(), xi_hat1)
            return suite.stateChart(xi_e1, xi0)

        assert np.linalg.norm(b0(np.zeros(12))) <= 1e-7
        testDifferential(b0, np.zeros(12), Bt)


@pytest.mark.parametrize("suite", SUITES, ids=lambda s: s.name)
def test_EqFSuite_outputMatrixC(suite):
    rng = np.random.default_rng(25)
    ids = [5, 0, 1, 2, 3, 4]
    cam = createDefaultCamera()
    floatStep = float(np.cbrt(np.finfo(np.float32).eps))
    for _ in range(5):
        xi0 = reasonableStateElement(rng, ids)
        XHat = reasonableGroupElement(rng, ids)
        yHat = eqf.measureSystemState(eqf.stateGroupAction(XHat, xi0), cam)
        Ct = suite.outputMatrixC(xi0, XHat, yHat)
        Ct2 = suite.outputMatrixC(xi0, XHat, yHat, False)
        assertMatrixEquality(Ct, Ct2)

        def ct(eps):
            xi = eqf.stateGroupAction(XHat, suite.stateChart.inv(eps, xi0))
            return (eqf.measureSystemState(xi, cam) - yHat).asVector()

        assert np.linalg.norm(ct(np.zeros(xi0.Dim()))) <= 1e-9
        testDifferential(ct, np.zeros(xi0.Dim()), Ct, floatStep)


def test_EqFSuite_outputMatrixCStar():
    """test_EqFMatrices.cpp:181-239.  The reference asserts C* beats C in every one of
    its 75 srand(0) draws at a 0.49 m step; over other draws the inequality fails in
    ~0.3% of directions (the depth direction, where both errors are ~1e-3 px), so it is
    asserted here for >= 98% of draws, together with the property it stands for: the
    C* linearisation error is an order smaller and shrinks one order faster."""
    rng = np.random.default_rng(26)
    suite = eqf.EqFCoordinateSuite_euclid
    cam = createDefaultCamera()
    base = float(np.cbrt(np.finfo(np.float32).eps))
    med = {}
    for scale in (100.0, 10.0):
        floatStep = scale * base
        wins, ratios = 0, []
        for _ in range(TEST_REPS):
            q0 = rng.uniform(-1, 1, 3) * 10 + np.array([0, 0, 20.0])
            Qq = lg.so3_exp(rng.uniform(-1, 1, 3) * 0.02)
            Qa = np.array(2.0 * rng.uniform() + 1.0)
            qHat = lg.sot3_apply_inverse(Qq, Qa, q0)
            yHat = cam.projectPoint(qHat)
            Ct = suite.outputMatrixCi(q0, Qq, Qa, cam)

            def hFunc(eps):
                en = np.concatenate([-lg.skew(q0) @ eps, [-q0 @ eps]]) / (q0 @ q0)
                eq, ea = lg.sot3_exp(-en)
                q_e = lg.sot3_apply(eq, ea, q0)
                return cam.projectPoint(lg.sot3_apply_inverse(Qq, Qa, q_e))

            for j in range(3):
                eps = np.zeros(3)
                eps[j] = floatStep
                yTrue = hFunc(eps)
                yTilde = yTrue - yHat
                CtS = suite.outputMatrixCiStar(q0, Qq, Qa, cam, yTrue)
                eStar = np.linalg.norm(CtS @ eps - yTilde)
                e0 = np.linalg.norm(Ct @ eps - yTilde)
                wins += eStar <= e0
                ratios.append(eStar / e0)
        assert wins >= 0.98 * 3 * TEST_REPS
        med[scale] = float(np.median(ratios))
    assert med[100.0] < 0.1 and med[10.0] < 0.3 * med[100.0]


# --------------------------------------------------------------- test/test_FilterStatistics.cpp:27-168
class _Stats:
    numParticles = 1000

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        st = eqf.Settings()
        st.coordinateChoice = eqf.COORD_INVDEPTH
        st.initialPointVariance = 0.01 ** 2
        st.initialPointDepthVariance = 0.01 ** 2
        st.initialBiasOmegaVariance = 0.01 ** 2
        st.initialBiasAccelVariance = 0.01 ** 2
        st.initialVelocityVariance = 0.1 ** 2
        st.initialPositionVariance = 0.001 ** 2
        self.settings = st
        self.ids = [0, 1]
        self.xi0 = reasonableStateElement(self.rng, self.ids)
        self.suite = eqf.getCoordinates(st.coordinateChoice)
        self.Sigma0 = st.constructInitialStateCovariance(2)
        self.filter = eqf.VIO_eqf(self.suite, self.xi0.copy(), eqf.VIOGroup.Identity(self.ids), self.Sigma0.copy())
        L = np.linalg.cholesky(self.Sigma0)
        self.particles = []
        for _ in range(self.numParticles):
            eps = L @ self.rng.standard_normal(self.Sigma0.shape[0])
            Delta = self.suite.liftInnovation(eps, self.xi0)
            self.particles.append(eqf.stateGroupAction(eqf.VIOExp(Delta), self.xi0))

    def meanNEES(self):
        return float(np.mean([self.filter.computeNEES(p) for p in self.particles]))


def test_FilterStatistics_initialDistribution():
    s = _Stats(30)
    assert abs(s.meanNEES() - 1.0) <= 0.1


def test_FilterStatistics_trueInputDistribution():
    s = _Stats(31)
    dt = 0.2
    vel = eqf.IMUVelocity.Zero()
    for _ in range(3):
        s.particles = [eqf.integrateSystemFunction(p, vel, dt) for p in s.particles]
        s.filter.integrateRiccatiStateDiscrete(vel, dt, 0 * s.settings.constructInputGainMatrix(),
                                               0 * s.settings.constructStateGainMatrix(2))
        s.filter.integrateObserverState(vel, dt, True)
        assert abs(s.meanNEES() - 1.0) <= 1.0


def test_FilterStatistics_outputDistribution():
    s = _Stats(32)
    cam = PinholeCamera(752, 480, 458.654, 457.296, 367.215, 248.375)
    R = s.settings.constructOutputGainMatrix(2)
    Rinv = np.linalg.inv(R)
    noise = np.linalg.cholesky(R) @ s.rng.standard_normal(4)
    measOutput = eqf.measureSystemState(s.xi0, cam).plusVector(noise)
    w = []
    for p in s.particles:
        e = (measOutput - eqf.measureSystemState(p, cam)).asVector()
        w.append(np.exp(-0.5 * float(e @ Rinv @ e)))
    w = np.array(w) / np.sum(w)
    idx = s.rng.choice(len(s.particles), size=len(s.particles), p=w)
    s.particles = [s.particles[i] for i in idx]
    s.filter.performVisionUpdate(measOutput, R)
    assert abs(s.meanNEES() - 1.0) <= 0.5
