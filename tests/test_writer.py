"""Output side: CSV formats of the reference's VIOWriter and the trajectory-error summary."""
import os

import numpy as np
import pytest

eb = pytest.importorskip("eqvio_b200")


def _state():
    s = eb.VIOSensorState()
    s.inputBias = np.array([0.001, -0.002, 0.003, 0.1, 0.2, -0.3])
    s.pose_q = np.array([np.cos(0.3), 0.0, 0.0, np.sin(0.3)])
    s.pose_x = np.array([1.0, 2.0, 1.0 / 3.0])
    s.velocity = np.array([0.5, 1e-5, -123456.789])
    s.cameraOffset_q = np.array([0.5, 0.5, -0.5, 0.5])
    s.cameraOffset_x = np.array([0.01, 0.02, 0.03])
    return eb.VIOState(s, np.array([[0.0, 0.0, 2.0], [1.0, -1.0, 3.0]]), np.array([7, 3]))


def test_csv_lines_follow_the_reference_format(tmp_path):
    xi = _state()
    with eb.VIOWriter(str(tmp_path / "out")) as w:
        w.writeStates(1403715273.2621431351, xi)
        w.writeFeatures(0.05, [9, 2], [[10.5, 20.25], [1.0 / 3.0, 400.0]])
        w.writeTiming(12.5, {"propagation": 1e-4, "correction": 0.25})
        w.writeLandmarkError(0.05, xi, eb.VIOState(xi.sensor, xi.p[:1] + 0.1, xi.ids[:1]))
    d = str(tmp_path / "out")
    imu = open(os.path.join(d, "IMUState.csv")).read().splitlines()
    assert imu[0] == "time, px, py, pz, qw, qx, qy, qz, vx, vy, vz"
    # stamp with setprecision(20) (%.20g), every other entry with the default six significant digits (%g)
    assert imu[1] == "1403715273.2621431351, 1, 2, 0.333333, 0.955336, 0, 0, 0.29552, 0.5, 1e-05, -123457"
    cam = open(os.path.join(d, "camera.csv")).read().splitlines()
    assert cam[1].endswith("0.01, 0.02, 0.03, 0.5, 0.5, -0.5, 0.5")
    pts = open(os.path.join(d, "points.csv")).read().splitlines()[1].split(", ")
    assert pts[1] == "7" and pts[5] == "3" and len(pts) == 9  # time, then (id, x, y, z) per landmark in state order
    # world-frame point = pose * cameraOffset * p
    from eqvio_b200.writer import _qrot, _se3_mul
    q, x = _se3_mul(xi.sensor.pose_q, xi.sensor.pose_x, xi.sensor.cameraOffset_q, xi.sensor.cameraOffset_x)
    np.testing.assert_allclose([float(v) for v in pts[2:5]], _qrot(q, xi.p[0]) + x, rtol=1e-5)
    feat = open(os.path.join(d, "features.csv")).read().splitlines()
    assert feat[1] == "0.050000000000000002776, 2, 0.333333, 400, 9, 10.5, 20.25"  # ids ascending (std::map)
    tim = open(os.path.join(d, "timing.csv")).read().splitlines()
    assert tim[0] == "time, correction, propagation" and tim[1] == "12.5, 0.25, 0.0001"
    lme = open(os.path.join(d, "landmarkError.csv")).read().splitlines()[1].split(", ")
    assert abs(float(lme[1]) - np.sqrt(0.03)) < 1e-6 and lme[2] == "nan"


def test_trajectory_errors_recover_a_similarity_transform():
    rng = np.random.default_rng(0)
    T = 200
    t = np.linspace(0, 20, T)
    tru = np.zeros((T, 11))
    tru[:, 0] = t
    tru[:, 1:4] = np.stack([np.cos(t), np.sin(t), 0.2 * np.sin(3 * t)], 1)
    tru[:, 4] = 1.0
    tru[:, 8:11] = rng.normal(size=(T, 3))
    # the estimate lives in a rotated, shifted, 2% larger frame: the errors after alignment must vanish
    ang = 0.7
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
    est = tru.copy()
    est[:, 1:4] = (R.T @ (tru[:, 1:4] - np.array([1.0, -2.0, 0.5])).T).T / 1.02
    est[:, 4:8] = np.array([np.cos(-ang / 2), 0, 0, np.sin(-ang / 2)])
    res = eb.trajectory_errors(est, tru)
    assert res["position (m)"]["rmse"] < 1e-9 and res["attitude (d)"]["max"] < 1e-5 and abs(res["scale"] - 1.02) < 1e-9
    assert res["velocity (m/s)"]["rmse"] == 0.0
    est[:, 1] += 0.01 * rng.normal(size=T)
    assert 0.004 < eb.trajectory_errors(est, tru)["position (m)"]["rmse"] < 0.02
