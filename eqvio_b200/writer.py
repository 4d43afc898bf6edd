"""Output side of the path: the reference's ``VIOWriter`` CSV files (src/VIOWriter.cpp:33-228) written
from the ctypes ``VIOFilter`` mirror, and the trajectory-error summary of scripts/analysis_tools.py:85-181.

Formatting follows the reference byte for byte: the time stamp goes through ``std::setprecision(20)`` (``%.20g``), every
other number through its own default ``std::stringstream`` (``%g``, six significant digits), entries are joined by ``", "``
(include/eqvio/csv/CSVLine.h:108-113,153-161); SE(3) elements are written position first, then the quaternion as w, x, y, z
(CSVLine.h:217-248).  The C++ ``VIOWriter`` itself keeps working unchanged above ``include/eqvio_b200_facade.hpp``; this
module is the same thing for Python hosts.  Host-side code only: no arithmetic of the filter lives here.
"""
from __future__ import annotations

import os

import numpy as np


def _g(x) -> str:
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    return "%g" % float(x)


def _line(stamp, entries) -> str:
    return "%.20g, " % float(stamp) + ", ".join(_g(e) for e in entries) + "\n"


def _qrot(q, v):
    w, x, y, z = q
    u = np.array([x, y, z])
    v = np.asarray(v, dtype=np.float64)
    return v + 2.0 * w * np.cross(u, v) + 2.0 * np.cross(u, np.cross(u, v))


def _qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def _qinv(q):
    q = np.asarray(q, dtype=np.float64)
    return np.array([q[0], -q[1], -q[2], -q[3]]) / float(np.dot(q, q))


def _qmat(q):
    return np.stack([_qrot(q, e) for e in np.eye(3)], axis=1)


def _so3_log(q):  # SO3.h:56-63
    R = _qmat(q)
    theta = np.arccos(np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0))
    coef = theta / (2.0 * np.sin(theta)) if abs(theta) > 1e-6 else 0.5
    return coef * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])


def _se3_log(q, x):  # SE3.h:85-104
    om = _so3_log(q)
    O = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    theta = float(np.linalg.norm(om))
    coef = 1.0 / 12.0
    if abs(theta) > 1e-6:
        coef = 1.0 / (theta * theta) * (1.0 - (theta * np.sin(theta)) / (2.0 * (1.0 - np.cos(theta))))
    VInv = np.eye(3) - 0.5 * O + coef * (O @ O)
    return np.concatenate([om, VInv @ np.asarray(x, dtype=np.float64)])


def _se3_mul(qa, xa, qb, xb):
    return _qmul(qa, qb), np.asarray(xa) + _qrot(qa, xb)


def _se3_inv(q, x):
    qi = _qinv(q)
    return qi, -_qrot(qi, x)


class VIOWriter:
    """Same files, headers and line formats as the reference's VIOWriter (src/VIOWriter.cpp)."""

    HEADERS = {
        "IMUState.csv": "time, px, py, pz, qw, qx, qy, qz, vx, vy, vz\n",
        "camera.csv": "time, px, py, pz, qw, qx, qy, qz\n",
        "bias.csv": "time, bias_gyr_x, bias_gyr_y, bias_gyr_z, bias_acc_x, bias_acc_y, bias_acc_z\n",
        "points.csv": "time, p1id, p1x, p1y, p1z, ...\n",
        "features.csv": "time, z1id, z1x, z1y, ...\n",
        "landmarkError.csv": "time, lm_err_1, lm_err_2, ...\n",
        # VIOWriter.cpp:146-151 (the reference's literals concatenate without a comma between bias_acc_z and num_lm; kept as is)
        "trueState.csv": "time, pose_tx, pose_ty, pose_tz, pose_qw, pose_qx, pose_qy, pose_qz,"
                         "pose_vx, pose_vy, pose_vz, cam_tx, cam_ty, cam_tz, cam_qw, cam_qx, cam_qy, cam_qz,"
                         "bias_gyr_x, bias_gyr_y, bias_gyr_z, bias_acc_x, bias_acc_y, bias_acc_z"
                         "num_lm, lm_1_id, lm_1_x, lm_1_y, lm_1_z, lm_2_id, lm_2_x, lm_2_y, lm_2_z, ...\n",
        "nees.csv": "time, NEES, DoF, PoseNEES, AttitudeNEES\n",
        "poseConsistency.csv": "time, eps_rx, eps_ry, eps_rz, eps_px, eps_py, eps_pz,"
                               "Sigma2_rx, Sigma2_ry, Sigma2_rz, Sigma2_px, Sigma2_py, Sigma2_pz\n",
        "cameraConsistency.csv": "time, eps_rx, eps_ry, eps_rz, eps_px, eps_py, eps_pz,"
                                 "Sigma2_rx, Sigma2_ry, Sigma2_rz, Sigma2_px, Sigma2_py, Sigma2_pz\n",
        "biasConsistency.csv": "time, eps_gyr_x, eps_gyr_y, eps_gyr_z, eps_acc_x, eps_acc_y, eps_acc_z,"
                               "Sigma2_gyr_x, Sigma2_gyr_y, Sigma2_gyr_z, Sigma2_acc_x, Sigma2_acc_y, Sigma2_acc_z\n",
    }

    def __init__(self, outputDir: str):  # VIOWriter.cpp:22-31
        self.outputDir = outputDir if outputDir.endswith("/") else outputDir + "/"
        os.makedirs(self.outputDir, exist_ok=True)
        self._files = {}

    def _file(self, name, header=None):
        f = self._files.get(name)
        if f is None:
            f = open(self.outputDir + name, "w")
            f.write(header if header is not None else self.HEADERS[name])
            self._files[name] = f
        return f

    def close(self):
        for f in self._files.values():
            f.close()
        self._files = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- VIOWriter.cpp:33-81 -------------------------------------------------------------------------------------------
    def writeStates(self, stamp, xi):
        s = xi.sensor
        self._file("IMUState.csv").write(_line(stamp, [*s.pose_x, *s.pose_q, *s.velocity]))
        self._file("camera.csv").write(_line(stamp, [*s.cameraOffset_x, *s.cameraOffset_q]))
        self._file("bias.csv").write(_line(stamp, list(s.inputBias)))
        qPC, xPC = _se3_mul(s.pose_q, s.pose_x, s.cameraOffset_q, s.cameraOffset_x)  # points in the world frame
        entries = []
        for pid, p in zip(xi.ids, np.asarray(xi.p).reshape(-1, 3)):
            entries += [int(pid), *(_qrot(qPC, p) + xPC)]
        self._file("points.csv").write(_line(stamp, entries))

    # -- VIOWriter.cpp:83-96 (ids ascending: std::map iteration) --------------------------------------------------------
    def writeFeatures(self, stamp, ids, y):
        ids = np.asarray(ids).reshape(-1)
        y = np.asarray(y, dtype=np.float64).reshape(-1, 2)
        order = np.argsort(ids, kind="stable")
        entries = []
        for k in order:
            entries += [int(ids[k]), y[k, 0], y[k, 1]]
        self._file("features.csv").write(_line(stamp, entries))

    # -- VIOWriter.cpp:98-116 ------------------------------------------------------------------------------------------
    def writeTiming(self, loopTimeStart, timings: dict):
        labels = sorted(timings)  # std::map iteration order
        f = self._file("timing.csv", ", ".join(["time"] + labels) + "\n")
        f.write(_line(loopTimeStart, [timings[k] for k in labels]))

    # -- VIOWriter.cpp:118-140 -----------------------------------------------------------------------------------------
    def writeLandmarkError(self, stamp, trueState, estState):
        est = {int(i): p for i, p in zip(estState.ids, np.asarray(estState.p).reshape(-1, 3))}
        entries = []
        for i, p in zip(trueState.ids, np.asarray(trueState.p).reshape(-1, 3)):
            entries.append(float(np.linalg.norm(est[int(i)] - p)) if int(i) in est else float("nan"))
        self._file("landmarkError.csv").write(_line(stamp, entries))

    # -- VIOWriter.cpp:142-228 (true-state dump, NEES and the per-block consistency files) -----------------------------
    def writeConsistency(self, stamp, trueState, flt):
        # VIOState.cpp:80-92: sensor (pose, velocity, camera offset, bias), landmark count, then id, x, y, z per landmark
        t = trueState.sensor
        entries = [*t.pose_x, *t.pose_q, *t.velocity, *t.cameraOffset_x, *t.cameraOffset_q, *t.inputBias, int(len(trueState.ids))]
        for pid, p in zip(trueState.ids, np.asarray(trueState.p).reshape(-1, 3)):
            entries += [int(pid), *p]
        self._file("trueState.csv").write(_line(stamp, entries))
        fs = flt.viewEqFState(withSigma=True)
        Sigma = fs.Sigma
        ts = trueState.sensor
        xi0 = fs.xi0.sensor
        Xs = np.asarray(fs.X_sensor, dtype=np.float64)  # beta6 | A (q wxyz, x) | w3 | B (q wxyz, x)
        beta, XA_q, XA_x, XB_q, XB_x = Xs[0:6], Xs[6:10], Xs[10:13], Xs[16:20], Xs[20:23]
        # errorPose = truePose * X.A^-1 ; epsilon = log(xi0.pose^-1 * errorPose)
        ep_q, ep_x = _se3_mul(ts.pose_q, ts.pose_x, *_se3_inv(XA_q, XA_x))
        poseEps = _se3_log(*_se3_mul(*_se3_inv(xi0.pose_q, xi0.pose_x), ep_q, ep_x))
        attEps = _so3_log(_qmul(_qinv(xi0.pose_q), ep_q))
        fullNEES = flt.computeNEES(trueState)
        poseNEES = float(poseEps @ np.linalg.solve(Sigma[6:12, 6:12], poseEps))
        attNEES = float(attEps @ np.linalg.solve(Sigma[6:9, 6:9], attEps))
        self._file("nees.csv").write(_line(stamp, [fullNEES, int(Sigma.shape[0]), poseNEES, attNEES]))
        self._file("poseConsistency.csv").write(_line(stamp, [*poseEps, *np.diag(Sigma)[6:12]]))
        # errorCamera = X.A * trueCameraOffset * X.B^-1
        ec = _se3_mul(*_se3_mul(XA_q, XA_x, ts.cameraOffset_q, ts.cameraOffset_x), *_se3_inv(XB_q, XB_x))
        camEps = _se3_log(*_se3_mul(*_se3_inv(xi0.cameraOffset_q, xi0.cameraOffset_x), *ec))
        self._file("cameraConsistency.csv").write(_line(stamp, [*camEps, *np.diag(Sigma)[15:21]]))
        biasEps = np.asarray(ts.inputBias) - beta - np.asarray(xi0.inputBias)
        self._file("biasConsistency.csv").write(_line(stamp, [*biasEps, *np.diag(Sigma)[0:6]]))


# ---------------------------------------------------------------------------------------------------------------------
# Trajectory error summary (scripts/analysis_tools.py:85-181): align the estimated positions to the true ones with a
# similarity transform (least squares, Umeyama), then position / attitude / velocity error statistics.
# ---------------------------------------------------------------------------------------------------------------------
def align_umeyama(est_xyz, tru_xyz, with_scale=True):
    """Least-squares similarity (s, R, t) with tru ~ s R est + t.  est_xyz, tru_xyz: (T, 3)."""
    est = np.asarray(est_xyz, dtype=np.float64)
    tru = np.asarray(tru_xyz, dtype=np.float64)
    mu_e, mu_t = est.mean(0), tru.mean(0)
    E, T = est - mu_e, tru - mu_t
    U, D, Vt = np.linalg.svd(T.T @ E / est.shape[0])
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1.0
    R = U @ S @ Vt
    var_e = (E ** 2).sum() / est.shape[0]
    s = float(np.trace(np.diag(D) @ S) / var_e) if with_scale and var_e > 0 else 1.0
    t = mu_t - s * R @ mu_e
    return s, R, t


def _stats(v):
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    return {"rmse": float(np.sqrt(np.mean(v ** 2))), "mean": float(np.mean(v)), "med": float(np.median(v)), "std": float(np.std(v)),
            "min": float(np.min(v)), "max": float(np.max(v))}


def trajectory_errors(est, tru):
    """est, tru: arrays (T, 11) of rows ``time, px, py, pz, qw, qx, qy, qz, vx, vy, vz`` (the IMUState.csv columns) at the SAME
    times.  Returns the reference's result dictionary: position (m) / attitude (d) / velocity (m/s) statistics and the scale."""
    est = np.asarray(est, dtype=np.float64)
    tru = np.asarray(tru, dtype=np.float64)
    s, R, t = align_umeyama(est[:, 1:4], tru[:, 1:4])
    pos_al = (s * (R @ est[:, 1:4].T)).T + t
    err_pos = np.linalg.norm(tru[:, 1:4] - pos_al, axis=1)
    err_att = []
    for qe, qt in zip(est[:, 4:8], tru[:, 4:8]):
        Re = R @ _qmat(qe)  # the alignment also rotates the estimated attitudes
        Rt = _qmat(qt)
        dR = Rt @ Re.T
        ang = np.arccos(np.clip((np.trace(dR) - 1.0) / 2.0, -1.0, 1.0))
        err_att.append(np.degrees(ang))
    err_vel = np.linalg.norm(tru[:, 8:11] - est[:, 8:11], axis=1)  # body-frame velocities are alignment-invariant
    length = float(np.sum(np.linalg.norm(np.diff(tru[:, 1:4], axis=0), axis=1)))
    return {"position (m)": _stats(err_pos), "attitude (d)": _stats(err_att), "velocity (m/s)": _stats(err_vel), "scale": s,
            "trajectory length (m)": length}


def read_imu_state_csv(path):
    """IMUState.csv -> (T, 11) array."""
    return np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
