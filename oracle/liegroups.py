"""Lie-group primitives of the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates the arithmetic of the vendored LiePP headers and the Eigen quaternion
routines they call.  Rotations are stored like LiePP stores them: as
quaternions ``(w, x, y, z)`` that are *never renormalised*
(external/LiePP/include/liepp/SO3.h:84-110).  All functions broadcast over
leading axes so that the N landmark transforms are handled as one array.

Reference files followed:
  external/LiePP/include/liepp/SO3.h   (exp :42-54, log :56-63, FromVectors :77-81)
  external/LiePP/include/liepp/SE3.h   (adjoint :50-58, exp :59-84, log :85-104, Adjoint :168-176)
  external/LiePP/include/liepp/SOT3.h  (exp :48-53, log :54-59, action :95-97)
  external/LiePP/include/liepp/SEn3.h  (exp :66-93, log :94-113, Adjoint :192-201)
Eigen semantics restated from the published Eigen 3 algorithm (Eigen itself is
not vendored by the reference): Quaternion product, ``toRotationMatrix``,
matrix->quaternion assignment, ``_transformVector``, ``inverse``,
``setFromTwoVectors``.
"""

import numpy as np

# ----------------------------------------------------------------------------
# small vector helpers
# ----------------------------------------------------------------------------


def skew(v):
    """3x3 cross-product matrix (SO3.h:33-35). Broadcasts over leading axes."""
    v = np.asarray(v, dtype=np.float64)
    z = np.zeros(v.shape[:-1])
    return np.stack(
        [
            np.stack([z, -v[..., 2], v[..., 1]], -1),
            np.stack([v[..., 2], z, -v[..., 0]], -1),
            np.stack([-v[..., 1], v[..., 0], z], -1),
        ],
        -2,
    )


def vex(M):
    """Inverse of skew (SO3.h:38)."""
    return np.stack([M[..., 2, 1], M[..., 0, 2], M[..., 1, 0]], -1)


def normalized(v):
    """Eigen ``normalized()``: v / sqrt(|v|^2) when |v|^2 > 0."""
    v = np.asarray(v, dtype=np.float64)
    n2 = np.sum(v * v, -1, keepdims=True)
    n = np.sqrt(n2)
    return np.where(n2 > 0, v / np.where(n2 > 0, n, 1.0), v)


def norm(v):
    return np.sqrt(np.sum(np.asarray(v) ** 2, -1))


def cross3(a, b):
    """a x b on the last axis (same arithmetic as np.cross, without its axis-shuffling overhead)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a0, a1, a2 = a[..., 0], a[..., 1], a[..., 2]
    b0, b1, b2 = b[..., 0], b[..., 1], b[..., 2]
    return np.stack([a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0], -1)


# ----------------------------------------------------------------------------
# Eigen quaternion arithmetic, (w, x, y, z)
# ----------------------------------------------------------------------------

QUAT_IDENTITY = np.array([1.0, 0.0, 0.0, 0.0])


def quat_mul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
        ],
        -1,
    )


def quat_inv(q):
    """Eigen ``Quaternion::inverse``: conjugate / squaredNorm."""
    n2 = np.sum(q * q, -1, keepdims=True)
    conj = q * np.array([1.0, -1.0, -1.0, -1.0])
    return conj / n2


def quat_rotate(q, v):
    """Eigen ``_transformVector``: v + w*(2 u x v) + u x (2 u x v)."""
    u = q[..., 1:]
    w = q[..., :1]
    uv = cross3(u, v)
    uv = uv + uv
    return v + w * uv + cross3(u, uv)


def quat_to_matrix(q):
    """Eigen ``toRotationMatrix`` (assumes, but does not enforce, unit norm)."""
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.stack(
        [
            np.stack([1.0 - (tyy + tzz), txy - twz, txz + twy], -1),
            np.stack([txy + twz, 1.0 - (txx + tzz), tyz - twx], -1),
            np.stack([txz - twy, tyz + twx, 1.0 - (txx + tyy)], -1),
        ],
        -2,
    )


def matrix_to_quat(m):
    """Eigen matrix->quaternion assignment (trace-positive branch, else the
    largest-diagonal branch).  Single 3x3 only (sensor-state sized work)."""
    m = np.asarray(m, dtype=np.float64)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.empty(4)
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
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def quat_from_two_vectors(a, b):
    """Eigen ``setFromTwoVectors(a, b)``: rotation taking a to b.

    Both inputs are normalised; c = v1.v0; axis = v0 x v1; s = sqrt(2(1+c));
    vec = axis / s, w = s / 2.  Near-antiparallel inputs (c < -1 + 1e-12) take
    the SVD branch.  Broadcasts over leading axes.
    """
    v0 = normalized(a)
    v1 = normalized(b)
    c = np.sum(v1 * v0, -1)
    axis = cross3(v0, v1)
    s = np.sqrt((1.0 + c) * 2.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        invs = 1.0 / s
        q = np.concatenate([(s * 0.5)[..., None], axis * invs[..., None]], -1)
    bad = c < -1.0 + 1e-12
    if np.any(bad):
        q = np.array(q, copy=True)
        v0b = np.broadcast_to(v0, q.shape[:-1] + (3,))
        v1b = np.broadcast_to(v1, q.shape[:-1] + (3,))
        for idx in zip(*np.nonzero(bad)) if q.ndim > 1 else [()]:
            cc = max(float(c[idx]) if q.ndim > 1 else float(c), -1.0)
            m = np.stack([v0b[idx], v1b[idx]], 0)
            _, _, vt = np.linalg.svd(m, full_matrices=True)
            ax = vt[2]
            w2 = (1.0 + cc) * 0.5
            q[idx] = np.concatenate([[np.sqrt(w2)], ax * np.sqrt(1.0 - w2)])
    return q


# ----------------------------------------------------------------------------
# SO(3)  (quaternion backed)
# ----------------------------------------------------------------------------


def so3_exp(w):
    """SO3::exp builds the quaternion directly; small-angle switch at
    theta/2 <= 1e-6 (SO3.h:42-54)."""
    w = np.asarray(w, dtype=np.float64)
    theta = norm(w) / 2.0
    big = theta > 1e-6
    wn = normalized(w)
    qb = np.concatenate([np.cos(theta)[..., None], np.sin(theta)[..., None] * wn], -1)
    qs = np.concatenate([np.ones_like(theta)[..., None], w / 2.0], -1)
    return np.where(big[..., None], qb, qs)


def so3_log(q):
    """SO3::log via acos((tr R - 1)/2) (SO3.h:56-63)."""
    R = quat_to_matrix(q)
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    with np.errstate(invalid="ignore"):
        theta = np.arccos((tr - 1.0) / 2.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        coef = np.where(np.abs(theta) > 1e-6, theta / (2.0 * np.sin(theta)), 0.5)
    Omega = coef[..., None, None] * (R - np.swapaxes(R, -1, -2))
    return vex(Omega)


# ----------------------------------------------------------------------------
# SE(3): pair (q, x).  Sensor-sized, not vectorised.
# ----------------------------------------------------------------------------


class SE3:
    __slots__ = ("q", "x")

    def __init__(self, q=None, x=None):
        self.q = np.array(QUAT_IDENTITY if q is None else q, dtype=np.float64)
        self.x = np.zeros(3) if x is None else np.array(x, dtype=np.float64)

    @staticmethod
    def Identity():
        return SE3()

    def copy(self):
        return SE3(self.q.copy(), self.x.copy())

    @property
    def R(self):
        return quat_to_matrix(self.q)

    def __mul__(self, other):
        if isinstance(other, SE3):  # SE3.h:156
            return SE3(quat_mul(self.q, other.q), self.x + quat_rotate(self.q, other.x))
        other = np.asarray(other, dtype=np.float64)  # SE3.h:155, point action
        return quat_rotate(self.q, other) + self.x

    def inverse(self):  # SE3.h:162
        qi = quat_inv(self.q)
        return SE3(qi, -quat_rotate(qi, self.x))

    def Adjoint(self):  # SE3.h:168-176
        Rm = self.R
        Ad = np.zeros((6, 6))
        Ad[0:3, 0:3] = Rm
        Ad[3:6, 0:3] = skew(self.x) @ Rm
        Ad[3:6, 3:6] = Rm
        return Ad

    def asMatrix(self):
        M = np.eye(4)
        M[0:3, 0:3] = self.R
        M[0:3, 3] = self.x
        return M


def se3_adjoint(u):
    """SE3::adjoint (little ad), u = (omega, v) (SE3.h:50-58)."""
    ad = np.zeros((6, 6))
    ad[0:3, 0:3] = skew(u[0:3])
    ad[3:6, 3:6] = skew(u[0:3])
    ad[3:6, 0:3] = skew(u[3:6])
    return ad


def _rodrigues_coeffs(th, strict):
    # SE3.h:64-73 uses |th| > 1e-12, SEn3.h:75-83 uses |th| >= 1e-12
    big = (abs(th) > 1e-12) if strict else (abs(th) >= 1e-12)
    if big:
        A = np.sin(th) / th
        B = (1.0 - np.cos(th)) / th**2
        C = (1.0 - A) / th**2
    else:
        A, B, C = 1.0, 0.5, 1.0 / 6.0
    return A, B, C


def se3_exp(u):
    """SE3::exp builds Rodrigues matrices and converts matrix->quaternion
    (SE3.h:59-84, :142-145)."""
    u = np.asarray(u, dtype=np.float64)
    w, v = u[0:3], u[3:6]
    th = np.sqrt(w @ w)
    A, B, C = _rodrigues_coeffs(th, True)
    wx = skew(w)
    R = np.eye(3) + A * wx + B * (wx @ wx)
    V = np.eye(3) + B * wx + C * (wx @ wx)
    return SE3(matrix_to_quat(R), V @ v)


def se3_log(P):
    """SE3::log (SE3.h:85-104)."""
    Omega = skew(so3_log(P.q))
    theta = np.sqrt(np.sum(vex(Omega) ** 2))
    coef = 1.0 / 12.0
    if abs(theta) > 1e-6:
        coef = 1.0 / (theta * theta) * (1.0 - (theta * np.sin(theta)) / (2.0 * (1.0 - np.cos(theta))))
    VInv = np.eye(3) - 0.5 * Omega + coef * (Omega @ Omega)
    return np.concatenate([vex(Omega), VInv @ P.x])


# ----------------------------------------------------------------------------
# SE_2(3): (q, x0, x1)
# ----------------------------------------------------------------------------


def se23_exp(u):
    """SEn3<2>::exp (SEn3.h:66-93). Returns (q, x0, x1)."""
    u = np.asarray(u, dtype=np.float64)
    w = u[0:3]
    th = np.sqrt(w @ w)
    A, B, C = _rodrigues_coeffs(th, False)
    wx = skew(w)
    R = np.eye(3) + A * wx + B * (wx @ wx)
    V = np.eye(3) + B * wx + C * (wx @ wx)
    return matrix_to_quat(R), V @ u[3:6], V @ u[6:9]


def se23_log(q, x0, x1):
    """SEn3<2>::log (SEn3.h:94-113)."""
    Omega = skew(so3_log(q))
    theta = np.sqrt(np.sum(vex(Omega) ** 2))
    coef = 1.0 / 12.0
    if abs(theta) > 1e-8:
        coef = 1.0 / (theta * theta) * (1.0 - (theta * np.sin(theta)) / (2.0 * (1.0 - np.cos(theta))))
    VInv = np.eye(3) - 0.5 * Omega + coef * (Omega @ Omega)
    return np.concatenate([vex(Omega), VInv @ x0, VInv @ x1])


# ----------------------------------------------------------------------------
# SOT(3): arrays (q[...,4], a[...])
# ----------------------------------------------------------------------------


def sot3_exp(W):
    """SOT3::exp (SOT3.h:48-53). W[...,4] -> (q, a)."""
    W = np.asarray(W, dtype=np.float64)
    return so3_exp(W[..., 0:3]), np.exp(W[..., 3])


def sot3_log(q, a):
    """SOT3::log (SOT3.h:54-59)."""
    return np.concatenate([so3_log(q), np.log(a)[..., None]], -1)


def sot3_apply(q, a, p):
    """Q * p = a * (R p) (SOT3.h:95)."""
    return np.asarray(a)[..., None] * quat_rotate(q, p)


def sot3_apply_inverse(q, a, p):
    """Q.inverse() * p: inverse() = (R^-1, 1/a) (SOT3.h:103) then action."""
    return (1.0 / np.asarray(a))[..., None] * quat_rotate(quat_inv(q), p)
