"""Camera models of the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates GIFT's pinhole and radial-tangential ("Standard") cameras:
  external/GIFT/GIFT/src/camera/PinholeCamera.cpp:57-74  (undistort / Jacobian / project)
  external/GIFT/GIFT/src/camera/PinholeCamera.cpp:85-90  (isInDomain, z > 0)
  external/GIFT/GIFT/include/GIFT/camera/GICamera.h:64-67 (isInDomain, pixel bounds)
  external/GIFT/GIFT/src/camera/StandardCamera.cpp:41-145 (radtan project / undistort /
      Jacobian / inverse-distortion least-squares fit)
All methods broadcast over leading axes (N landmarks at once).
"""

import numpy as np

from .liegroups import normalized

MODEL_PINHOLE = 0
MODEL_RADTAN = 1
MODEL_EQUIDISTANT = 2


class PinholeCamera:
    model = MODEL_PINHOLE

    def __init__(self, width, height, fx, fy, cx, cy):
        self.width, self.height = int(width), int(height)
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.dist = []
        self.invDist = []

    # PinholeCamera.cpp:57-61
    def undistortPoint(self, y):
        y = np.asarray(y, dtype=np.float64)
        v = np.stack([(y[..., 0] - self.cx) / self.fx, (y[..., 1] - self.cy) / self.fy, np.ones(y.shape[:-1])], -1)
        return normalized(v)

    # PinholeCamera.cpp:70-74
    def projectPoint(self, p):
        p = np.asarray(p, dtype=np.float64)
        return np.stack(
            [self.fx * p[..., 0] / p[..., 2] + self.cx, self.fy * p[..., 1] / p[..., 2] + self.cy], -1
        )

    # PinholeCamera.cpp:63-68
    def projectionJacobian(self, p):
        p = np.asarray(p, dtype=np.float64)
        x, y, z = p[..., 0], p[..., 1], p[..., 2]
        zero = np.zeros_like(z)
        return np.stack(
            [
                np.stack([self.fx / z, zero, -self.fx * x / (z * z)], -1),
                np.stack([zero, self.fy / z, -self.fy * y / (z * z)], -1),
            ],
            -2,
        )

    # PinholeCamera.cpp:85-90 + GICamera.h:64-67
    def isInDomain(self, p):
        p = np.asarray(p, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            px = self.projectPoint(p)
        ok = (px[..., 0] >= 0) & (px[..., 1] >= 0) & (px[..., 0] < self.width) & (px[..., 1] < self.height)
        return ok & (p[..., 2] > 0)

    def pod(self):
        """Flat description handed to the C-ABI (include/eqvio_b200.h: eqvio_camera)."""
        d = list(self.dist) + [0.0] * (5 - len(self.dist))
        i = list(self.invDist) + [0.0] * (5 - len(self.invDist))
        return dict(model=self.model, fx=self.fx, fy=self.fy, cx=self.cx, cy=self.cy, ndist=len(self.dist),
                    dist=d, inv_dist=i, width=self.width, height=self.height)


def _distort_homogeneous(pt, dist):
    """StandardCamera::distortHomogeneousPoint (StandardCamera.cpp:57-75)."""
    x, y = pt[..., 0], pt[..., 1]
    r2 = x * x + y * y
    dx, dy = x.copy() if isinstance(x, np.ndarray) else x, y.copy() if isinstance(y, np.ndarray) else y
    n = len(dist)
    if n >= 2:
        dx = dx + x * (dist[0] * r2 + dist[1] * r2 * r2)
        dy = dy + y * (dist[0] * r2 + dist[1] * r2 * r2)
    if n >= 4:
        dx = dx + 2 * dist[2] * x * y + dist[3] * (r2 + 2 * x * x)
        dy = dy + 2 * dist[3] * x * y + dist[2] * (r2 + 2 * y * y)
    if n >= 5:
        dx = dx + x * dist[4] * r2 * r2 * r2
        dy = dy + y * dist[4] * r2 * r2 * r2
    return np.stack([dx, dy], -1)


class StandardCamera(PinholeCamera):
    """Pinhole + radtan distortion (k1, k2, p1, p2[, k3])."""

    model = MODEL_RADTAN

    def __init__(self, width, height, fx, fy, cx, cy, dist):
        super().__init__(width, height, fx, fy, cx, cy)
        self.dist = [float(d) for d in dist]
        self.invDist = self.computeInverseDistortion()

    # StandardCamera.cpp:41-48
    def projectPoint(self, p):
        p = np.asarray(p, dtype=np.float64)
        h = np.stack([p[..., 0] / p[..., 2], p[..., 1] / p[..., 2]], -1)
        d = _distort_homogeneous(h, self.dist)
        # PinholeCamera::projectPointEigen on (d, 1)
        return np.stack([self.fx * d[..., 0] / 1.0 + self.cx, self.fy * d[..., 1] / 1.0 + self.cy], -1)

    # StandardCamera.cpp:50-56
    def undistortPoint(self, y):
        u = PinholeCamera.undistortPoint(self, y)
        h = np.stack([u[..., 0] / u[..., 2], u[..., 1] / u[..., 2]], -1)
        d = _distort_homogeneous(h, self.invDist)
        v = np.stack([d[..., 0], d[..., 1], np.ones(d.shape[:-1])], -1)
        return normalized(v)

    # StandardCamera.cpp:77-111
    def projectionJacobian(self, p):
        p = np.asarray(p, dtype=np.float64)
        x, y, z = p[..., 0], p[..., 1], p[..., 2]
        zero = np.zeros_like(z)
        Jh = np.stack(
            [np.stack([1.0 / z, zero, -1.0 * x / (z * z)], -1), np.stack([zero, 1.0 / z, -1.0 * y / (z * z)], -1)], -2
        )
        px, py = x / z, y / z
        r2 = px * px + py * py
        d = self.dist
        n = len(d)
        J = np.zeros(p.shape[:-1] + (2, 2))
        J[..., 0, 0] = 1.0
        J[..., 1, 1] = 1.0
        hp = np.stack([px, py], -1)
        Dr2 = 2.0 * hp
        if n >= 2:
            s = d[0] * r2 + d[1] * r2 * r2
            J[..., 0, 0] += s
            J[..., 1, 1] += s
            J += (hp * (d[0] + 2 * r2 * d[1])[..., None])[..., :, None] * Dr2[..., None, :]
        if n >= 4:
            J[..., 0, 0] += 2.0 * d[2] * py + 6.0 * d[3] * px
            J[..., 0, 1] += 2.0 * d[2] * px + 2.0 * d[3] * py
            J[..., 1, 0] += 2.0 * d[2] * px + 2.0 * d[3] * py
            J[..., 1, 1] += 6.0 * d[2] * py + 2.0 * d[3] * px
        if n >= 5:
            s = d[4] * r2 * r2 * r2
            J[..., 0, 0] += s
            J[..., 1, 1] += s
            J += (hp * (d[4] * 3 * r2 * r2)[..., None])[..., :, None] * Dr2[..., None, :]
        K2 = np.array([[self.fx, 0.0], [0.0, self.fy]])
        return K2 @ J @ Jh

    # StandardCamera.cpp:113-145.  The reference solves the 5-parameter least
    # squares with Eigen's colPivHouseholderQr; numpy's lstsq (SVD) gives the
    # same minimiser to rounding for this full-rank system.
    def computeInverseDistortion(self):
        if self.width * self.height == 0:
            w, h = int(round(self.cx * 2)), int(round(self.cy * 2))
        else:
            w, h = self.width, self.height
        maxPoints = 30
        rows, rhs = [], []
        for x in range(0, w, w // maxPoints):
            for y in range(0, h, h // maxPoints):
                npt = np.array([(x - self.cx) / self.fx, (y - self.cy) / self.fy])
                p = _distort_homogeneous(npt, self.dist)
                r2 = p[0] * p[0] + p[1] * p[1]
                rows.append([p[0] * r2, p[0] * r2 * r2, 2 * p[0] * p[1], r2 + 2 * p[0] * p[0], p[0] * r2 * r2 * r2])
                rows.append([p[1] * r2, p[1] * r2 * r2, r2 + 2 * p[1] * p[1], 2 * p[0] * p[1], p[1] * r2 * r2 * r2])
                rhs.append(npt[0] - p[0])
                rhs.append(npt[1] - p[1])
        sol, *_ = np.linalg.lstsq(np.array(rows), np.array(rhs), rcond=None)
        return [float(s) for s in sol]


class EquidistantCamera(PinholeCamera):
    """Pinhole + Kannala-Brandt equidistant (fisheye) distortion, 4 coefficients
    (external/GIFT/GIFT/src/camera/EquidistantCamera.cpp)."""

    model = MODEL_EQUIDISTANT

    def __init__(self, width, height, fx, fy, cx, cy, dist):
        super().__init__(width, height, fx, fy, cx, cy)
        self.dist = [float(d) for d in dist][:4]
        self.invDist = []

    # EquidistantCamera.cpp:70-81
    def _distort(self, h):
        h = np.asarray(h, dtype=np.float64)
        r = np.sqrt(h[..., 0] * h[..., 0] + h[..., 1] * h[..., 1])
        theta = np.arctan(r)
        d = self.dist
        temp = theta * (1.0 + d[0] * theta**2 + d[1] * theta**4 + d[2] * theta**6 + d[3] * theta**8)
        with np.errstate(divide="ignore", invalid="ignore"):
            scale = np.where(r > 1e-6, temp / np.where(r > 1e-6, r, 1.0), 1.0)
        return scale[..., None] * h

    # EquidistantCamera.cpp:38-46
    def projectPoint(self, p):
        p = np.asarray(p, dtype=np.float64)
        h = np.stack([p[..., 0] / p[..., 2], p[..., 1] / p[..., 2]], -1)
        d = self._distort(h)
        return np.stack([self.fx * d[..., 0] / 1.0 + self.cx, self.fy * d[..., 1] / 1.0 + self.cy], -1)

    # EquidistantCamera.cpp:48-68: damped Gauss-Newton on the sphere, stops at 0.1 px residual / 0.005 step
    def undistortPoint(self, y):
        y = np.asarray(y, dtype=np.float64)
        if y.ndim > 1:
            return np.stack([self.undistortPoint(v) for v in y.reshape(-1, 2)], 0).reshape(y.shape[:-1] + (3,))
        result = PinholeCamera.undistortPoint(self, y)
        for _ in range(30):
            res = y - self.projectPoint(result)
            if np.sqrt(res @ res) < 0.1:
                break
            J = self.projectionJacobian(result)
            H = J.T @ J + np.eye(3) * 1000.0
            step = np.linalg.solve(H, J.T @ res)
            result = normalized(result + step)
            if np.sqrt(step @ step) < 0.005:
                break
        return result

    # EquidistantCamera.cpp:83-119
    def projectionJacobian(self, p):
        p = np.asarray(p, dtype=np.float64)
        if p.ndim > 1:
            return np.stack([self.projectionJacobian(v) for v in p.reshape(-1, 3)], 0).reshape(p.shape[:-1] + (2, 3))
        x, yy, z = p
        Jh = np.array([[1.0 / z, 0.0, -1.0 * x / (z * z)], [0.0, 1.0 / z, -1.0 * yy / (z * z)]])
        h = np.array([x / z, yy / z])
        D = np.eye(2)
        r = np.sqrt(h @ h)
        if r > 1e-6:
            d = self.dist
            theta = np.arctan(r)
            temp = 1.0 + d[0] * theta**2 + d[1] * theta**4 + d[2] * theta**6 + d[3] * theta**8
            D = temp * theta / r * np.eye(2)
            Dr = h / r
            Dth = Dr / (1.0 + r * r)
            DTemp = temp / r
            for i in range(1, 5):
                DTemp += theta / r * d[i - 1] * (2 * i) * theta ** (2 * i - 1)
            D = D + np.outer(h, Dth) * DTemp
            D = D + (-theta / (r * r) * temp) * np.outer(h, Dr)
        K2 = np.array([[self.fx, 0.0], [0.0, self.fy]])
        return K2 @ D @ Jh


def createDefaultCamera():
    """test/testing_utilities.cpp:175-184."""
    return PinholeCamera(800, 480, 450.0, 450.0, 400.0, 240.0)


def simulationCamera():
    """src/dataserver/SimulationDataServer.cpp:156-171 (generatePinholeCameraSquare)."""
    return PinholeCamera(752, 480, 458.654, 457.296, 367.215, 248.375)
