// Small fixed-size Lie-group arithmetic shared by host and device code.
//
// Restates, in plain structs, the arithmetic the reference takes from LiePP
// (external/LiePP/include/liepp/{SO3,SE3,SOT3,SEn3}.h) and from Eigen's
// Quaternion (product, inverse = conj/|q|^2, _transformVector,
// toRotationMatrix, matrix->quaternion, setFromTwoVectors).  Rotations are
// quaternions (w,x,y,z) that are never renormalised, like LiePP's SO3.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace eqvio {

constexpr double GRAVITY_CONSTANT = 9.80665;  // include/eqvio/mathematical/IMUVelocity.h:26

struct V3 {
    double x, y, z;
};
struct Quat {
    double w, x, y, z;
};
struct M3 {
    double m[9];  // row-major
    HD double& operator()(int r, int c) { return m[3 * r + c]; }
    HD double operator()(int r, int c) const { return m[3 * r + c]; }
};
struct SE3 {
    Quat q;
    V3 x;
};

HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
HD V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
HD V3 operator*(V3 a, double s) { return V3{s * a.x, s * a.y, s * a.z}; }
HD V3 operator/(V3 a, double s) { return V3{a.x / s, a.y / s, a.z / s}; }
HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
HD double norm2(V3 a) { return dot(a, a); }
HD double norm(V3 a) { return sqrt(dot(a, a)); }
HD V3 normalized(V3 a) {
    double n2 = dot(a, a);
    return n2 > 0 ? a / sqrt(n2) : a;
}
HD double get(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

HD M3 m3_identity() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
HD M3 m3_zero() { return M3{{0, 0, 0, 0, 0, 0, 0, 0, 0}}; }
HD M3 skew(V3 v) { return M3{{0, -v.z, v.y, v.z, 0, -v.x, -v.y, v.x, 0}}; }  // SO3.h:33-35
HD M3 operator*(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
    return r;
}
HD V3 operator*(const M3& a, V3 v) {
    return V3{a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
              a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
HD M3 operator+(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i];
    return r;
}
HD M3 operator-(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i];
    return r;
}
HD M3 operator*(double s, const M3& a) {
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = s * a.m[i];
    return r;
}
HD M3 transpose(const M3& a) { return M3{{a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8]}}; }
HD M3 outer(V3 a, V3 b) { return M3{{a.x * b.x, a.x * b.y, a.x * b.z, a.y * b.x, a.y * b.y, a.y * b.z, a.z * b.x, a.z * b.y, a.z * b.z}}; }
// General 3x3 inverse by cofactors (what Eigen does for fixed 3x3; euclid.cpp:153).
HD M3 inverse(const M3& a) {
    double c00 = a.m[4] * a.m[8] - a.m[5] * a.m[7];
    double c01 = a.m[5] * a.m[6] - a.m[3] * a.m[8];
    double c02 = a.m[3] * a.m[7] - a.m[4] * a.m[6];
    double det = a.m[0] * c00 + a.m[1] * c01 + a.m[2] * c02;
    double id = 1.0 / det;
    M3 r;
    r.m[0] = c00 * id;
    r.m[1] = (a.m[2] * a.m[7] - a.m[1] * a.m[8]) * id;
    r.m[2] = (a.m[1] * a.m[5] - a.m[2] * a.m[4]) * id;
    r.m[3] = c01 * id;
    r.m[4] = (a.m[0] * a.m[8] - a.m[2] * a.m[6]) * id;
    r.m[5] = (a.m[2] * a.m[3] - a.m[0] * a.m[5]) * id;
    r.m[6] = c02 * id;
    r.m[7] = (a.m[1] * a.m[6] - a.m[0] * a.m[7]) * id;
    r.m[8] = (a.m[0] * a.m[4] - a.m[1] * a.m[3]) * id;
    return r;
}

// ---------------------------------------------------------------- quaternions
HD Quat quat_identity() { return Quat{1, 0, 0, 0}; }
HD Quat qmul(Quat a, Quat b) {
    return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
HD Quat qinv(Quat q) {  // conjugate / squared norm (Eigen's quaternion inverse); ONE reciprocal instead of four fp64 divisions (~25
    // instructions each): the inverse sits on every serial chain of the update (observer segments, Riccati prologue, lifts) and the
    // divisions were a quarter of the observer chain's stall samples.  Differs from the four divisions by at most one ulp per coefficient.
    const double inv = 1.0 / (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return Quat{q.w * inv, -q.x * inv, -q.y * inv, -q.z * inv};
}
HD V3 qrot(Quat q, V3 v) {
    V3 u = V3{q.x, q.y, q.z};
    V3 uv = cross(u, v);
    uv = uv + uv;
    return v + q.w * uv + cross(u, uv);
}
HD M3 qmat(Quat q) {
    double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    return M3{{1.0 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.0 - (txx + tzz), tyz - twx, txz - twy, tyz + twx,
               1.0 - (txx + tyy)}};
}
HD Quat mat2quat(const M3& m) {
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    double q[4];  // w,x,y,z
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[0] = 0.5 * t;
        t = 0.5 / t;
        q[1] = (m(2, 1) - m(1, 2)) * t;
        q[2] = (m(0, 2) - m(2, 0)) * t;
        q[3] = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        q[1 + i] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m(k, j) - m(j, k)) * t;
        q[1 + j] = (m(j, i) + m(i, j)) * t;
        q[1 + k] = (m(k, i) + m(i, k)) * t;
    }
    return Quat{q[0], q[1], q[2], q[3]};
}
// Eigen setFromTwoVectors(a, b).  The near-antiparallel branch (c < -1+1e-12)
// of Eigen picks the rotation axis from an SVD null vector; any unit axis
// orthogonal to v0 is an equally valid half-turn, so a closed-form
// perpendicular is used there instead (measure-zero case, never reached by
// the filter: it would need a landmark to jump to the opposite bearing).
HD Quat quat_from_two_vectors(V3 a, V3 b) {
    V3 v0 = normalized(a), v1 = normalized(b);
    double c = dot(v1, v0);
    if (c < -1.0 + 1e-12) {
        c = c > -1.0 ? c : -1.0;
        V3 ax = fabs(v0.x) < 0.9 ? cross(v0, V3{1, 0, 0}) : cross(v0, V3{0, 1, 0});
        ax = normalized(ax);
        double w2 = (1.0 + c) * 0.5;
        double s = sqrt(1.0 - w2);
        return Quat{sqrt(w2), ax.x * s, ax.y * s, ax.z * s};
    }
    V3 axis = cross(v0, v1);
    double s = sqrt((1.0 + c) * 2.0);
    double invs = 1.0 / s;
    return Quat{s * 0.5, axis.x * invs, axis.y * invs, axis.z * invs};
}

// ---------------------------------------------------------------- SO3 / SE3 / SE2(3) / SOT3
HD Quat so3_exp(V3 w) {  // SO3.h:42-54
    double theta = norm(w) / 2.0;
    if (theta > 1e-6) {
        V3 n = normalized(w);
        double s = sin(theta);
        return Quat{cos(theta), s * n.x, s * n.y, s * n.z};
    }
    return Quat{1.0, w.x / 2.0, w.y / 2.0, w.z / 2.0};
}
HD V3 so3_log(Quat q) {  // SO3.h:56-63
    M3 R = qmat(q);
    double theta = acos((R(0, 0) + R(1, 1) + R(2, 2) - 1.0) / 2.0);
    double coef = (fabs(theta) > 1e-6) ? theta / (2.0 * sin(theta)) : 0.5;
    return V3{coef * (R(2, 1) - R(1, 2)), coef * (R(0, 2) - R(2, 0)), coef * (R(1, 0) - R(0, 1))};
}
HD SE3 se3_identity() { return SE3{quat_identity(), V3{0, 0, 0}}; }
HD SE3 se3_mul(const SE3& a, const SE3& b) { return SE3{qmul(a.q, b.q), a.x + qrot(a.q, b.x)}; }  // SE3.h:156
HD SE3 se3_inv(const SE3& a) {                                                                   // SE3.h:162
    Quat qi = qinv(a.q);
    return SE3{qi, -qrot(qi, a.x)};
}
HD V3 se3_apply(const SE3& a, V3 p) { return qrot(a.q, p) + a.x; }
// SE3::log (SE3.h:85-104): out = (omega, v)
HD void se3_log(const SE3& P, double* out) {
    V3 om = so3_log(P.q);
    M3 O = skew(om);
    double theta = sqrt(om.x * om.x + om.y * om.y + om.z * om.z);
    double coef = 1.0 / 12.0;
    if (fabs(theta) > 1e-6) coef = 1.0 / (theta * theta) * (1.0 - (theta * sin(theta)) / (2.0 * (1.0 - cos(theta))));
    M3 VInv = m3_identity() - 0.5 * O + coef * (O * O);
    V3 v = VInv * P.x;
    out[0] = om.x; out[1] = om.y; out[2] = om.z;
    out[3] = v.x; out[4] = v.y; out[5] = v.z;
}
HD void rodrigues(V3 w, bool strict, M3& R, M3& V) {  // SE3.h:59-77 (strict: >1e-12), SEn3.h:66-87 (>=1e-12)
    double th = norm(w);
    double A, B, C;
    bool big = strict ? (fabs(th) > 1e-12) : (fabs(th) >= 1e-12);
    if (big) {
        A = sin(th) / th;
        B = (1.0 - cos(th)) / (th * th);
        C = (1.0 - A) / (th * th);
    } else {
        A = 1.0;
        B = 0.5;
        C = 1.0 / 6.0;
    }
    M3 wx = skew(w);
    M3 wx2 = wx * wx;
    R = m3_identity() + A * wx + B * wx2;
    V = m3_identity() + B * wx + C * wx2;
}
HD SE3 se3_exp(V3 w, V3 v) {  // SE3.h:59-84
    M3 R, V;
    rodrigues(w, true, R, V);
    return SE3{mat2quat(R), V * v};
}
HD void se23_exp(V3 w, V3 v0, V3 v1, Quat& q, V3& x0, V3& x1) {  // SEn3.h:66-93
    M3 R, V;
    rodrigues(w, false, R, V);
    q = mat2quat(R);
    x0 = V * v0;
    x1 = V * v1;
}
// 6x6 matrices, row-major double[36]
HD void se3_Adjoint(const SE3& a, double* Ad) {  // SE3.h:168-176
    M3 R = qmat(a.q);
    M3 xR = skew(a.x) * R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Ad[6 * i + j] = R(i, j);
            Ad[6 * i + 3 + j] = 0.0;
            Ad[6 * (3 + i) + j] = xR(i, j);
            Ad[6 * (3 + i) + 3 + j] = R(i, j);
        }
}
HD void se3_adjoint(const double* u, double* ad) {  // SE3.h:50-58
    M3 wx = skew(V3{u[0], u[1], u[2]});
    M3 vx = skew(V3{u[3], u[4], u[5]});
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            ad[6 * i + j] = wx(i, j);
            ad[6 * i + 3 + j] = 0.0;
            ad[6 * (3 + i) + j] = vx(i, j);
            ad[6 * (3 + i) + 3 + j] = wx(i, j);
        }
}
HD void mat6_vec(const double* M, const double* v, double* r) {
    for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int j = 0; j < 6; ++j) s += M[6 * i + j] * v[j];
        r[i] = s;
    }
}
HD void mat6_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += A[6 * i + k] * B[6 * k + j];
            C[6 * i + j] = s;
        }
}

}  // namespace eqvio
