// Sensor-state algebra, camera models, point charts and the per-landmark output
// block of the EqF -- shared by host and device code.
//
// Reference functions restated here (path:line relative to the reference):
//   sensorStateGroupAction            src/mathematical/VIOGroup.cpp:25-32
//   GIFT pinhole / radtan camera      external/GIFT/GIFT/src/camera/PinholeCamera.cpp:57-74,
//                                     external/GIFT/GIFT/src/camera/StandardCamera.cpp:41-111
//   stereographic sphere chart        src/mathematical/VIOState.cpp:246-307
//   pointChart_invdepth.inv           src/mathematical/VIOState.cpp:174-188
//   conv_euc2ind / conv_ind2euc       src/mathematical/coordinateSuite/invdepth.cpp:65-81
//   ind2euc (lift / C*)               src/mathematical/coordinateSuite/invdepth.cpp:203-209, 259-263
//   EqFoutputMatrixCiStar_{euclid,invdepth}, outputMatrixCi
//                                     coordinateSuite/euclid.cpp:162-184, invdepth.cpp:255-266,
//                                     src/mathematical/EqFMatrices.cpp:84-89
#pragma once
#include "lie.cuh"

namespace eqvio {

constexpr int SENSOR_DIM = 21;  // VIOSensorState::CompDim
constexpr int SOFF = 24;        // internal row offset of landmark 0 in Sigma (sensor block padded 21 -> 24)

enum { COORD_EUCLIDEAN = 0, COORD_INVDEPTH = 1, COORD_NORMAL = 2 };
enum { CAM_PINHOLE = 0, CAM_RADTAN = 1, CAM_EQUIDISTANT = 2 };

struct Camera {
    int model, width, height, ndist;
    double fx, fy, cx, cy;
    double dist[5];
    double inv_dist[5];
};

struct SensorState {  // VIOSensorState (VIOState.h:41-62)
    double bias[6];
    SE3 pose;
    V3 vel;
    SE3 cam;
};
struct GroupSensor {  // sensor part of VIOGroup (VIOGroup.h:32-38)
    double beta[6];
    SE3 A;
    V3 w;
    SE3 B;
};

HD SE3 unpack_se3(const double* f) { return SE3{Quat{f[0], f[1], f[2], f[3]}, V3{f[4], f[5], f[6]}}; }
HD void pack_se3(const SE3& s, double* f) {
    f[0] = s.q.w; f[1] = s.q.x; f[2] = s.q.y; f[3] = s.q.z; f[4] = s.x.x; f[5] = s.x.y; f[6] = s.x.z;
}
HD SensorState unpack_sensor(const double* f) {
    SensorState s;
    for (int i = 0; i < 6; ++i) s.bias[i] = f[i];
    s.pose = unpack_se3(f + 6);
    s.vel = V3{f[13], f[14], f[15]};
    s.cam = unpack_se3(f + 16);
    return s;
}
HD void pack_sensor(const SensorState& s, double* f) {
    for (int i = 0; i < 6; ++i) f[i] = s.bias[i];
    pack_se3(s.pose, f + 6);
    f[13] = s.vel.x; f[14] = s.vel.y; f[15] = s.vel.z;
    pack_se3(s.cam, f + 16);
}
HD GroupSensor unpack_group(const double* f) {
    GroupSensor g;
    for (int i = 0; i < 6; ++i) g.beta[i] = f[i];
    g.A = unpack_se3(f + 6);
    g.w = V3{f[13], f[14], f[15]};
    g.B = unpack_se3(f + 16);
    return g;
}
HD void pack_group(const GroupSensor& g, double* f) {
    for (int i = 0; i < 6; ++i) f[i] = g.beta[i];
    pack_se3(g.A, f + 6);
    f[13] = g.w.x; f[14] = g.w.y; f[15] = g.w.z;
    pack_se3(g.B, f + 16);
}
HD GroupSensor group_identity() {
    GroupSensor g;
    for (int i = 0; i < 6; ++i) g.beta[i] = 0;
    g.A = se3_identity();
    g.w = V3{0, 0, 0};
    g.B = se3_identity();
    return g;
}

// VIOGroup.cpp:25-32
HD SensorState sensor_group_action(const GroupSensor& X, const SensorState& s) {
    SensorState r;
    for (int i = 0; i < 6; ++i) r.bias[i] = s.bias[i] + X.beta[i];
    r.pose = se3_mul(s.pose, X.A);
    r.vel = qrot(qinv(X.A.q), s.vel - X.w);
    r.cam = se3_mul(se3_mul(se3_inv(X.A), s.cam), X.B);
    return r;
}
// landmark part of stateGroupAction: Q^-1 q0 = (1/a) R_Q^-1 q0 (VIOGroup.cpp:44-52, SOT3.h:95-103)
HD V3 landmark_action(Quat Qq, double Qa, V3 q0) { return (1.0 / Qa) * qrot(qinv(Qq), q0); }

// ---------------------------------------------------------------- cameras
HD void distort_homogeneous(double x, double y, const double* d, int n, double& ox, double& oy) {
    double r2 = x * x + y * y;  // StandardCamera.cpp:57-75
    ox = x;
    oy = y;
    if (n >= 2) {
        ox += x * (d[0] * r2 + d[1] * r2 * r2);
        oy += y * (d[0] * r2 + d[1] * r2 * r2);
    }
    if (n >= 4) {
        ox += 2 * d[2] * x * y + d[3] * (r2 + 2 * x * x);
        oy += 2 * d[3] * x * y + d[2] * (r2 + 2 * y * y);
    }
    if (n >= 5) {
        ox += x * d[4] * r2 * r2 * r2;
        oy += y * d[4] * r2 * r2 * r2;
    }
}
// EquidistantCamera::distortHomogeneousPoint (external/GIFT/GIFT/src/camera/EquidistantCamera.cpp:70-81)
HD void distort_equidistant(double x, double y, const double* d, double& ox, double& oy) {
    const double r = sqrt(x * x + y * y);
    const double th = atan(r);
    const double t2 = th * th;
    const double temp = th * (1.0 + d[0] * t2 + d[1] * t2 * t2 + d[2] * t2 * t2 * t2 + d[3] * t2 * t2 * t2 * t2);
    const double scale = (r > 1e-6) ? temp / r : 1.0;
    ox = scale * x;
    oy = scale * y;
}
HD void cam_project(const Camera& c, V3 p, double& u, double& v) {
    if (c.model == CAM_EQUIDISTANT) {  // EquidistantCamera.cpp:38-46
        double dx, dy;
        distort_equidistant(p.x / p.z, p.y / p.z, c.dist, dx, dy);
        u = c.fx * dx / 1.0 + c.cx;
        v = c.fy * dy / 1.0 + c.cy;
    } else if (c.model == CAM_RADTAN) {  // StandardCamera.cpp:41-48
        double dx, dy;
        distort_homogeneous(p.x / p.z, p.y / p.z, c.dist, c.ndist, dx, dy);
        u = c.fx * dx / 1.0 + c.cx;
        v = c.fy * dy / 1.0 + c.cy;
    } else {  // PinholeCamera.cpp:70-74
        u = c.fx * p.x / p.z + c.cx;
        v = c.fy * p.y / p.z + c.cy;
    }
}
// J: 2x3 row-major
HD void cam_jacobian(const Camera& c, V3 p, double* J) {
    double iz = 1.0 / p.z;
    if (c.model == CAM_EQUIDISTANT) {  // EquidistantCamera.cpp:83-119
        const double Jh[6] = {1.0 / p.z, 0, -1.0 * p.x / (p.z * p.z), 0, 1.0 / p.z, -1.0 * p.y / (p.z * p.z)};
        const double hx = p.x / p.z, hy = p.y / p.z;
        double D00 = 1, D01 = 0, D10 = 0, D11 = 1;
        const double r = sqrt(hx * hx + hy * hy);
        if (r > 1e-6) {
            const double* d = c.dist;
            const double th = atan(r), t2 = th * th;
            const double temp = 1.0 + d[0] * t2 + d[1] * t2 * t2 + d[2] * t2 * t2 * t2 + d[3] * t2 * t2 * t2 * t2;
            D00 = D11 = temp * th / r;
            D01 = D10 = 0;
            const double Drx = hx / r, Dry = hy / r;
            const double Dthx = Drx / (1.0 + r * r), Dthy = Dry / (1.0 + r * r);
            double DTemp = temp / r;
            double tp = th;  // theta^(2i-1)
            for (int i = 1; i < 5; ++i) {
                DTemp += th / r * d[i - 1] * (2 * i) * tp;
                tp *= t2;
            }
            D00 += hx * DTemp * Dthx; D01 += hx * DTemp * Dthy; D10 += hy * DTemp * Dthx; D11 += hy * DTemp * Dthy;
            const double k = -th / (r * r) * temp;
            D00 += k * hx * Drx; D01 += k * hx * Dry; D10 += k * hy * Drx; D11 += k * hy * Dry;
        }
        for (int j = 0; j < 3; ++j) {
            J[j] = c.fx * (D00 * Jh[j] + D01 * Jh[3 + j]);
            J[3 + j] = c.fy * (D10 * Jh[j] + D11 * Jh[3 + j]);
        }
    } else if (c.model == CAM_RADTAN) {  // StandardCamera.cpp:77-111
        double Jh[6] = {1.0 / p.z, 0, -1.0 * p.x / (p.z * p.z), 0, 1.0 / p.z, -1.0 * p.y / (p.z * p.z)};
        double px = p.x / p.z, py = p.y / p.z;
        double r2 = px * px + py * py;
        const double* d = c.dist;
        double D00 = 1, D01 = 0, D10 = 0, D11 = 1;
        if (c.ndist >= 2) {
            double s = d[0] * r2 + d[1] * r2 * r2;
            D00 += s;
            D11 += s;
            double k = d[0] + 2 * r2 * d[1];
            D00 += px * k * 2 * px; D01 += px * k * 2 * py; D10 += py * k * 2 * px; D11 += py * k * 2 * py;
        }
        if (c.ndist >= 4) {
            D00 += 2.0 * d[2] * py + 6.0 * d[3] * px;
            D01 += 2.0 * d[2] * px + 2.0 * d[3] * py;
            D10 += 2.0 * d[2] * px + 2.0 * d[3] * py;
            D11 += 6.0 * d[2] * py + 2.0 * d[3] * px;
        }
        if (c.ndist >= 5) {
            double s = d[4] * r2 * r2 * r2;
            D00 += s;
            D11 += s;
            double k = d[4] * 3 * r2 * r2;
            D00 += px * k * 2 * px; D01 += px * k * 2 * py; D10 += py * k * 2 * px; D11 += py * k * 2 * py;
        }
        for (int j = 0; j < 3; ++j) {
            J[j] = c.fx * (D00 * Jh[j] + D01 * Jh[3 + j]);
            J[3 + j] = c.fy * (D10 * Jh[j] + D11 * Jh[3 + j]);
        }
    } else {  // PinholeCamera.cpp:63-68
        J[0] = c.fx / p.z; J[1] = 0; J[2] = -c.fx * p.x / (p.z * p.z);
        J[3] = 0; J[4] = c.fy / p.z; J[5] = -c.fy * p.y / (p.z * p.z);
    }
    (void)iz;
}

HD V3 cam_undistort(const Camera& c, double u, double v) {
    V3 b = normalized(V3{(u - c.cx) / c.fx, (v - c.cy) / c.fy, 1.0});  // PinholeCamera.cpp:57-61
    if (c.model == CAM_RADTAN) {                                        // StandardCamera.cpp:50-56
        double dx, dy;
        // invDist ALWAYS has five entries (the least-squares fit of StandardCamera.cpp:117-147 solves for five whatever the length of
        // the forward model), so a four-coefficient camera -- EuRoC's sensor.yaml -- still undistorts with the r^6 term
        distort_homogeneous(b.x / b.z, b.y / b.z, c.inv_dist, 5, dx, dy);
        b = normalized(V3{dx, dy, 1.0});
    } else if (c.model == CAM_EQUIDISTANT) {  // EquidistantCamera.cpp:48-68: damped Gauss-Newton on the sphere
        for (int iter = 0; iter < 30; ++iter) {
            double eu, ev;
            cam_project(c, b, eu, ev);
            const double r0 = u - eu, r1 = v - ev;
            if (sqrt(r0 * r0 + r1 * r1) < 0.1) break;
            double J[6];
            cam_jacobian(c, b, J);
            M3 H;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) H(i, j) = J[i] * J[j] + J[3 + i] * J[3 + j] + (i == j ? 1000.0 : 0.0);
            const V3 g = V3{J[0] * r0 + J[3] * r1, J[1] * r0 + J[4] * r1, J[2] * r0 + J[5] * r1};
            const V3 step = inverse(H) * g;
            b = normalized(b + step);
            if (norm(step) < 0.005) break;
        }
    }
    return b;
}

// ---------------------------------------------------------------- stereographic chart about a pole
// rot = SO3FromVectors(-pole, e3)  (VIOState.cpp:286-307)
HD Quat stereo_rot(V3 pole) { return quat_from_two_vectors(-pole, V3{0, 0, 1}); }
// chartDiff0(pole): 2x3 row-major  = e3ProjectSphereDiff(rot*pole) * R(rot)
HD void stereo_diff0(V3 pole, double* D) {
    Quat rot = stereo_rot(pole);
    V3 eta = qrot(rot, pole);
    M3 R = qmat(rot);
    double s = 1.0 - eta.z;
    double f = 1.0 / (s * s);  // pow(1 - e3.eta, -2)
    // I23 * (I*(1-eta.z) + (eta - e3) e3^T): rows 0,1: [s,0,eta.x],[0,s,eta.y]
    double P[6] = {f * s, 0, f * eta.x, 0, f * s, f * eta.y};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) D[3 * i + j] = P[3 * i] * R(0, j) + P[3 * i + 1] * R(1, j) + P[3 * i + 2] * R(2, j);
}
// chartInvDiff0(pole): 3x2 row-major = R(rot^-1) * e3ProjectSphereInvDiff(0),  e3ProjectSphereInvDiff(0) = 2*[I2;0]
HD void stereo_inv_diff0(V3 pole, double* D) {
    M3 Ri = qmat(qinv(stereo_rot(pole)));
    for (int i = 0; i < 3; ++i) {
        D[2 * i] = 2.0 * Ri(i, 0);
        D[2 * i + 1] = 2.0 * Ri(i, 1);
    }
}
// sphereChart_stereo.inv(y, pole)
HD V3 stereo_inv(double y0, double y1, V3 pole) {
    double k = 2.0 / (y0 * y0 + y1 * y1 + 1.0);
    V3 eta = V3{k * y0, k * y1, 1.0 + k * (0.0 - 1.0)};
    return qrot(qinv(stereo_rot(pole)), eta);
}
// sphereChart_stereo(eta, pole) (VIOState.cpp:286-290) with e3ProjectSphere (:246-251)
HD void stereo_chart(V3 eta, V3 pole, double& y0, double& y1) {
    V3 r = qrot(stereo_rot(pole), eta);
    y0 = (r.x - 0.0) / (1.0 - r.z);
    y1 = (r.y - 0.0) / (1.0 - r.z);
}
// pointChart_invdepth (VIOState.cpp:160-173)
HD V3 invdepth_chart(V3 p, V3 p0) {
    double rho = 1.0 / norm(p), rho0 = 1.0 / norm(p0);
    double a, b;
    stereo_chart(p * rho, p0 * rho0, a, b);
    return V3{a, b, rho - rho0};
}
// sensor part of VIOGroup::inverse (VIOGroup.cpp:108-121)
HD GroupSensor group_inverse(const GroupSensor& X) {
    GroupSensor r;
    for (int i = 0; i < 6; ++i) r.beta[i] = -X.beta[i];
    r.A = se3_inv(X.A);
    r.B = se3_inv(X.B);
    r.w = -qrot(qinv(X.A.q), X.w);
    return r;
}
// sensorChart_std (VIOState.cpp:104-113): 21 coordinates of Xi about Xi0
HD void sensor_chart_std(const SensorState& Xi, const SensorState& Xi0, double* eps) {
    for (int i = 0; i < 6; ++i) eps[i] = Xi.bias[i] - Xi0.bias[i];
    se3_log(se3_mul(se3_inv(Xi0.pose), Xi.pose), eps + 6);
    eps[12] = Xi.vel.x - Xi0.vel.x;
    eps[13] = Xi.vel.y - Xi0.vel.y;
    eps[14] = Xi.vel.z - Xi0.vel.z;
    se3_log(se3_mul(se3_inv(Xi0.cam), Xi.cam), eps + 15);
}
HD V3 invdepth_chart_inv(V3 eps, V3 q0);
HD V3 point_chart_normal(V3 p, V3 p0);
HD V3 point_chart_normal_inv(V3 eps, V3 p0);
HD void sensor_chart_normal(const SensorState& Xi, const SensorState& Xi0, double* eps);
HD SensorState sensor_chart_normal_inv(const double* eps, const SensorState& Xi0);
// sensorChart_std inverse (VIOState.cpp:114-121)
HD SensorState sensor_chart_std_inv(const double* eps, const SensorState& Xi0) {
    SensorState Xi;
    for (int i = 0; i < 6; ++i) Xi.bias[i] = Xi0.bias[i] + eps[i];
    Xi.pose = se3_mul(Xi0.pose, se3_exp(V3{eps[6], eps[7], eps[8]}, V3{eps[9], eps[10], eps[11]}));
    Xi.vel = V3{Xi0.vel.x + eps[12], Xi0.vel.y + eps[13], Xi0.vel.z + eps[14]};
    Xi.cam = se3_mul(Xi0.cam, se3_exp(V3{eps[15], eps[16], eps[17]}, V3{eps[18], eps[19], eps[20]}));
    return Xi;
}
// VIOGroup product, sensor part (VIOGroup.cpp:71-92)
HD GroupSensor group_mul(const GroupSensor& a, const GroupSensor& b) {
    GroupSensor r;
    for (int i = 0; i < 6; ++i) r.beta[i] = a.beta[i] + b.beta[i];
    r.A = se3_mul(a.A, b.A);
    r.B = se3_mul(a.B, b.B);
    r.w = a.w + qrot(a.A.q, b.w);
    return r;
}
// liftVelocityDiscrete, sensor part (VIOGroup.cpp:229-257): the lift of one IMU segment u = (gyr3, acc3, gyrBiasVel3,
// accBiasVel3) of length dt at the state xh, and the camera-frame change the landmark part needs (:259)
HD void lift_velocity_discrete_sensor(const SensorState& xh, const double* u, double dt, GroupSensor& L, SE3& camChangeInv) {
    V3 gyr = V3{u[0] - xh.bias[0], u[1] - xh.bias[1], u[2] - xh.bias[2]};
    V3 acc = V3{u[3] - xh.bias[3], u[4] - xh.bias[4], u[5] - xh.bias[5]};
    V3 gdir = qrot(qinv(xh.pose.q), V3{0, 0, 1});
    for (int i = 0; i < 6; ++i) L.beta[i] = dt * u[6 + i];
    L.A.q = so3_exp(dt * gyr);
    V3 x = dt * qrot(xh.pose.q, xh.vel) + (0.5 * dt * dt) * (qrot(xh.pose.q, acc) + V3{0, 0, -GRAVITY_CONSTANT});
    L.A.x = qrot(qinv(xh.pose.q), x);
    L.B = se3_mul(se3_mul(se3_inv(xh.cam), L.A), xh.cam);
    V3 bvd = acc - GRAVITY_CONSTANT * gdir;
    L.w = xh.vel - (xh.vel + dt * bvd);
    camChangeInv = se3_mul(se3_mul(se3_inv(xh.cam), se3_inv(L.A)), xh.cam);
}
// a0Discrete of stateMatrixADiscrete (EqFMatrices.cpp:27-37), sensor part: the chart coordinates eps1 of
// (X LambdaTilde X^-1) . chart^-1(eps) for the sensor coordinates eps (21); also returns the camera-frame change of
// Lambda(xi) that the landmark part of the same evaluation needs.
HD void a0_discrete_sensor(int coord, const GroupSensor& X, const SensorState& xi0, const double* u, double dt, const double* eps,
                           double* eps1, SE3& camChangeInv) {
    const SensorState xe = coord == COORD_NORMAL ? sensor_chart_normal_inv(eps, xi0) : sensor_chart_std_inv(eps, xi0);
    const SensorState xhat = sensor_group_action(X, xi0);
    const SensorState xi = sensor_group_action(X, xe);
    GroupSensor L1, L0;
    SE3 cc0;
    lift_velocity_discrete_sensor(xi, u, dt, L1, camChangeInv);
    lift_velocity_discrete_sensor(xhat, u, dt, L0, cc0);
    const GroupSensor G = group_mul(group_mul(X, group_mul(L1, group_inverse(L0))), group_inverse(X));
    if (coord == COORD_NORMAL)
        sensor_chart_normal(sensor_group_action(G, xe), xi0, eps1);
    else
        sensor_chart_std(sensor_group_action(G, xe), xi0, eps1);
}
// ... landmark part: eps (3) are the chart coordinates of landmark (p0; Q, a), cc1 / cc0 the camera-frame changes of
// Lambda(xi) / Lambda(xi_hat).  Point chart: Euclidean, inverse depth or normal.
HD V3 a0_discrete_landmark(int coord, V3 p0, Quat Q, double a, V3 eps, const SE3& cc1, const SE3& cc0) {
    const V3 pe = coord == COORD_INVDEPTH ? invdepth_chart_inv(eps, p0) : (coord == COORD_NORMAL ? point_chart_normal_inv(eps, p0) : p0 + eps);
    const V3 p = landmark_action(Q, a, pe), phat = landmark_action(Q, a, p0);
    const V3 p1 = se3_apply(cc1, p), ph1 = se3_apply(cc0, phat);
    const Quat R1 = quat_from_two_vectors(normalized(p1), normalized(p)), R0 = quat_from_two_vectors(normalized(ph1), normalized(phat));
    const double a1 = norm(p) / norm(p1), a0 = norm(phat) / norm(ph1);
    const Quat Rt = qmul(R1, qinv(R0));          // LambdaTilde = Lambda(xi) Lambda(xi_hat)^-1
    const double at = a1 * (1.0 / a0);
    const Quat Rg = qmul(qmul(Q, Rt), qinv(Q));  // X LambdaTilde X^-1
    const double ag = (a * at) * (1.0 / a);
    const V3 pe1 = landmark_action(Rg, ag, pe);
    return coord == COORD_INVDEPTH ? invdepth_chart(pe1, p0) : (coord == COORD_NORMAL ? point_chart_normal(pe1, p0) : pe1 - p0);
}
// integrateSystemFunction, sensor part (VIOState.cpp:27-68): advances s by one IMU segment u = (gyr, acc, gyrBiasVel,
// accBiasVel) of length dt and returns the camera-frame change applied to every landmark.
HD SE3 integrate_system_sensor(SensorState& s, const double* u, double dt) {
    V3 gyr = V3{u[0] - s.bias[0], u[1] - s.bias[1], u[2] - s.bias[2]};
    V3 acc = V3{u[3] - s.bias[3], u[4] - s.bias[4], u[5] - s.bias[5]};
    SensorState n;
    for (int i = 0; i < 6; ++i) n.bias[i] = s.bias[i] + dt * u[6 + i];
    SE3 poseChange;
    poseChange.q = so3_exp(dt * gyr);
    V3 x = dt * qrot(s.pose.q, s.vel) + (0.5 * dt * dt) * (qrot(s.pose.q, acc) + V3{0, 0, -GRAVITY_CONSTANT});
    poseChange.x = qrot(qinv(s.pose.q), x);
    n.pose = se3_mul(s.pose, poseChange);
    V3 inertialVelocityDiff = qmat(s.pose.q) * acc + V3{0, 0, -GRAVITY_CONSTANT};
    n.vel = qrot(qinv(n.pose.q), qrot(s.pose.q, s.vel) + dt * inertialVelocityDiff);
    n.cam = s.cam;
    SE3 camChangeInv = se3_mul(se3_mul(se3_inv(s.cam), se3_inv(poseChange)), s.cam);
    s = n;
    return camChangeInv;
}

// pointChart_invdepth.inv (VIOState.cpp:174-188)
HD V3 invdepth_chart_inv(V3 eps, V3 q0) {
    double rho0 = 1.0 / norm(q0);
    V3 y0 = q0 * rho0;
    V3 y = stereo_inv(eps.x, eps.y, y0);
    double rho = eps.z + rho0;
    if (rho <= 0.0) rho = 1e-6;
    return y / rho;
}
// conv_euc2ind(q0) (invdepth.cpp:65-73): rows 0,1 = rho0 * DPhi(y0) (I - y0 y0^T); row 2 = -rho0^2 y0^T
HD M3 conv_euc2ind(V3 q0) {
    double rho0 = 1.0 / norm(q0);
    V3 y0 = q0 * rho0;
    double D[6];
    stereo_diff0(y0, D);
    M3 P = m3_identity() - outer(y0, y0);
    M3 M;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) M(i, j) = rho0 * (D[3 * i] * P(0, j) + D[3 * i + 1] * P(1, j) + D[3 * i + 2] * P(2, j));
    M(2, 0) = -rho0 * rho0 * y0.x;
    M(2, 1) = -rho0 * rho0 * y0.y;
    M(2, 2) = -rho0 * rho0 * y0.z;
    return M;
}
// conv_ind2euc(q0) (invdepth.cpp:74-81): cols 0,1 = DPhi^-1(y0)/rho0; col 2 = -y0/rho0^2
HD M3 conv_ind2euc(V3 q0) {
    double rho0 = 1.0 / norm(q0);
    V3 y0 = q0 * rho0;
    double D[6];
    stereo_inv_diff0(y0, D);
    M3 M;
    for (int i = 0; i < 3; ++i) {
        M(i, 0) = D[2 * i] / rho0;
        M(i, 1) = D[2 * i + 1] / rho0;
    }
    M(0, 2) = -y0.x / (rho0 * rho0);
    M(1, 2) = -y0.y / (rho0 * rho0);
    M(2, 2) = -y0.z / (rho0 * rho0);
    return M;
}
// ind2euc of the innovation lift and of C* (invdepth.cpp:203-209, 259-263): [r0 * DPhi^-1(y0), -r0 * q0]
HD M3 ind2euc_lift(V3 q0) {
    double r0 = norm(q0);
    V3 y0 = q0 / r0;
    double D[6];
    stereo_inv_diff0(y0, D);
    M3 M;
    for (int i = 0; i < 3; ++i) {
        M(i, 0) = r0 * D[2 * i];
        M(i, 1) = r0 * D[2 * i + 1];
    }
    M(0, 2) = -r0 * q0.x;
    M(1, 2) = -r0 * q0.y;
    M(2, 2) = -r0 * q0.z;
    return M;
}

// ---------------------------------------------------------------- output block C*_i (2x3 row-major)
// DRho(v)[:, 0:3] = J_pi(v) * skew(v); the 4th column of DRho is zero, so only the rotation block
// of Ad_{Q^-1} and the first three rows of m2g contribute (euclid.cpp:166-183).
HD void drho3(const Camera& cam, V3 v, double* out /*2x3*/) {
    double J[6];
    cam_jacobian(cam, v, J);
    M3 S = skew(v);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) out[3 * i + j] = J[3 * i] * S(0, j) + J[3 * i + 1] * S(1, j) + J[3 * i + 2] * S(2, j);
}
// useStar: y = measured pixel; otherwise y := project(qHat) (outputMatrixCi, EqFMatrices.cpp:84-89)
// ------------------------------------------------------------------------------------------------
// Normal coordinates (coordinateSuite/normal.cpp, VIOState.cpp:123-152,190-215,310-353)
// ------------------------------------------------------------------------------------------------
HD void sphere_chart_normal(V3 eta, V3 pole, double& e0, double& e1) {  // VIOState.cpp:310-327
    const V3 y = qrot(quat_from_two_vectors(pole, V3{0, 0, 1}), eta);
    const V3 ye3 = cross(y, V3{0, 0, 1});
    const double sin_th = norm(ye3), cos_th = y.z;
    const double th = atan2(sin_th, cos_th);
    const double scale = fabs(th) < 1e-8 ? 1.0 : th / sin_th;
    e0 = ye3.x * scale;
    e1 = ye3.y * scale;
}
HD V3 sphere_chart_normal_inv(double e0, double e1, V3 pole) {  // VIOState.cpp:328-337
    const V3 y = qrot(so3_exp(V3{-e0, -e1, -0.0}), V3{0, 0, 1});
    return qrot(qinv(quat_from_two_vectors(pole, V3{0, 0, 1})), y);
}
HD void sphere_normal_inv_diff0(V3 pole, double* D /*3x2 row-major*/) {  // VIOState.cpp:346-353
    const M3 Ri = qmat(qinv(quat_from_two_vectors(pole, V3{0, 0, 1})));
    for (int r = 0; r < 3; ++r) {
        D[2 * r] = Ri(r, 1);
        D[2 * r + 1] = -Ri(r, 0);
    }
}
HD V3 point_chart_normal(V3 p, V3 p0) {  // VIOState.cpp:190-203
    const double rho = 1.0 / norm(p), rho0 = 1.0 / norm(p0);
    double e0, e1;
    sphere_chart_normal(rho * p, rho0 * p0, e0, e1);
    return V3{e0, e1, log(rho / rho0)};
}
HD V3 point_chart_normal_inv(V3 eps, V3 p0) {  // VIOState.cpp:204-215
    const double rho0 = 1.0 / norm(p0);
    const V3 y = sphere_chart_normal_inv(eps.x, eps.y, rho0 * p0);
    return y / (rho0 * exp(eps.z));
}
HD void se23_log(Quat q, V3 x0, V3 x1, double* out /*9*/) {  // SEn3.h:94-113
    const V3 om = so3_log(q);
    const M3 O = skew(om);
    const double theta = sqrt(om.x * om.x + om.y * om.y + om.z * om.z);
    double coef = 1.0 / 12.0;
    if (fabs(theta) > 1e-8) coef = 1.0 / (theta * theta) * (1.0 - (theta * sin(theta)) / (2.0 * (1.0 - cos(theta))));
    const M3 VInv = m3_identity() - 0.5 * O + coef * (O * O);
    const V3 a = VInv * x0, b = VInv * x1;
    out[0] = om.x; out[1] = om.y; out[2] = om.z;
    out[3] = a.x; out[4] = a.y; out[5] = a.z;
    out[6] = b.x; out[7] = b.y; out[8] = b.z;
}
HD void sensor_chart_normal(const SensorState& Xi, const SensorState& Xi0, double* eps) {  // VIOState.cpp:123-137
    const SE3 A = se3_mul(se3_inv(Xi0.pose), Xi.pose);
    const V3 v_xi0 = qrot(Xi0.pose.q, Xi0.vel), v_xi = qrot(Xi.pose.q, Xi.vel);
    const V3 v_A = qrot(qinv(Xi0.pose.q), v_xi - v_xi0);
    const SE3 B = se3_mul(se3_mul(se3_inv(Xi0.cam), A), Xi.cam);
    for (int i = 0; i < 6; ++i) eps[i] = Xi.bias[i] - Xi0.bias[i];
    se23_log(A.q, A.x, v_A, eps + 6);
    se3_log(B, eps + 15);
}
HD SensorState sensor_chart_normal_inv(const double* eps, const SensorState& Xi0) {  // VIOState.cpp:138-152
    Quat q;
    V3 x0, x1;
    se23_exp(V3{eps[6], eps[7], eps[8]}, V3{eps[9], eps[10], eps[11]}, V3{eps[12], eps[13], eps[14]}, q, x0, x1);
    const SE3 B = se3_exp(V3{eps[15], eps[16], eps[17]}, V3{eps[18], eps[19], eps[20]});
    const SE3 A = SE3{q, x0};
    SensorState Xi;
    for (int i = 0; i < 6; ++i) Xi.bias[i] = Xi0.bias[i] + eps[i];
    Xi.pose = se3_mul(Xi0.pose, A);
    const V3 v_xi0 = qrot(Xi0.pose.q, Xi0.vel);
    Xi.vel = qrot(qinv(Xi.pose.q), v_xi0 + qrot(Xi0.pose.q, x1));
    Xi.cam = se3_mul(se3_mul(se3_inv(A), Xi0.cam), B);
    return Xi;
}
HD double normal_diff_step() { return cbrt(2.220446049250313e-16); }  // numericalDifferential's default h (Geometry.cpp:27-29)
// landmark block of coordinateDifferential_normal_euclid (VIOState.cpp:391-401): central differences of
// eps -> pointChart_normal(pointChart_euclid^-1(eps)); M row-major 3x3
HD M3 normal_M_landmark(V3 p0) {
    const double h = normal_diff_step();
    M3 M;
    for (int j = 0; j < 3; ++j) {
        const V3 d = V3{j == 0 ? h : 0.0, j == 1 ? h : 0.0, j == 2 ? h : 0.0};
        const V3 a = point_chart_normal(p0 + d, p0), b = point_chart_normal(p0 - d, p0);
        M.m[0 * 3 + j] = (a.x - b.x) / (2 * h);
        M.m[1 * 3 + j] = (a.y - b.y) / (2 * h);
        M.m[2 * 3 + j] = (a.z - b.z) / (2 * h);
    }
    return M;
}

HD void output_block(const Camera& cam, int coord, V3 q0, Quat Qq, double Qa, bool useStar, double yu, double yv,
                     double* C /*2x3*/) {
    if (coord == COORD_NORMAL) {  // normal.cpp:57-65: [J_pi(yHat) R_Q^T chartInvDiff0(q0), 0]; the measurement is not used
        const V3 yHatN = qrot(qinv(Qq), normalized(q0));
        double J[6], D[6];
        cam_jacobian(cam, yHatN, J);
        sphere_normal_inv_diff0(q0, D);  // the reference passes q0, not its bearing, as the pole (normalised inside)
        const M3 Rt = qmat(qinv(Qq));
        for (int i = 0; i < 2; ++i) {
            double JR[3];
            for (int k = 0; k < 3; ++k) JR[k] = J[3 * i] * Rt(0, k) + J[3 * i + 1] * Rt(1, k) + J[3 * i + 2] * Rt(2, k);
            for (int j = 0; j < 2; ++j) C[3 * i + j] = JR[0] * D[j] + JR[1] * D[2 + j] + JR[2] * D[4 + j];
            C[3 * i + 2] = 0.0;
        }
        return;
    }
    V3 qHat = landmark_action(Qq, Qa, q0);
    V3 yHat = normalized(qHat);
    if (!useStar) cam_project(cam, qHat, yu, yv);
    V3 yTru = cam_undistort(cam, yu, yv);
    double Da[6], Db[6];
    drho3(cam, yTru, Da);
    drho3(cam, yHat, Db);
    M3 Rinv = qmat(qinv(Qq));
    double n2 = norm2(q0);
    M3 m2g = (-1.0 / n2) * skew(q0);  // -skew(q0)/|q0|^2
    M3 T = Rinv * m2g;
    double Ce[6];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) {
            double a0 = 0.5 * (Da[3 * i] + Db[3 * i]), a1 = 0.5 * (Da[3 * i + 1] + Db[3 * i + 1]),
                   a2 = 0.5 * (Da[3 * i + 2] + Db[3 * i + 2]);
            Ce[3 * i + j] = a0 * T(0, j) + a1 * T(1, j) + a2 * T(2, j);
        }
    if (coord == COORD_INVDEPTH) {
        M3 L = ind2euc_lift(q0);
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 3; ++j) C[3 * i + j] = Ce[3 * i] * L(0, j) + Ce[3 * i + 1] * L(1, j) + Ce[3 * i + 2] * L(2, j);
    } else {
        for (int i = 0; i < 6; ++i) C[i] = Ce[i];
    }
}

}  // namespace eqvio
