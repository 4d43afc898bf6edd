// Header-only C++ facade over the C ABI of include/eqvio_b200.h with the member names, argument
// meaning and silent-return behaviour of the reference's `VIOFilter`
// (include/eqvio/VIOFilter.h:36-192, src/VIOFilter.cpp).  The reference's own headers need Eigen,
// LiePP, GIFT/OpenCV and yaml-cpp; this facade uses plain structs instead so that it builds with
// nothing but a C++17 compiler.  INTEGRATION.md shows the thin adapter that maps the reference's
// Eigen/LiePP types onto these structs inside a re-bodied eqvio `VIOFilter`.
#pragma once
#include <array>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "eqvio_b200.h"

namespace eqvio_b200 {

struct IMUVelocity {  // include/eqvio/mathematical/IMUVelocity.h:33-84
    double stamp = 0;
    std::array<double, 3> gyr{}, acc{}, gyrBiasVel{}, accBiasVel{};
};

struct SE3 {  // liepp::SE3d: quaternion (w,x,y,z) + translation
    std::array<double, 4> q{1, 0, 0, 0};
    std::array<double, 3> x{};
};

struct Landmark {  // include/eqvio/mathematical/VIOState.h:64-67
    std::array<double, 3> p{};
    int id = -1;
};

struct VIOSensorState {  // VIOState.h:41-62
    std::array<double, 6> inputBias{};
    SE3 pose;
    std::array<double, 3> velocity{};
    SE3 cameraOffset;
};

struct VIOState {  // VIOState.h:69-90
    VIOSensorState sensor;
    std::vector<Landmark> cameraLandmarks;
    std::vector<int> getIds() const {
        std::vector<int> ids;
        ids.reserve(cameraLandmarks.size());
        for (const auto& lm : cameraLandmarks) ids.push_back(lm.id);
        return ids;
    }
    int Dim() const { return 21 + 3 * (int)cameraLandmarks.size(); }
};

struct VisionMeasurement {  // include/eqvio/mathematical/VisionMeasurement.h:35-40
    double stamp = 0;
    std::map<int, std::array<double, 2>> camCoordinates;  // ascending id, like the reference's std::map
    eqvio_camera camera{};                                 // flattened GIFT::GICamera (cameraPtr)
    std::vector<int> getIds() const {
        std::vector<int> ids;
        for (const auto& kv : camCoordinates) ids.push_back(kv.first);
        return ids;
    }
};

struct EqFStateView {  // what viewEqFState() exposes of VIO_eqf (VIO_eqf.h:34-42)
    VIOState xi0;
    std::array<double, 23> X{};   // beta6 | A(q,x) | w3 | B(q,x)
    std::vector<double> Q;        // 5 per landmark: quaternion wxyz, scale a
    std::vector<double> Sigma;    // dim x dim, column-major
    int dim = 0;
    double currentTime = -1;
};

class Error : public std::runtime_error {
public:
    Error(int code, const std::string& msg) : std::runtime_error(msg), code(code) {}
    int code;
};

class VIOFilter {
public:
    struct Settings : eqvio_settings {  // VIOFilter::Settings (VIOFilterSettings.h:58-99)
        Settings() { eqvio_settings_default(this); }
    };

    std::unique_ptr<Settings> settings;  // public in the reference too (VIOFilter.h:86)

    // VIOFilter(const Settings&) -- VIOFilter.cpp:31-41
    explicit VIOFilter(const Settings& s, int capacity = 256, int device = 0, void* stream = nullptr)
        : settings(std::make_unique<Settings>(s)) {
        check_create(eqvio_create(settings.get(), device, capacity, stream, &h_));
    }
    // VIOFilter(const VIOState& xi0, const Settings&, const double& time = 0) -- VIOFilter.cpp:43-56
    VIOFilter(const VIOState& xi0, const Settings& s, const double& time = 0, int capacity = 256, int device = 0,
              void* stream = nullptr)
        : settings(std::make_unique<Settings>(s)) {
        double sensor[23];
        pack(xi0.sensor, sensor);
        std::vector<int> ids;
        std::vector<double> p;
        flatten(xi0.cameraLandmarks, ids, p);
        check_create(eqvio_create_from_state(settings.get(), device, capacity, stream, sensor, (int)ids.size(), ids.data(),
                                             p.data(), time, &h_));
    }
    ~VIOFilter() { eqvio_destroy(h_); }
    VIOFilter(const VIOFilter&) = delete;
    VIOFilter& operator=(const VIOFilter&) = delete;

    void processIMUData(const IMUVelocity& u) {  // VIOFilter.cpp:58-63
        check(eqvio_process_imu(h_, u.stamp, u.gyr.data(), u.acc.data(), u.gyrBiasVel.data(), u.accBiasVel.data()));
    }
    void initialiseFromIMUData(const IMUVelocity& u) {  // VIOFilter.cpp:65-78
        check(eqvio_initialise_from_imu(h_, u.stamp, u.gyr.data(), u.acc.data()));
    }
    void setState(const VIOState& xi) {  // VIOFilter.cpp:80-92
        double sensor[23];
        pack(xi.sensor, sensor);
        std::vector<int> ids;
        std::vector<double> p;
        flatten(xi.cameraLandmarks, ids, p);
        check(eqvio_set_state(h_, sensor, (int)ids.size(), ids.data(), p.data()));
    }
    void setLandmarks(const std::vector<Landmark>& lms) {  // VIOFilter.cpp:94-110
        std::vector<int> ids;
        std::vector<double> p;
        flatten(lms, ids, p);
        check(eqvio_set_landmarks(h_, (int)ids.size(), ids.data(), p.data()));
    }
    void augmentLandmarkStates(const std::vector<int>& newIds, const VIOState& provided) {  // VIOFilter.cpp:112-132
        std::vector<int> ids;
        std::vector<double> p;
        flatten(provided.cameraLandmarks, ids, p);
        check(eqvio_augment_landmark_states(h_, (int)newIds.size(), newIds.data(), (int)ids.size(), ids.data(), p.data()));
    }
    // VIOFilter.cpp:194-241.  Returns silently (like the reference) when the filter is not initialised,
    // time does not advance or the gated measurement is empty; lastUpdateRan() tells which.
    void processVisionData(const VisionMeasurement& y) {
        std::vector<int> ids;
        std::vector<double> px;
        ids.reserve(y.camCoordinates.size());
        for (const auto& kv : y.camCoordinates) {
            ids.push_back(kv.first);
            px.push_back(kv.second[0]);
            px.push_back(kv.second[1]);
        }
        int did = 0;
        check(eqvio_process_vision(h_, y.stamp, (int)ids.size(), ids.data(), px.data(), &y.camera, &did));
        ran_ = did != 0;
    }
    bool lastUpdateRan() const { return ran_; }

    double getTime() const { return eqvio_get_time(h_); }              // VIOFilter.cpp:256
    bool isInitialised() const { return eqvio_is_initialised(h_) != 0; }  // VIOFilter.h:170

    VIOState stateEstimate() const {  // VIOFilter.cpp:243
        const int N = eqvio_num_landmarks(h_);
        double sensor[23];
        std::vector<int> ids(N > 0 ? N : 1);
        std::vector<double> p(3 * (N > 0 ? N : 1));
        int n = 0;
        check(eqvio_get_state_estimate(h_, sensor, ids.data(), p.data(), &n));
        VIOState xi;
        unpack(sensor, xi.sensor);
        xi.cameraLandmarks.resize(n);
        for (int i = 0; i < n; ++i) xi.cameraLandmarks[i] = Landmark{{p[3 * i], p[3 * i + 1], p[3 * i + 2]}, ids[i]};
        return xi;
    }
    EqFStateView viewEqFState(bool withSigma = true) const {  // VIOFilter.cpp:245
        const int N = eqvio_num_landmarks(h_);
        EqFStateView v;
        v.dim = 21 + 3 * N;
        double sensor[23];
        std::vector<int> ids(N > 0 ? N : 1);
        std::vector<double> p(3 * (N > 0 ? N : 1));
        v.Q.assign(5 * (size_t)(N > 0 ? N : 1), 0.0);
        if (withSigma) v.Sigma.assign((size_t)v.dim * v.dim, 0.0);
        check(eqvio_get_eqf_state(h_, sensor, ids.data(), p.data(), v.X.data(), v.Q.data(), withSigma ? v.Sigma.data() : nullptr,
                                  v.dim));
        v.Q.resize(5 * (size_t)N);
        unpack(sensor, v.xi0.sensor);
        v.xi0.cameraLandmarks.resize(N);
        for (int i = 0; i < N; ++i) v.xi0.cameraLandmarks[i] = Landmark{{p[3 * i], p[3 * i + 1], p[3 * i + 2]}, ids[i]};
        v.currentTime = getTime();
        return v;
    }
    VisionMeasurement getFeaturePredictions(const eqvio_camera& cam, const double& stamp = -1) {  // VIOFilter.cpp:247-252
        const int N = eqvio_num_landmarks(h_);
        std::vector<int> ids(N > 0 ? N : 1);
        std::vector<double> px(2 * (N > 0 ? N : 1));
        int n = 0;
        check(eqvio_get_feature_predictions(h_, &cam, stamp, ids.data(), px.data(), &n));
        VisionMeasurement m;
        m.stamp = stamp;
        m.camera = cam;
        for (int j = 0; j < n; ++j) m.camCoordinates[ids[j]] = {px[2 * j], px[2 * j + 1]};
        return m;
    }
    // viewEqFState().computeNEES(trueState) (VIO_eqf.cpp:153-170) without moving Sigma to the host
    double computeNEES(const VIOState& trueState) const {
        double sensor[23];
        pack(trueState.sensor, sensor);
        std::vector<int> ids;
        std::vector<double> p;
        flatten(trueState.cameraLandmarks, ids, p);
        double nees = 0;
        check(eqvio_compute_nees(h_, sensor, (int)ids.size(), ids.data(), p.data(), &nees));
        return nees;
    }
    eqvio_filter* handle() const { return h_; }

private:
    static void pack(const SE3& s, double* f) {
        for (int i = 0; i < 4; ++i) f[i] = s.q[i];
        for (int i = 0; i < 3; ++i) f[4 + i] = s.x[i];
    }
    static void unpackSE3(const double* f, SE3& s) {
        for (int i = 0; i < 4; ++i) s.q[i] = f[i];
        for (int i = 0; i < 3; ++i) s.x[i] = f[4 + i];
    }
    static void pack(const VIOSensorState& s, double* f) {
        for (int i = 0; i < 6; ++i) f[i] = s.inputBias[i];
        pack(s.pose, f + 6);
        for (int i = 0; i < 3; ++i) f[13 + i] = s.velocity[i];
        pack(s.cameraOffset, f + 16);
    }
    static void unpack(const double* f, VIOSensorState& s) {
        for (int i = 0; i < 6; ++i) s.inputBias[i] = f[i];
        unpackSE3(f + 6, s.pose);
        for (int i = 0; i < 3; ++i) s.velocity[i] = f[13 + i];
        unpackSE3(f + 16, s.cameraOffset);
    }
    static void flatten(const std::vector<Landmark>& lms, std::vector<int>& ids, std::vector<double>& p) {
        ids.clear();
        p.clear();
        for (const auto& lm : lms) {
            ids.push_back(lm.id);
            p.insert(p.end(), lm.p.begin(), lm.p.end());
        }
        if (ids.empty()) {  // keep data() non-null
            ids.reserve(1);
            p.reserve(3);
        }
    }
    void check(int rc) const {
        if (rc != EQVIO_OK) throw Error(rc, eqvio_last_error(h_));
    }
    void check_create(int rc) {
        if (rc != EQVIO_OK) throw Error(rc, eqvio_last_error(nullptr));
    }
    eqvio_filter* h_ = nullptr;
    bool ran_ = false;
};

}  // namespace eqvio_b200
