"""Recorded feature streams: what ``eqvio_opt`` hands to the filter AFTER the image front-end -- per image the feature ids and pixels
(``VisionMeasurement``), between images the IMU samples -- read from the files a stock EqVIO run and an ASL (EuRoC) dataset leave
on disk, replayed through the C ABI (``eqvio_replay``) and scored like ``scripts/analysis_tools.py`` scores a run.

This is the bridge for BASELINE configs[3] (EuRoC V1_01_easy): the dataset and the GIFT/OpenCV tracker are absent here, but the
filter path does not need images -- a stock ``eqvio_opt --output <dir>`` run writes ``features.csv`` (VIOWriter.cpp:83-96: stamp,
then ``id, u, v`` per tracked feature), the dataset has ``mav0/imu0/data.csv``, ``mav0/cam0/sensor.yaml`` and
``mav0/state_groundtruth_estimate0/data.csv`` (src/dataserver/ASLDatasetReader.cpp:22-131).  Those four files are a complete input
of the hot path; ``FeatureStream.fromASL`` merges them in the order of the reference's event loop (src/main_opt.cpp:178-262 with
SimpleDataServer.cpp:20-29: the image goes first when stamps tie).

Host-side plumbing only: no arithmetic of the filter lives here.
"""
import os
from dataclasses import dataclass, field

import numpy as np

from .filter import COORD_EUCLIDEAN, COORD_INVDEPTH, COORD_NORMAL, Camera, Settings, VIOFilter
from .writer import VIOWriter, trajectory_errors

# VIOFilter::Settings(const YAML::Node&), VIOFilterSettings.h:123-174: YAML key -> field
_YAML_KEYS = {
    "processVariance": dict(biasGyr="biasOmegaProcessVariance", biasAcc="biasAccelProcessVariance", attitude="attitudeProcessVariance",
                            position="positionProcessVariance", velocity="velocityProcessVariance", point="pointProcessVariance",
                            cameraAttitude="cameraAttitudeProcessVariance", cameraPosition="cameraPositionProcessVariance"),
    "measurementNoise": dict(feature="measurementNoise", featureOutlierAbs="outlierThresholdAbs",
                             featureOutlierProb="outlierThresholdProb", featureRetention="featureRetention"),
    "velocityNoise": dict(gyr="velGyrNoise", acc="velAccNoise", gyrBias="velGyrBiasWalk", accBias="velAccBiasWalk"),
    "initialVariance": dict(attitude="initialAttitudeVariance", position="initialPositionVariance", velocity="initialVelocityVariance",
                            point="initialPointVariance", pointDepth="initialPointDepthVariance", biasGyr="initialBiasOmegaVariance",
                            biasAcc="initialBiasAccelVariance", cameraAttitude="initialCameraAttitudeVariance",
                            cameraPosition="initialCameraPositionVariance"),
    "settings": dict(useDiscreteInnovationLift="useDiscreteInnovationLift", useDiscreteVelocityLift="useDiscreteVelocityLift",
                     useDiscreteStateMatrix="useDiscreteStateMatrix", fastRiccati="fastRiccati", useMedianDepth="useMedianDepth",
                     useFeaturePredictions="useFeaturePredictions", useEquivariantOutput="useEquivariantOutput",
                     removeLostLandmarks="removeLostLandmarks"),
    "initialValue": dict(sceneDepth="initialSceneDepth"),
}
_COORDS = {"Euclidean": COORD_EUCLIDEAN, "InvDepth": COORD_INVDEPTH, "Normal": COORD_NORMAL}


def _se3_from_yaml(v):
    """safeConfig's SE(3) list: a format tag of 'x' and 'w' | 'q' characters followed by seven numbers, e.g. ``[xw, x, y, z, qw, qx,
    qy, qz]`` (configs/EQVIO_config_EuRoC_stationary.yaml:72-80).  Returns (q wxyz, x) as the 7-vector ``Settings.cameraOffset`` takes."""
    tag, nums = str(v[0]), [float(t) for t in v[1:8]]
    if tag == "xw":
        x, q = nums[0:3], nums[3:7]
    elif tag == "wx":
        q, x = nums[0:4], nums[4:7]
    elif tag == "xq":  # quaternion with w last
        x, q = nums[0:3], [nums[6], nums[3], nums[4], nums[5]]
    elif tag == "qx":
        q, x = [nums[3], nums[0], nums[1], nums[2]], nums[4:7]
    else:
        raise ValueError(f"unknown SE(3) format tag {tag!r}")
    return np.array(list(q) + list(x), dtype=np.float64)


def settings_from_yaml(node) -> Settings:
    """``VIOFilter::Settings(configNode)``: ``node`` is the ``eqf`` mapping of an EqVIO configuration (a dict, or a path to the YAML
    file).  Keys that are absent keep the struct defaults (safeConfig); a missing or unknown ``coordinateChoice`` is an error, as in
    the reference (VIOFilterSettings.h:34-46)."""
    if isinstance(node, (str, os.PathLike)):
        import yaml

        with open(node) as f:
            node = yaml.safe_load(f)["eqf"]
    st = Settings()
    for group, keys in _YAML_KEYS.items():
        sub = node.get(group) or {}
        for k, name in keys.items():
            if k in sub:
                setattr(st, name, sub[k])
    choice = (node.get("settings") or {}).get("coordinateChoice")
    if choice not in _COORDS:
        raise ValueError("Invalid coordinate choice. Valid choices are Euclidean, InvDepth, Normal.")
    st.coordinateChoice = _COORDS[choice]
    off = (node.get("initialValue") or {}).get("cameraOffset")
    if off is not None:
        st.cameraOffset = _se3_from_yaml(off)
    return st


def read_features_csv(path):
    """``features.csv`` of a VIOWriter (VIOWriter.cpp:83-96) -> list of (stamp, ids int64 (n,), y float64 (n, 2))."""
    out = []
    with open(path) as f:
        next(f)  # header
        for line in f:
            t = [s for s in line.strip().split(",") if s.strip() != ""]
            if not t:
                continue
            vals = np.array([float(s) for s in t[1:]], dtype=np.float64).reshape(-1, 3)
            out.append((float(t[0]), vals[:, 0].astype(np.int64), vals[:, 1:3].copy()))
    return out


def read_asl_imu_csv(path):
    """``mav0/imu0/data.csv`` (stamp [ns], gyr xyz, acc xyz; optional bias velocities) -> (M, 13) rows ``stamp [s], gyr, acc,
    gyrBiasVel, accBiasVel`` (ASLDatasetReader.cpp:43-53, IMUVelocity.cpp:82-89)."""
    raw = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
    rows = np.zeros((raw.shape[0], 13))
    rows[:, : min(13, raw.shape[1])] = raw[:, :13]
    rows[:, 0] *= 1e-9
    return rows


def read_asl_groundtruth_csv(path):
    """``mav0/state_groundtruth_estimate0/data.csv`` -> (T, 11) rows ``time, p xyz, q wxyz, v xyz`` with the body-frame velocity the
    IMUState.csv of a run holds; stamps closer than 1e-8 s to their predecessor are dropped (ASLDatasetReader.cpp:112-131)."""
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # an empty file is a dataset without ground truth, not an error
        raw = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
    if raw.size == 0:
        return np.zeros((0, 11))
    keep, prev = [], -1e8
    for k, t in enumerate(raw[:, 0] * 1e-9):
        if t > prev + 1e-8:
            keep.append(k)
            prev = t
    raw = raw[keep]
    out = np.zeros((raw.shape[0], 11))
    out[:, 0] = raw[:, 0] * 1e-9
    out[:, 1:8] = raw[:, 1:8]
    if raw.shape[1] >= 11:  # v_RS_R is a world-frame velocity: rotate into the body frame
        for k in range(raw.shape[0]):
            w, x, y, z = raw[k, 4:8]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            out[k, 8:11] = R.T @ raw[k, 8:11]
    return out


def read_asl_camera_yaml(path):
    """``mav0/cam0/sensor.yaml`` -> (camera keyword dict for ``Camera``, cameraOffset 7-vector (q wxyz, x) from ``T_BS``)
    (ASLDatasetReader.cpp:78-105: resolution, intrinsics fu fv cu cv, radtan distortion, row-major 4 x 4 extrinsics)."""
    import yaml

    with open(path) as f:
        node = yaml.safe_load(f)
    fx, fy, cx, cy = (float(v) for v in node["intrinsics"])
    cam = dict(width=int(node["resolution"][0]), height=int(node["resolution"][1]), fx=fx, fy=fy, cx=cx, cy=cy,
               dist=[float(v) for v in node["distortion_coefficients"]])
    T = np.array([float(v) for v in node["T_BS"]["data"]], dtype=np.float64).reshape(4, 4)
    R, x = T[:3, :3], T[:3, 3]
    w = np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2.0
    if w > 1e-6:
        q = np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])
    else:  # rotation by pi: largest diagonal entry decides the branch
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(max(0.0, 1.0 + R[i, i] - R[j, j] - R[k, k])) * 2.0
        v = np.zeros(3)
        v[i] = s / 4.0
        v[j] = (R[j, i] + R[i, j]) / s
        v[k] = (R[k, i] + R[i, k]) / s
        q = np.array([(R[k, j] - R[j, k]) / s, *v])
    return cam, np.concatenate([q / np.linalg.norm(q), x])


@dataclass
class StreamFrame:
    """One image of the stream with the IMU samples the event loop delivers before it."""

    stamp: float
    imu: np.ndarray  # (k, 13)
    ids: np.ndarray  # (n,) int64
    y: np.ndarray  # (n, 2) pixels
    provided_p: object = None  # eqvio_opt's flow: landmarks enter inside processVisionData


@dataclass
class FeatureStream:
    imu: np.ndarray  # (M, 13): stamp [s], gyr, acc, gyrBiasVel, accBiasVel
    features: list  # [(stamp, ids, y)]
    camera: dict  # keyword arguments of eqvio_b200.Camera
    cameraOffset: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0, 0, 0, 0]))  # (q wxyz, x)
    groundtruth: np.ndarray = None  # (T, 11) or None

    @staticmethod
    def fromASL(datasetDir, featuresCsv, cameraLag=0.0):
        """An ASL / EuRoC dataset directory (the one holding ``mav0/``) plus the ``features.csv`` of a stock eqvio_opt run on it.
        ``cameraLag`` is ``main:cameraLag`` of the configuration; it was already subtracted from the stamps in features.csv
        (ASLDatasetReader.cpp:71), so it is NOT applied again -- the argument only documents the run."""
        del cameraLag
        mav = os.path.join(datasetDir, "mav0")
        cam, off = read_asl_camera_yaml(os.path.join(mav, "cam0", "sensor.yaml"))
        gt_path = os.path.join(mav, "state_groundtruth_estimate0", "data.csv")
        return FeatureStream(imu=read_asl_imu_csv(os.path.join(mav, "imu0", "data.csv")), features=read_features_csv(featuresCsv), camera=cam,
                             cameraOffset=off, groundtruth=(read_asl_groundtruth_csv(gt_path) if os.path.exists(gt_path) else None))

    def frames(self, startTime=0.0):
        """The reference's event loop (main_opt.cpp:178-262): measurements in stamp order, the image first when an image and an IMU
        sample carry the same stamp (SimpleDataServer.cpp:20-29), everything before ``startTime`` skipped (when it is positive);
        IMU samples after the last image are dropped (they change no output row)."""
        out, k, M = [], 0, self.imu.shape[0]
        for stamp, ids, y in self.features:
            k0 = k
            while k < M and self.imu[k, 0] < stamp:
                k += 1
            if startTime > 0 and stamp < startTime:
                continue
            rows = self.imu[k0:k]
            if startTime > 0:
                rows = rows[rows[:, 0] >= startTime]
            out.append(StreamFrame(float(stamp), np.ascontiguousarray(rows), np.asarray(ids, dtype=np.int64), np.asarray(y, dtype=np.float64)))
        return out

    def save(self, path):
        """One ``.npz``: flat arrays + offsets (a few MB for a EuRoC sequence at 40 features)."""
        n = np.array([len(f[1]) for f in self.features], dtype=np.int64)
        cam = self.camera
        np.savez_compressed(
            path, imu=self.imu, stamps=np.array([f[0] for f in self.features]), counts=n,
            ids=np.concatenate([f[1] for f in self.features]) if len(n) else np.zeros(0, dtype=np.int64),
            y=np.concatenate([f[2] for f in self.features]) if len(n) else np.zeros((0, 2)),
            cam_size=np.array([cam["width"], cam["height"]], dtype=np.int64), cam_k=np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]]),
            cam_dist=np.array(cam.get("dist", ()), dtype=np.float64), cameraOffset=np.asarray(self.cameraOffset, dtype=np.float64),
            groundtruth=self.groundtruth if self.groundtruth is not None else np.zeros((0, 11)))

    @staticmethod
    def load(path):
        z = np.load(path)
        off = np.concatenate([[0], np.cumsum(z["counts"])])
        feats = [(float(z["stamps"][k]), z["ids"][off[k]:off[k + 1]].astype(np.int64), z["y"][off[k]:off[k + 1]].copy())
                 for k in range(len(z["counts"]))]
        cam = dict(width=int(z["cam_size"][0]), height=int(z["cam_size"][1]), fx=float(z["cam_k"][0]), fy=float(z["cam_k"][1]),
                   cx=float(z["cam_k"][2]), cy=float(z["cam_k"][3]), dist=[float(v) for v in z["cam_dist"]])
        gt = z["groundtruth"]
        return FeatureStream(imu=z["imu"], features=feats, camera=cam, cameraOffset=z["cameraOffset"], groundtruth=gt if gt.shape[0] else None)


def interpolate_groundtruth(gt, times):
    """Ground-truth rows at ``times``: linear in position and velocity, normalised linear in the quaternion (sign-aligned) -- the
    comparison-time resampling of scripts/analysis_tools.py:113-128."""
    gt = np.asarray(gt, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    idx = np.clip(np.searchsorted(gt[:, 0], times), 1, gt.shape[0] - 1)
    a, b = gt[idx - 1], gt[idx]
    w = np.clip((times - a[:, 0]) / np.maximum(b[:, 0] - a[:, 0], 1e-300), 0.0, 1.0)[:, None]
    out = a + w * (b - a)
    qa, qb = a[:, 4:8], b[:, 4:8].copy()
    qb[np.sum(qa * qb, axis=1) < 0] *= -1.0
    q = qa + w * (qb - qa)
    out[:, 4:8] = q / np.linalg.norm(q, axis=1, keepdims=True)
    out[:, 0] = times
    return out


def run_stream(stream: FeatureStream, settings: Settings, outputDir=None, capacity=None, startTime=0.0, useReplay=True, device=0):
    """``eqvio_opt`` on a recorded stream: a filter built from the settings alone (initialised by its first IMU sample,
    VIOFilter.cpp:58-78) with the dataset's camera extrinsics as ``cameraOffset`` (main_opt.cpp), every frame through
    ``processIMUData`` x k / ``processVisionData`` / ``stateEstimate``, the VIOWriter files when ``outputDir`` is given.
    Returns dict(IMUState (T, 11) rows as in IMUState.csv, frame_ms, errors = trajectory_errors against the ground truth or None)."""
    settings.cameraOffset = stream.cameraOffset
    frames = stream.frames(startTime)
    if capacity is None:
        capacity = max([len(f.ids) for f in frames] + [8]) * 2 + 8
    cam = Camera(**stream.camera)
    flt = VIOFilter(settings, capacity=capacity, device=device)
    rows, frame_ms = [], None
    writer = VIOWriter(outputDir) if outputDir else None
    try:
        if useReplay and writer is None:
            frame_ms, est = flt.replay(frames, cam)
            est = np.asarray(est).reshape(len(frames), -1)
            # est_sensor: bias6 | pose q(wxyz) x | velocity | camera offset -> IMUState.csv columns
            seen = False  # getTime() of a filter that has not seen an IMU sample yet is -1 (VIOFilter.cpp:58-78, :194-200)
            for fr, e in zip(frames, est):
                seen = seen or len(fr.imu) > 0
                rows.append([fr.stamp if seen else -1.0, *e[10:13], *e[6:10], *e[13:16]])
        else:
            for fr in frames:
                if len(fr.imu):
                    flt.processIMUArray(fr.imu)
                flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
                xi = flt.stateEstimate()
                s = xi.sensor
                rows.append([flt.getTime(), *s.pose_x, *s.pose_q, *s.velocity])
                if writer:
                    writer.writeStates(flt.getTime(), xi)
                    writer.writeFeatures(fr.stamp, fr.ids, fr.y)
    finally:
        if writer:
            writer.close()
        flt.close()
    rows = np.array(rows, dtype=np.float64).reshape(-1, 11)
    errors = None
    if stream.groundtruth is not None and len(stream.groundtruth) > 1 and len(rows) > 1:
        gt = stream.groundtruth
        sel = (rows[:, 0] >= gt[0, 0]) & (rows[:, 0] <= gt[-1, 0])
        if sel.sum() > 2:
            errors = trajectory_errors(rows[sel], interpolate_groundtruth(gt, rows[sel, 0]))
    return dict(IMUState=rows, frame_ms=frame_ms, errors=errors)
