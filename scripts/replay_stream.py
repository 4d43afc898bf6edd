"""eqvio_opt on a recorded feature stream (BASELINE configs[3] bridge, eqvio_b200/stream.py):

    python scripts/replay_stream.py <ASL dataset dir (holds mav0/)> <features.csv of a stock eqvio_opt run> <EqVIO config .yaml>
                                    [--output DIR] [--start T] [--save-npz FILE]

Reads mav0/imu0/data.csv, mav0/cam0/sensor.yaml, mav0/state_groundtruth_estimate0/data.csv and the features.csv, runs the filter on
cuda:0 through the C ABI with the configuration's `eqf` settings and prints the frame rate and the trajectory-error summary of
scripts/analysis_tools.py (Sim(3)-aligned position RMSE, attitude, velocity, scale) as one JSON line.  With --output the VIOWriter
files (IMUState.csv, camera.csv, bias.csv, points.csv, features.csv) are written there, ready for the reference's analysis scripts."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("dataset")
    ap.add_argument("features")
    ap.add_argument("config")
    ap.add_argument("--output", default=None)
    ap.add_argument("--start", type=float, default=0.0, help="main:startTime of the configuration")
    ap.add_argument("--save-npz", default=None, help="also store the merged stream as one .npz (FeatureStream.save)")
    a = ap.parse_args()
    import numpy as np

    from eqvio_b200.stream import FeatureStream, run_stream, settings_from_yaml

    stream = FeatureStream.load(a.dataset) if a.dataset.endswith(".npz") else FeatureStream.fromASL(a.dataset, a.features)
    if a.save_npz:
        stream.save(a.save_npz)
    out = run_stream(stream, settings_from_yaml(a.config), outputDir=a.output, startTime=a.start)
    line = dict(frames=int(out["IMUState"].shape[0]), errors=out["errors"])
    if out["frame_ms"] is not None and len(out["frame_ms"]):
        ms = np.asarray(out["frame_ms"], dtype=np.float64)
        line.update(median_frame_ms=float(np.median(ms)), frames_per_s=float(1e3 * len(ms) / ms.sum()))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
