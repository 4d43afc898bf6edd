"""Generate the committed golden fixtures: recorded VIOSimulator input streams and the oracle's
per-update outputs for them.  The reference cannot be built in this image, so these
vectors originate from the oracle restatement, not from the Eigen binary; they pin the oracle against
regressions and give the GPU tests inputs that do not depend on the simulator code.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from parity_utils import make_stream, run_oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "euclid_n16": dict(N=16, frames=6, coord=0),
    "invdepth_n16": dict(N=16, frames=6, coord=1),
    "euclid_n24_gated_noisy": dict(N=24, frames=6, coord=0,
                                   settings_overrides=dict(outlierThresholdAbs=3.0, outlierThresholdProb=6.0,
                                                           measurementNoise=0.5, featureRetention=0.2),
                                   sim_overrides=dict(outputNoise=True, inputNoise=True)),
}


def pack(stream, outs):
    d = {}
    st = stream["settings"]
    d["settings_names"] = np.array([k for k in vars(st) if k != "cameraOffset"])
    d["settings_values"] = np.array([float(getattr(st, k)) for k in d["settings_names"]])
    d["cam_pod"] = np.array([stream["cam"].width, stream["cam"].height, stream["cam"].fx, stream["cam"].fy,
                             stream["cam"].cx, stream["cam"].cy])
    d["init_sensor"] = stream["init"].sensor.flat()
    d["init_p"] = stream["init"].p
    d["init_ids"] = stream["init"].ids
    d["num_frames"] = np.array(len(stream["frames"]))
    for k, fr in enumerate(stream["frames"]):
        d[f"f{k}_stamp"] = np.array(fr.stamp)
        d[f"f{k}_ids"] = fr.ids
        d[f"f{k}_y"] = fr.y
        d[f"f{k}_provided_p"] = fr.provided_p
        d[f"f{k}_imu"] = fr.imu
        o = outs[k]
        d[f"o{k}_ids"] = o["ids"]
        d[f"o{k}_sensor"] = o["sensor"]
        d[f"o{k}_p"] = o["p"]
        d[f"o{k}_Sigma"] = o["Sigma"]
        d[f"o{k}_time"] = np.array(o["time"])
    return d


if __name__ == "__main__":
    for name, kw in CASES.items():
        stream = make_stream(**kw)
        outs = run_oracle(stream)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack(stream, outs))
        print(name, "frames", len(outs), "N", [len(o["ids"]) for o in outs])
