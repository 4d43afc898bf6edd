"""Loader of the committed golden fixtures (tests/golden/*.npz, written by make_golden.py)."""
import os

import numpy as np

from oracle import eqf
from oracle.camera import PinholeCamera
from oracle.simulator import Frame

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["euclid_n16", "invdepth_n16", "euclid_n24_gated_noisy"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    st = eqf.Settings()
    for k, v in zip(z["settings_names"], z["settings_values"]):
        k = str(k)
        cur = getattr(st, k)
        setattr(st, k, bool(v) if isinstance(cur, bool) else (int(v) if isinstance(cur, int) else float(v)))
    c = z["cam_pod"]
    cam = PinholeCamera(int(c[0]), int(c[1]), c[2], c[3], c[4], c[5])
    init = eqf.VIOState(eqf.VIOSensorState.fromFlat(z["init_sensor"]), z["init_p"], z["init_ids"])
    st.cameraOffset = init.sensor.cameraOffset.copy()
    frames, outs = [], []
    for k in range(int(z["num_frames"])):
        frames.append(Frame(float(z[f"f{k}_stamp"]), z[f"f{k}_ids"], z[f"f{k}_y"], z[f"f{k}_provided_p"], z[f"f{k}_imu"]))
        outs.append(dict(ids=z[f"o{k}_ids"], sensor=z[f"o{k}_sensor"], p=z[f"o{k}_p"], Sigma=z[f"o{k}_Sigma"],
                         time=float(z[f"o{k}_time"])))
    stream = dict(settings=st, cam=cam, init=init, frames=frames, N=max(len(f.ids) for f in frames))
    return stream, outs
