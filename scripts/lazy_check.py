"""Lazy trailing updates of the look-ahead correction (EQVIO_TUNE_LAZY_DOWNDATE) against the every-chunk form: the same stream through
filters with M = 0, 1, 2, 3, ... must leave bit-identical Sigma / state (per tile the same products in the same order).
    python scripts/lazy_check.py [N] [frames] [graph 0|1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import eqvio_b200 as eb
from simdata import SimConfig, record_stream

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
graph = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sm = record_stream(SimConfig.benchmark(N, 0), frames)
cam = eb.Camera(**sm.camera)


def run(M, split=1):
    os.environ["EQVIO_B200_BAND_SPLIT"] = str(split)
    flt = eb.VIOFilter(eb.Settings(fastRiccati=1), eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0,
                       capacity=N + 8)
    flt.setTuning(graph=graph, lookahead=1, correction=0, lazyDowndate=M)
    for fr in sm.frames:
        flt.processIMUArray(fr.imu)
        flt.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
    st = flt.viewEqFState(withSigma=True)
    flt.close()
    return np.array(st.Sigma), np.concatenate([np.ravel(st.X_sensor), np.ravel(st.X_Qq), np.ravel(st.X_Qa)])


ref = run(0)
print("N", N, "frames", frames, "graph", graph, "Sigma", ref[0].shape, "finite", bool(np.isfinite(ref[0]).all()),
      "sym", float(np.abs(ref[0] - ref[0].T).max()))
bad = 0
for M, split in ((1, 0), (1, 1), (2, 1), (2, 0), (3, 1), (4, 1), (8, 1)):
    got = run(M, split)
    dS = float(np.abs(got[0] - ref[0]).max())
    dx = float(np.abs(got[1] - ref[1]).max())
    same = bool(np.array_equal(got[0], ref[0]))
    print(f"M={M} split={split}: max|dSigma|={dS:.3e} max|dstate|={dx:.3e} bit-identical={same}")
    bad += not same
sys.exit(1 if bad else 0)
