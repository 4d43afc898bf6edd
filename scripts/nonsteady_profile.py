"""Steady vs non-steady frames: the same simulated stream driven (a) as eqvio_sim does -- augmentLandmarkStates before every
processVisionData (main_sim.cpp:139-142), so that the vision call sees a steady frame -- and (b) as eqvio_opt does on real
data -- no augment call, lost ids are pruned and new ids added INSIDE processVisionData (VIOFilter.cpp:203-219).
    python scripts/nonsteady_profile.py [N] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import eqvio_b200 as eb
from simdata import SimConfig, record_stream

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 60
sm = record_stream(SimConfig.benchmark(N, 0), 10 + K)
cam = eb.Camera(**sm.camera)


class Fr:
    def __init__(self, f, aug):
        self.stamp, self.imu, self.ids, self.y = f.stamp, f.imu, f.ids, f.y
        self.provided_p = f.provided_p if aug else None


for aug in (True, False):
    flt = eb.VIOFilter(eb.Settings(fastRiccati=1), eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0,
                       capacity=N + 64)
    frames = [Fr(f, aug) for f in sm.frames]
    flt.replay(frames[:10], cam)
    flt.hostProfile(reset=True)
    ms, _ = flt.replay(frames[10:], cam, flushBytes=256 << 20)
    hp = flt.hostProfile()
    changed = np.mean([len(set(a.ids) ^ set(b.ids)) for a, b in zip(sm.frames[9:-1], sm.frames[10:])])
    print(f"N={N} augment={aug}: {1000.0 / ms.mean():.0f} updates/s ({ms.mean() * 1e3:.0f} us/frame), "
          f"ids changing per frame {changed:.1f}, landmarks at the end {flt.numLandmarks()}, host profile {({k: round(v, 1) for k, v in hp.items()})}")
    flt.close()
