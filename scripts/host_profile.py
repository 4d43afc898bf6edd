"""Where does the host time of one e2e step go?  (GPU box)  python scripts/host_profile.py [N]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
import numpy as np  # noqa: E402

import eqvio_b200 as eb  # noqa: E402
from simdata import SimConfig, record_stream  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sm = record_stream(SimConfig.benchmark(N, 0), 120)
st = eb.Settings(fastRiccati=1)
flt = eb.VIOFilter(st, eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0, capacity=N + 8)
cam = eb.Camera(**sm.camera)
empty = eb.VIOSensorState()


def step(fr):
    flt.processIMUArray(fr.imu)
    flt.augmentLandmarkStates(fr.ids, eb.VIOState(empty, fr.provided_p, fr.ids))
    flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
    return flt.stateEstimate()


for fr in sm.frames[:20]:
    step(fr)
flt.hostProfile(reset=True)
t = {"imu": 0.0, "aug": 0.0, "vis": 0.0, "est": 0.0}
for fr in sm.frames[20:70]:
    t0 = time.perf_counter(); flt.processIMUArray(fr.imu)
    t1 = time.perf_counter(); flt.augmentLandmarkStates(fr.ids, eb.VIOState(empty, fr.provided_p, fr.ids))
    t2 = time.perf_counter(); flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
    t3 = time.perf_counter(); flt.stateEstimate()
    t4 = time.perf_counter()
    t["imu"] += t1 - t0; t["aug"] += t2 - t1; t["vis"] += t3 - t2; t["est"] += t4 - t3
print({k: round(1e6 * v / 50, 1) for k, v in t.items()}, "us per step; total", round(1e6 * sum(t.values()) / 50, 1))
print("inside eqvio_process_vision:", {k: round(v, 1) for k, v in flt.hostProfile().items()})
pr = cProfile.Profile()
pr.enable()
for fr in sm.frames[70:120]:
    step(fr)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
