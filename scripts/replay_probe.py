"""Per-frame host wall clock of the C++ replay loop (eqvio_replay) for several simulator instances: which frames leave the steady path.
    python scripts/replay_probe.py [N] [instances] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import eqvio_b200 as eb
from simdata import SimConfig, record_stream

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
I = int(sys.argv[2]) if len(sys.argv) > 2 else 4
F = int(sys.argv[3]) if len(sys.argv) > 3 else 50
for inst in range(I):
    sm = record_stream(SimConfig.benchmark(N, inst, duration=20.0), 1 + F)
    flt = eb.VIOFilter(eb.Settings(fastRiccati=1), eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0,
                       capacity=N + 8)
    cam = eb.Camera(**sm.camera)
    g0 = flt.graphStats()
    ms, est = flt.replay(sm.frames[1:1 + F], cam, flushBytes=256 << 20)
    g1 = flt.graphStats()
    print(f"instance {inst}: median {np.median(ms):.3f} ms, mean {ms.mean():.3f}, max {ms.max():.3f}; graphs captured {g1[0] - g0[0]} replayed {g1[1] - g0[1]}; "
          f"landmarks {flt.numLandmarks()}; frames > 0.3 ms: {[int(k) for k in np.nonzero(ms > 0.3)[0]]}")
    print("   ", " ".join(f"{v:.2f}" for v in ms))
    flt.close()
