"""Device timeline of the chunk kernels of ONE steady update (globaltimer stamps, first start / last end per launch).
Needs a library built with -DEQVIO_TIMELINE:
    nvcc <flags of __graft_entry__> -DEQVIO_TIMELINE -o eqvio_b200/lib/libeqvio_b200_tl.so eqvio_b200/csrc/filter.cu -ldl
    EQVIO_B200_LIB=eqvio_b200/lib/libeqvio_b200_tl.so python scripts/timeline.py [N] [lookahead 0|1] [graph 0|1] [correction 0|1|2]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import eqvio_b200 as eb
from eqvio_b200 import _capi
from simdata import SimConfig, record_stream

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
look = int(sys.argv[2]) if len(sys.argv) > 2 else 1
graph = int(sys.argv[3]) if len(sys.argv) > 3 else 0
corr = int(sys.argv[4]) if len(sys.argv) > 4 else None
sm = record_stream(SimConfig.benchmark(N, 0), 14)
flt = eb.VIOFilter(eb.Settings(fastRiccati=1), eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0,
                   capacity=N + 8)
flt.setTuning(graph=graph, lookahead=look, correction=corr)
cam = eb.Camera(**sm.camera)
fn = _capi.lib.eqvio_debug_timeline
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
buf = (C.c_ulonglong * 1024)()
try:
    for k, fr in enumerate(sm.frames):
        flt.processIMUArray(fr.imu)
        flt.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        if k == len(sm.frames) - 1 or os.environ.get("EQVIO_TL_FIRST"):
            fn(flt._h, buf, 512, 1)
            if os.environ.get("EQVIO_TL_FLUSH"):  # cold L2, as in bench.py: 256 MiB write (EQVIO_TL_FLUSH=2: followed by a read sweep)
                import torch
                scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
                scratch.zero_()
                if os.environ["EQVIO_TL_FLUSH"] == "2":
                    scratch2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
                    scratch2.zero_()
                    torch.cuda.synchronize()
                    scratch.sum()
                torch.cuda.synchronize()
        flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
except Exception as e:  # timing-only kernel variants produce garbage: the stamps of the failing update are still there
    print("update failed:", e)
n = fn(flt._h, buf, 512, 0)
nm = _capi.lib.eqvio_debug_timeline_name
nm.restype = C.c_char_p
nm.argtypes = [C.c_void_p, C.c_int]
t = np.array(buf[:2 * n], dtype=np.float64).reshape(n, 2)
valid = np.nonzero(t[:, 1] > 0)[0]  # slots stamped since the reset (a replayed graph keeps the slot numbers of its capture)
names = [nm(flt._h, int(i)).decode() for i in valid]
t = t[valid]
order = np.argsort(t[:, 0], kind="stable")
t, names = t[order], [names[i] for i in order]
n = len(t)
t0 = t[:, 0].min()
print(f"N={N} lookahead={look} graph={graph}: {n} launches, span {(t[:, 1].max() - t0) / 1e3:.1f} us")
for i in range(n):
    print(f"{i:3d}  start {(t[i, 0] - t0) / 1e3:8.1f}  end {(t[i, 1] - t0) / 1e3:8.1f}  dur {(t[i, 1] - t[i, 0]) / 1e3:6.1f}  {names[i]}")
