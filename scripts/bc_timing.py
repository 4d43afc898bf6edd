"""Phase timestamps (clock64, thread 0) of bc_diag_kernel at block column 1 -- needs a library built with -DEQVIO_CHUNK_TIMING.
    EQVIO_B200_LIB=eqvio_b200/lib/<timing build>.so python scripts/bc_timing.py [N]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("EQVIO_B200_LIB", os.path.join(ROOT, "eqvio_b200", "lib", "libeqvio_b200_timing.so"))
import numpy as np

import eqvio_b200 as eb
from eqvio_b200 import _capi
from simdata import SimConfig, record_stream

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sm = record_stream(SimConfig.benchmark(N, 0), 12)
flt = eb.VIOFilter(eb.Settings(fastRiccati=1), eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0,
                   capacity=N + 8)
flt.setTuning(graph=1, correction=2)
cam = eb.Camera(**sm.camera)
try:
    for fr in sm.frames:
        flt.processIMUArray(fr.imu)
        flt.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
except Exception as e:  # timing-only builds (e.g. -DBC_NO_INV) produce garbage: the stamps of the first update are still valid
    print("update failed (expected for timing-only variants):", e)
fn = _capi.lib.eqvio_debug_bc_timing
fn.restype = C.c_int
out = (C.c_longlong * 16)()
warps = (C.c_int * (9 * 64))()
assert fn(out, warps) == 0
t = np.array(list(out), dtype=np.int64)
names = {0: "start", 1: "past the dependency wait", 2: "T, M^T in shared memory", 3: "P = T M^T done", 4: "D = T - P P^T in tiles",
         5: "factor / inverse loop done", 6: "M^T stored"}
for i in sorted(names):
    print(f"{names[i]:>28s}: {t[i] - t[0]:8d} cycles")
w = np.array(list(warps), dtype=np.int64).reshape(9, 16, 4)
t0 = w[:, 0, 0].min()
print("per S-group warp (lane 0), cycles since the loop start: loop top | after barrier 1 | after barrier 2")
for J in range(16):
    print(f"  J={J:2d}  " + "   ".join(" ".join(f"{(w[q, J, i] - t0) & 0xffffffff:6d}" for i in (0, 2, 3)) for q in range(9)))

gt = getattr(_capi.lib, "eqvio_debug_bc_gt", None)
if gt is not None:
    buf = (C.c_ulonglong * 32)()
    if gt(buf) == 0:
        v = np.array(list(buf), dtype=np.float64)
        t0 = v[0]
        lab = {0: "diag(0) start", 1: "diag(0) loop start", 2: "diag(0) loop end", 3: "diag(0) XT flag 0", 4: "diag(0) XT flag 1", 5: "diag(0) XT flag 2",
               6: "diag(0) XT flag 3", 8: "diag(1) start (T fetched)", 9: "diag(1) loop start", 10: "diag(1) loop end", 16: "diag(1) saw XT flag 0",
               17: "diag(1) saw XT flag 1", 18: "diag(1) saw XT flag 2", 19: "diag(1) saw XT flag 3"}
        print("hand-over between the first two diagonal steps (globaltimer, us since diag(0) start):")
        for i in sorted(lab, key=lambda i: v[i]):
            print(f"  {lab[i]:>28s}: {(v[i] - t0) / 1e3:7.2f}")
