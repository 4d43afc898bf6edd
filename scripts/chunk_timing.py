"""Phase timestamps (clock64) of chunk_factor_kernel, CTA 0 -- needs a library built with -DEQVIO_CHUNK_TIMING at
eqvio_b200/lib/libeqvio_b200_timing.so (the stamps come from the look-ahead kernel in the chained correction).   python scripts/chunk_timing.py [N]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shutil

os.environ.setdefault("EQVIO_B200_LIB", os.path.join(ROOT, "eqvio_b200", "lib", "libeqvio_b200_timing.so"))
import numpy as np

import eqvio_b200 as eb
from eqvio_b200 import _capi
from simdata import SimConfig, record_stream

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sm = record_stream(SimConfig.benchmark(N, 0), 12)
flt = eb.VIOFilter(eb.Settings(fastRiccati=1), eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0,
                   capacity=N + 8)
flt.setTuning(graph=0, stageS=int(os.environ.get('EQVIO_STAGE', '1')))
cam = eb.Camera(**sm.camera)
for fr in sm.frames:
    flt.processIMUArray(fr.imu)
    flt.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
    flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
fn = _capi.lib.eqvio_debug_chunk_timing
fn.restype = C.c_int
out = (C.c_longlong * 16)()
assert fn(out) == 0
t = np.array(list(out), dtype=np.int64)
factor = 0
flt_names_old = {0: "start", 1: "C/Idx loaded", 2: "S gather done", 3: "S loop done", 4: "final barrier", 5: "Yt staged", 6: "end",
                 8: "RHS gather done", 9: "RHS loop done"}
flt_names_df = {0: "start", 1: "C/Idx loaded", 2: "chain warp starts waiting", 3: "chain done", 4: "final barrier", 5: "Yt staged", 6: "end",
                8: "S warp 3: tile projected", 9: "S warp 3: past the elimination"}
names = flt_names_df if factor else flt_names_old
if factor == 2:
    names = {0: 'start', 1: 'C/Idx loaded', 2: 'tiles initialised', 3: 'elimination done', 4: '1/L_kk', 5: 'Yt staged', 6: 'end'}
base = t[0]
for i in sorted(names):
    print(f"{names[i]:>18s}: {t[i] - base:8d} cycles")

fine = getattr(_capi.lib, "eqvio_debug_chunk_fine", None)
if fine is not None:
    fine.restype = C.c_int
    buf = (C.c_longlong * 128)()
    if fine(buf) == 0:
        f = np.array(list(buf), dtype=np.int64)
        print("block column: diag-phase | barrier 1 | panel-phase + barrier 2 | trailing (thread 0's view; thread 0 owns tile (0,0))" if not factor else
              ("chain step: wait for the handed tiles | load + panel op + rank-4 update | eliminate + publish | loop back (chain warp)" if factor == 1 else
               "block column (thread 0): diagonal phase | barrier | panel phase + barrier | trailing DMMAs"))
        for J in range(16):
            a, b, c, d = f[4 * J:4 * J + 4]
            nxt = f[4 * J + 4] if J < 15 else d
            print(f"  J={J:2d}  {b - a:6d} {c - b:6d} {d - c:6d} {nxt - d:6d}   total {nxt - a:6d}")
