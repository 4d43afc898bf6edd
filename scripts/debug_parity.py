"""Block-wise parity diagnosis of one propagate / correct step against the oracle (GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import eqvio_b200 as eb  # noqa: E402
from oracle import eqf  # noqa: E402
from parity_utils import gpu_filter, make_stream, rel_fro, snapshot_gpu, snapshot_oracle  # noqa: E402


def blocks(g, r, tag):
    S, T = g["Sigma"], r["Sigma"]
    print(f"[{tag}] Sigma all {rel_fro(S, T):.3e} | ss {rel_fro(S[:21, :21], T[:21, :21]):.3e} | sl {rel_fro(S[:21, 21:], T[:21, 21:]):.3e} "
          f"| ls {rel_fro(S[21:, :21], T[21:, :21]):.3e} | ll {rel_fro(S[21:, 21:], T[21:, 21:]):.3e} | asym {np.abs(S - S.T).max():.2e}")
    print(f"[{tag}] sensor {rel_fro(g['sensor'], r['sensor']):.3e} p {rel_fro(g['p'], r['p']):.3e} X {rel_fro(g['X_sensor'], r['X_sensor']):.3e} "
          f"Qq {rel_fro(g['Qq'], r['Qq']):.3e} Qa {rel_fro(g['Qa'], r['Qa']):.3e}")
    d = np.abs(S - T)
    i, j = np.unravel_index(np.argmax(d), d.shape)
    print(f"[{tag}] worst entry ({i},{j}) gpu {S[i, j]:.6e} ref {T[i, j]:.6e}")


for coord in (0, 1):
    over = dict(removeLostLandmarks=False)
    stream = make_stream(N=8, frames=3, coord=coord, settings_overrides=over)
    o = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
    g, cam = gpu_filter(stream)
    blocks(snapshot_gpu(g), snapshot_oracle(o), f"c{coord} init")
    fr = stream["frames"][1]
    for row in fr.imu:
        o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
    g.processIMUArray(fr.imu)
    # propagation only: empty measurement, landmarks kept
    o.processVisionData(eqf.VisionMeasurement(fr.stamp, {}, stream["cam"]))
    g.processVisionArrays(fr.stamp, np.zeros(0, dtype=np.int32), np.zeros((0, 2)), cam)
    blocks(snapshot_gpu(g), snapshot_oracle(o), f"c{coord} propagate")
    fr = stream["frames"][2]
    for row in fr.imu:
        o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
    g.processIMUArray(fr.imu)
    o.processVisionData(eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, stream["cam"]))
    g.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
    blocks(snapshot_gpu(g), snapshot_oracle(o), f"c{coord} prop+correct")
    g.close()

# detail: sensor block after one propagation
stream = make_stream(N=8, frames=3, coord=0, settings_overrides=dict(removeLostLandmarks=False))
o = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
g, cam = gpu_filter(stream)
fr = stream["frames"][1]
for row in fr.imu:
    o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
g.processIMUArray(fr.imu)
o.processVisionData(eqf.VisionMeasurement(fr.stamp, {}, stream["cam"]))
g.processVisionArrays(fr.stamp, np.zeros(0, dtype=np.int32), np.zeros((0, 2)), cam)
S, T = snapshot_gpu(g)["Sigma"], snapshot_oracle(o)["Sigma"]
d = np.abs(S - T)[:21, :21]
np.set_printoptions(linewidth=250, precision=3, suppress=False)
for (i, j) in np.argwhere(d > 1e-12):
    print("ss diff", i, j, "gpu", S[i, j], "ref", T[i, j])
d = np.abs(S - T)[21:, :21]
for (i, j) in np.argwhere(d > 1e-12)[:20]:
    print("ls diff", i, j, "gpu", S[21 + i, j], "ref", T[21 + i, j])
