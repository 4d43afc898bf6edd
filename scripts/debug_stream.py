"""Debug: recorded-stream test case, GPU vs oracle frame by frame (ids, outliers, state error)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_stream as T
import eqvio_b200 as eb
from eqvio_b200 import VIOWriter
from eqvio_b200.stream import FeatureStream, settings_from_yaml
from oracle import eqf
from oracle.camera import StandardCamera
from oracle.simulator import SimulationDataServer, benchmarkSim
from parity_utils import make_stream, snapshot_gpu, snapshot_oracle, compare_states

tmp = tempfile.mkdtemp()
N, frames = 40, 8
ocam5 = StandardCamera(752, 480, *T.EUROC_CAM0["intrinsics"], T.EUROC_CAM0["distortion"] + [0.0])
base = make_stream(N=N, frames=1, coord=1)
server = SimulationDataServer(benchmarkSim(N, 0, outputNoise=True, inputNoise=True), base["settings"])
server.simulator.cameraPtr = ocam5
rec = server.record(frames)
csv_round = int(os.environ.get("CSV", "1"))
ost = base["settings"]
gst = settings_from_yaml(T.EUROC_EQF)
for name in gst._names:
    if name != "cameraOffset":
        setattr(ost, name, getattr(gst, name))
ext = server.cameraExtrinsics()
ost.cameraOffset.q = np.asarray(ext.q).copy(); ost.cameraOffset.x = np.asarray(ext.x).copy()
gst.cameraOffset = np.concatenate([ext.q, ext.x])
ocam = StandardCamera(752, 480, *T.EUROC_CAM0["intrinsics"], T.EUROC_CAM0["distortion"])
cam = eb.Camera(752, 480, *T.EUROC_CAM0["intrinsics"], T.EUROC_CAM0["distortion"])
o = eqf.VIOFilter(ost)
g = eb.VIOFilter(gst, capacity=96)
if os.environ.get("NOGRAPH"): g.setTuning(graph=0, speculate=0)
for k, fr in enumerate(rec):
    y = np.array([[float("%g" % v) for v in row] for row in fr.y]) if csv_round else fr.y
    order = np.argsort(fr.ids); ids = np.asarray(fr.ids)[order]; y = y[order]
    for r in fr.imu:
        o.processIMUData(eqf.IMUVelocity(r[0], r[1:4], r[4:7], r[7:10], r[10:13]))
    if len(fr.imu): g.processIMUArray(fr.imu)
    o.processVisionData(eqf.VisionMeasurement.fromArrays(fr.stamp, ids, y, ocam))
    g.processVisionArrays(fr.stamp, ids, y, cam)
    so, sg = snapshot_oracle(o), snapshot_gpu(g)
    e = compare_states(sg, so)
    print(k, "n_o", len(so["ids"]), "n_g", len(sg["ids"]), "ids_equal", e["ids_equal"], "sigma", e.get("sigma"), "state", e.get("state"),
          "gpu outliers", sorted(g.lastOutliers()), "only_o", sorted(set(so["ids"]) - set(sg["ids"])), "only_g", sorted(set(sg["ids"]) - set(so["ids"])))
