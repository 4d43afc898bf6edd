"""Device VIOSimulator throughput vs the host generator: frames (vision measurements) per second for I Monte-Carlo instances
(BASELINE config 5: 16 instances per GPU).   python scripts/sim_throughput.py [N] [instances] [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from eqvio_b200.simulator import DeviceSimulator
from simdata import SimConfig
from simdata.vio_simulator import _Simulator

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
I = int(sys.argv[2]) if len(sys.argv) > 2 else 16
K = int(sys.argv[3]) if len(sys.argv) > 3 else 400
cfgs = [SimConfig.benchmark(N, s) for s in range(I)]
t0 = time.perf_counter()
sim = DeviceSimulator(cfgs)
t_setup = time.perf_counter() - t0
stamps = np.arange(K) / 20.0
sim.vision(stamps[:8])
t0 = time.perf_counter()
n, ids, y, p, sensor = sim.vision(stamps)
t_dev = time.perf_counter() - t0
host = _Simulator(cfgs[0])
t0 = time.perf_counter()
for t in stamps[:40]:
    host.vision(t)
    host.full_state(t)
t_host = (time.perf_counter() - t0) / 40
print(f"N={N} ({cfgs[0].numPoints} world points), {I} instances x {K} frames: device kernel {sim.last_vision_ms:.2f} ms, call incl. copies "
      f"{1e3 * t_dev:.1f} ms = {I * K / t_dev:.0f} frames/s ({I * K / (sim.last_vision_ms * 1e-3):.0f} frames/s in the kernel); host generator "
      f"{1e3 * t_host:.2f} ms/frame = {1.0 / t_host:.0f} frames/s on one core; world-point setup {t_setup:.2f} s; min visible {int(n.min())}")
