"""The standalone input generator (simdata/, used by bench.py) produces the same streams as the
oracle's VIOSimulator restatement: two independently written restatements of
src/VIOSimulator.cpp agree to rounding."""
import numpy as np
import pytest

from oracle.simulator import SimulationDataServer, benchmarkSettings, benchmarkSim
from simdata import SimConfig, record_stream


@pytest.mark.parametrize("noisy", [False, True])
def test_simdata_matches_oracle_simulator(noisy):
    N, frames = 24, 5
    st = benchmarkSettings(0, measurementNoise=0.5)
    sim = benchmarkSim(N, 3, outputNoise=noisy, inputNoise=noisy)
    server = SimulationDataServer(sim, st)
    init = server.initialCondition()
    ref = server.record(frames)
    got = record_stream(SimConfig.benchmark(N, 3, outputNoise=noisy, inputNoise=noisy, measurementNoise=0.5), frames)
    assert np.array_equal(got.init_ids, init.ids)
    assert np.allclose(got.init_p, init.p, rtol=0, atol=1e-12)
    assert np.allclose(got.init_sensor, init.sensor.flat(), rtol=0, atol=1e-12)
    assert len(got.frames) == len(ref) == frames
    for a, b in zip(got.frames, ref):
        assert a.stamp == b.stamp
        assert np.array_equal(a.ids, b.ids)
        assert np.allclose(a.y, b.y, rtol=0, atol=1e-9)
        assert np.allclose(a.provided_p, b.provided_p, rtol=0, atol=1e-12)
        assert a.imu.shape == b.imu.shape and np.allclose(a.imu, b.imu, rtol=0, atol=1e-10)
        assert np.allclose(a.true_sensor, b.true_sensor, rtol=0, atol=1e-12)


def test_visible_count_holds_over_a_lap():
    """with 4 walls at 2 m and 20 N points, >= N landmarks stay in view."""
    s = record_stream(SimConfig.benchmark(32, 0), 400)
    assert len(s.frames) == 400
    assert min(len(f.ids) for f in s.frames) == 32
    assert all(f.imu.shape[0] == 10 for f in s.frames[1:]) and s.frames[0].imu.shape[0] == 0
