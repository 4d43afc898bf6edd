"""Device VIOSimulator vs the host generator (simdata): identical ids (visibility, selection in shuffled order, sorting),
pixels / positions / IMU to rounding (the trajectory is re-evaluated with the device's sin / cos, and the accelerations come
from a cubic fit through pose samples 0.5 ms apart, which amplifies last-bit differences by ~1 / dt^2)."""
import numpy as np
import pytest

from simdata import SimConfig, record_stream

pytestmark = pytest.mark.gpu


def test_device_simulator_matches_host_generator():
    from eqvio_b200.simulator import DeviceSimulator

    N, frames = 48, 30
    cfgs = [SimConfig.benchmark(N, seed) for seed in (0, 3, 11)]
    sim = DeviceSimulator(cfgs)
    got = sim.record_streams(frames)
    for cfg, g in zip(cfgs, got):
        ref = record_stream(cfg, frames)
        assert np.array_equal(g.init_ids, ref.init_ids)
        np.testing.assert_allclose(g.init_p, ref.init_p, rtol=0, atol=1e-9)
        np.testing.assert_allclose(g.init_sensor, ref.init_sensor, rtol=0, atol=1e-7)
        assert len(g.frames) == len(ref.frames) == frames
        for a, b in zip(g.frames, ref.frames):
            assert a.stamp == b.stamp
            assert np.array_equal(a.ids, b.ids)  # discrete part: bit-identical indexing
            np.testing.assert_allclose(a.y, b.y, rtol=0, atol=1e-7)
            np.testing.assert_allclose(a.provided_p, b.provided_p, rtol=0, atol=1e-9)
            assert a.imu.shape == b.imu.shape
            np.testing.assert_allclose(a.imu[:, :4], b.imu[:, :4], rtol=0, atol=1e-9)   # stamp, gyro
            np.testing.assert_allclose(a.imu[:, 4:7], b.imu[:, 4:7], rtol=0, atol=1e-5)  # accelerometer (cubic fit)
            np.testing.assert_allclose(a.true_sensor, b.true_sensor, rtol=0, atol=1e-7)
    sim.close()


def test_filter_runs_on_a_device_generated_stream():
    """A stream from the device simulator drives the filter to the same estimates as the host-generated one."""
    import eqvio_b200 as eb
    from eqvio_b200.simulator import DeviceSimulator

    cfg = SimConfig.benchmark(32, 5)
    sim = DeviceSimulator([cfg])
    outs = []
    for sm in (sim.record_streams(12)[0], record_stream(cfg, 12)):
        flt = eb.VIOFilter(eb.Settings(fastRiccati=1), eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0,
                           capacity=40)
        cam = eb.Camera(**sm.camera)
        for fr in sm.frames:
            flt.processIMUArray(fr.imu)
            flt.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
            flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        outs.append(flt.stateEstimate())
        flt.close()
    assert np.array_equal(outs[0].ids, outs[1].ids)
    np.testing.assert_allclose(outs[0].sensor.flat(), outs[1].sensor.flat(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(outs[0].p, outs[1].p, rtol=0, atol=1e-6)
    sim.close()
