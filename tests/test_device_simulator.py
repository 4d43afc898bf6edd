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


def test_device_noise_is_the_philox_function():
    """Input / output noise on the device (VIOSimulator.cpp:163-167, 258-262): noisy stream - noise-free stream must be exactly the
    Philox / Box-Muller draws of simdata/philox.py for (noise seed, stream, event index, component), scaled by the reference's standard
    deviations; ids are untouched (noise is added after the visibility test); instances with different seeds get different draws;
    the draws do not depend on how the calls are batched."""
    from eqvio_b200.simulator import DeviceSimulator
    from simdata import philox

    N, frames = 32, 8
    seeds = (0, 7)
    clean = DeviceSimulator([SimConfig.benchmark(N, s) for s in seeds])
    noisy = DeviceSimulator([SimConfig.benchmark(N, s, inputNoise=True, outputNoise=True, noiseSeed=100 + s) for s in seeds])
    a, b = clean.record_streams(frames), noisy.record_streams(frames)
    c0 = noisy.cfg
    for inst, (sa, sb) in enumerate(zip(a, b)):
        seed = 100 + seeds[inst]
        for fa, fb in zip(sa.frames, sb.frames):
            assert np.array_equal(fa.ids, fb.ids) and fa.stamp == fb.stamp
            np.testing.assert_array_equal(fa.provided_p, fb.provided_p)  # the truth handed to augmentLandmarkStates stays clean
            ev = int(round(fa.stamp * c0.imageFreq))
            z0, z1 = philox.normal_pair(seed, philox.STREAM_VISION, ev, np.arange(len(fa.ids)))
            np.testing.assert_allclose(fb.y - fa.y, c0.measurementNoise * np.stack([z0, z1], axis=1), rtol=0, atol=1e-9)
            if len(fa.imu):
                evi = np.rint(fa.imu[:, 0] * c0.imuFreq).astype(np.int64)
                z0, z1 = philox.normal_pair(seed, philox.STREAM_IMU, evi[:, None], np.arange(6)[None, :])
                z = np.stack([z0, z1], axis=2).reshape(len(evi), 12)
                np.testing.assert_allclose(fb.imu[:, 1:] - fa.imu[:, 1:], noisy.imu_sigma * z, rtol=0, atol=1e-9)
                assert np.array_equal(fa.imu[:, 0], fb.imu[:, 0])
    assert not np.allclose(b[0].frames[3].imu[:, 1:4], b[1].frames[3].imu[:, 1:4])  # per-instance draws
    # batching independence: one stamp at a time gives the same pixels
    t = np.array([fr.stamp for fr in b[0].frames])
    _, _, y_all, _, _ = noisy.vision(t)
    _, _, y_one, _, _ = noisy.vision(t[5:6])
    np.testing.assert_array_equal(y_all[:, 5], y_one[:, 0])
    clean.close()
    noisy.close()


def test_noisy_device_stream_drives_the_filter():
    """A noisy Monte-Carlo instance from the device simulator through the filter (gating off as in the struct defaults): finite,
    and the landmark set follows the measured ids."""
    import eqvio_b200 as eb
    from eqvio_b200.simulator import DeviceSimulator

    cfg = SimConfig.benchmark(32, 2, inputNoise=True, outputNoise=True)
    sim = DeviceSimulator([cfg])
    sm = sim.record_streams(15)[0]
    flt = eb.VIOFilter(eb.Settings(fastRiccati=1), eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0, capacity=40)
    cam = eb.Camera(**sm.camera)
    for fr in sm.frames:
        flt.processIMUArray(fr.imu)
        flt.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        flt.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
    est = flt.stateEstimate()
    assert np.isfinite(est.sensor.flat()).all() and np.isfinite(est.p).all()
    assert np.array_equal(np.sort(np.asarray(est.ids)), np.sort(sm.frames[-1].ids))
    err = np.linalg.norm(est.sensor.flat()[10:13] - sm.frames[-1].true_sensor[10:13])
    assert err < 0.5, err
    flt.close()
    sim.close()
