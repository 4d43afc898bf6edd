"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
VIOSimulator streams, plus the reference's edge cases (silent returns, empty / ragged measurements,
capacity, unsupported switches) and size-independent properties at BASELINE's full size.

Tolerances: BASELINE.json asks for pose / Sigma within 1e-6 relative Frobenius and identical landmark
indexing.  The path is fp64 throughout, so sequences are held to 1e-9 here.
"""
import numpy as np
import pytest

from parity_utils import compare_states, gpu_filter, make_stream, replay_gpu, run_gpu, run_oracle, snapshot_gpu

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _check(gpu, ref, tol=TOL):
    assert len(gpu) == len(ref)
    worst = 0.0
    for k, (g, r) in enumerate(zip(gpu, ref)):
        e = compare_states(g, r)
        assert e["ids_equal"], f"update {k}: landmark ids differ"
        assert e["sigma"] < tol, f"update {k}: Sigma rel-Frobenius {e['sigma']:.3e}"
        assert e["state"] < tol, f"update {k}: state rel-Frobenius {e['state']:.3e}"
        assert e["pose"] < tol, f"update {k}: pose rel-Frobenius {e['pose']:.3e}"
        worst = max(worst, e["sigma"], e["state"])
    return worst


@pytest.mark.parametrize("discrete", [1, 0])
def test_fused_observer_matches_two_kernel_form(discrete):
    """integrateObserverState as one software-pipelined kernel (helper warp / sensor chain / landmark warps) vs the sensor kernel
    followed by the landmark kernel, which keeps the reference's operation order: the pipelined form carries the moved point and shares
    reciprocal norms, so the two agree to rounding (a few ulp per segment), and both are checked against the oracle."""
    stream = make_stream(N=70, frames=8, coord=0, settings_overrides=dict(useDiscreteVelocityLift=bool(discrete)))
    ref = run_gpu(stream, tuning=dict(fuseObserver=0))
    got = run_gpu(stream, tuning=dict(fuseObserver=1))
    for g, r in zip(got, ref):
        e = compare_states(g, r)
        assert e["ids_equal"] and e["sigma"] < 1e-11 and e["state"] < 1e-11
    _check(got, run_oracle(stream))
    _check(ref, run_oracle(stream))


@pytest.mark.parametrize("N,chunk", [(40, 8), (100, 32), (150, 32)])
@pytest.mark.parametrize("graph", [0, 1])
def test_lookahead_downdate_is_bit_identical(N, chunk, graph):
    """The look-ahead split (band tiles at once, the other tiles beside the next factor kernel on a second stream)
    runs the same arithmetic per tile as one in-order downdate per chunk: identical bits, with and without graphs."""
    stream = make_stream(N=N, frames=8, coord=0)
    ref = run_gpu(stream, tuning=dict(lookahead=0, graph=0, chunkLandmarks=chunk))
    got = run_gpu(stream, tuning=dict(lookahead=1, graph=graph, chunkLandmarks=chunk))
    for g, r in zip(got, ref):
        e = compare_states(g, r)
        assert e["ids_equal"] and e["sigma"] == 0.0 and e["state"] == 0.0
    _check(got, run_oracle(stream))


@pytest.mark.parametrize("coord", [0, 1])
@pytest.mark.parametrize("N", [8, 64])
def test_sequence_matches_oracle(N, coord):
    stream = make_stream(N=N, frames=8, coord=coord)
    _check(run_gpu(stream), run_oracle(stream))


@pytest.mark.parametrize("tuning", [dict(correction=1), dict(correction=2), dict(correction=2, graph=0), dict(correction=2, graph=0, speculate=0),
                                    dict(correction=0, chunkLandmarks=5), dict(correction=0, chunkLandmarks=16),
                                    dict(correction=0, chunkLandmarks=1), dict(correction=0, chunkLandmarks=7)])
def test_correction_evaluation_orders_agree(tuning):
    """Batch Cholesky sweep, block sweep with look-ahead (blockchol.cuh) and sequential chunks of any size: same result to
    rounding, all match the oracle.  N = 40: one full 64-row block and a ragged one of 16 rows."""
    stream = make_stream(N=40, frames=6, coord=1)
    ref = run_oracle(stream)
    _check(run_gpu(stream, tuning=tuning), ref)


@pytest.mark.parametrize("tuning", [dict(graph=0), dict(graph=0, speculate=0), dict(graph=1), dict(graph=1, pdl=0), dict(graph=0, pdl=0), dict(graph=1, fuseSmall=0), dict(graph=0, fuseSmall=0),
                                    dict(graph=1, propFusion=0), dict(graph=0, propFusion=0)])
def test_steady_path_variants_agree(tuning):
    """CUDA-graph replay, plain speculative launches and the wait-for-the-gate path give identical results
    (same kernels, same order), over enough frames for graphs to be captured AND replayed."""
    stream = make_stream(N=40, frames=12, coord=0)
    ref = run_gpu(stream, tuning=dict(graph=0, speculate=0))
    got = run_gpu(stream, tuning=tuning)
    for g, r in zip(got, ref):
        e = compare_states(g, r)
        assert e["ids_equal"] and e["sigma"] == 0.0 and e["state"] == 0.0


@pytest.mark.parametrize("graph", [0, 1])
def test_staged_s_gather(graph):
    """EQVIO_TUNE_STAGE_S (default on): the chunk factor kernel fetches Sigma[L_c, L_c] as one 2-D TMA tensor copy instead of
    per-thread gathers.  The same entries enter the same arithmetic, so both settings agree bit for bit.  N = 40: one full
    chunk of 32 landmarks and a ragged one of 8; landmark-set changes make some chunks non-contiguous in the state, which falls
    back to the gather inside the same launch sequence."""
    stream = make_stream(N=40, frames=12, coord=0)
    ref = run_gpu(stream, tuning=dict(graph=0, speculate=0, stageS=0))
    got = run_gpu(stream, tuning=dict(graph=graph, stageS=1))
    for g, r in zip(got, ref):
        e = compare_states(g, r)
        assert e["ids_equal"] and e["sigma"] == 0.0 and e["state"] == 0.0
    _check(got, run_oracle(stream))


@pytest.mark.parametrize("coord", [0, 1])
def test_continuous_lifts(coord):
    """useDiscreteVelocityLift = useDiscreteInnovationLift = false (the EuRoC config's innovation lift)."""
    stream = make_stream(N=24, frames=6, coord=coord,
                         settings_overrides=dict(useDiscreteVelocityLift=False, useDiscreteInnovationLift=False))
    _check(run_gpu(stream), run_oracle(stream))


def test_non_equivariant_output():
    stream = make_stream(N=24, frames=5, coord=0, settings_overrides=dict(useEquivariantOutput=False))
    _check(run_gpu(stream), run_oracle(stream))


@pytest.mark.parametrize("speculate", [1, 0])
def test_gating_and_noise_identical_indexing(speculate):
    """Outlier gating (absolute + probabilistic, capped by featureRetention) with noisy inputs: the
    discrete decisions must match the oracle's, update by update -- both when the correction is launched
    speculatively (and redone after a gate hit) and when the host waits for the gate."""
    stream = make_stream(N=48, frames=10, coord=0,
                         settings_overrides=dict(outlierThresholdAbs=3.0, outlierThresholdProb=6.0, measurementNoise=0.5,
                                                 featureRetention=0.2),
                         sim_overrides=dict(outputNoise=True, inputNoise=True))
    gpu, ref = run_gpu(stream, tuning=dict(speculate=speculate)), run_oracle(stream)
    _check(gpu, ref)
    assert any(len(r["ids"]) < 48 for r in ref), "the case is meant to exercise outlier removal"


def test_new_landmarks_from_bearings_median_depth():
    """Without augmentLandmarkStates new ids enter through addNewLandmarks: bearing x median depth."""
    import eqvio_b200 as eb
    from oracle import eqf

    stream = make_stream(N=32, frames=8, coord=0)
    o = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
    g, cam = gpu_filter(stream)
    # drop a third of the initial landmarks so that later frames re-introduce them as new ids
    keep = stream["init"].ids[::3]
    o.removeOldLandmarks(list(keep))
    g.augmentLandmarkStates(keep, eb.VIOState(eb.VIOSensorState(), stream["init"].p, stream["init"].ids))
    for fr in stream["frames"]:
        for row in fr.imu:
            o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
        g.processIMUArray(fr.imu)
        o.processVisionData(eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, stream["cam"]))
        g.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        from parity_utils import snapshot_oracle
        e = compare_states(snapshot_gpu(g), snapshot_oracle(o))
        assert e["ids_equal"] and e["sigma"] < TOL and e["state"] < TOL
    assert g.numLandmarks() == 32
    g.close()


@pytest.mark.parametrize("speculateNew", [1, 0])
@pytest.mark.parametrize("noisy", [False, True])
def test_real_data_flow_new_and_lost_ids_inside_the_vision_call(speculateNew, noisy):
    """eqvio_opt's flow (no augmentLandmarkStates): lost ids are pruned and new ids added INSIDE processVisionData, on frames
    whose gates may trip as well (noisy case: outliers + new ids on the same frame exercise the drop-and-re-add redo of the
    speculative append).  Discrete decisions and values must match the oracle either way."""
    import eqvio_b200 as eb
    from oracle import eqf
    from parity_utils import snapshot_oracle

    ov = dict(outlierThresholdAbs=3.0, outlierThresholdProb=6.0, measurementNoise=0.5, featureRetention=0.2) if noisy else {}
    so = dict(outputNoise=True, inputNoise=True) if noisy else {}
    stream = make_stream(N=40, frames=12, coord=0, settings_overrides=ov, sim_overrides=so)
    o = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
    g, cam = gpu_filter(stream)
    g.setTuning(speculateNew=speculateNew)
    removed = 0
    for fr in stream["frames"]:
        for row in fr.imu:
            o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
        g.processIMUArray(fr.imu)
        o.processVisionData(eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, stream["cam"]))
        g.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        e = compare_states(snapshot_gpu(g), snapshot_oracle(o))
        assert e["ids_equal"] and e["sigma"] < TOL and e["state"] < TOL
        removed += len(g.lastOutliers())
    if noisy:
        assert removed > 0, "the noisy case is meant to trip gates"
    g.close()


def test_keep_lost_landmarks():
    """removeLostLandmarks = false: unmeasured landmarks stay in the state with zero C columns."""
    import eqvio_b200 as eb
    from oracle import eqf
    from parity_utils import snapshot_oracle

    stream = make_stream(N=24, frames=5, coord=0, settings_overrides=dict(removeLostLandmarks=False))
    o = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
    g, cam = gpu_filter(stream)
    for fr in stream["frames"]:
        for row in fr.imu:
            o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
        g.processIMUArray(fr.imu)
        sel = np.arange(len(fr.ids)) % 4 != 1  # ragged: a quarter of the tracks missing each frame
        o.processVisionData(eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids[sel], fr.y[sel], stream["cam"]))
        g.processVisionArrays(fr.stamp, fr.ids[sel], fr.y[sel], cam)
        e = compare_states(snapshot_gpu(g), snapshot_oracle(o))
        assert e["ids_equal"] and e["sigma"] < TOL and e["state"] < TOL
    g.close()


@pytest.mark.parametrize("ncoef", [5, 4])
def test_radtan_camera(ncoef):
    """Radtan camera with five and with FOUR distortion coefficients (EuRoC's sensor.yaml gives four): the inverse model always has
    five (StandardCamera.cpp:117-147), so the undistortion of a four-coefficient camera must still apply the r^6 term."""
    from oracle.camera import StandardCamera

    stream = make_stream(N=24, frames=1, coord=1)
    dist = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0][:ncoef]
    cam = StandardCamera(752, 480, 458.654, 457.296, 367.215, 248.375, dist)
    stream["cam"] = cam
    # re-project the recorded measurements through the distorted camera
    from oracle.simulator import SimulationDataServer, benchmarkSim
    server = SimulationDataServer(benchmarkSim(24, 0), stream["settings"])
    server.simulator.cameraPtr = cam
    stream["init"] = server.initialCondition()
    stream["frames"] = server.record(6)
    _check(run_gpu(stream), run_oracle(stream))


def test_equidistant_camera():
    """GIFT EquidistantCamera (fisheye): iterative undistort + its Jacobian inside C* on the device."""
    from oracle.camera import EquidistantCamera
    from oracle.simulator import SimulationDataServer, benchmarkSim

    for coord in (0, 1):
        stream = make_stream(N=24, frames=1, coord=coord)
        cam = EquidistantCamera(752, 480, 458.654, 457.296, 367.215, 248.375,
                                [-0.013721808247486035, 0.020727425669427896, -0.012786476702685545, 0.0025242267320687625])
        stream["cam"] = cam
        server = SimulationDataServer(benchmarkSim(24, 0), stream["settings"])
        server.simulator.cameraPtr = cam
        stream["init"] = server.initialCondition()
        stream["frames"] = server.record(6)
        _check(run_gpu(stream), run_oracle(stream))


def test_config2_n256():
    """BASELINE configs[1]: N=256 landmarks, fp64 Sigma, correctness vs the CPU reference path."""
    stream = make_stream(N=256, frames=4, coord=0)
    ref = run_oracle(stream, dense_lazy=True)
    assert _check(run_gpu(stream), ref) < 1e-9
    assert _check(run_gpu(stream, tuning=dict(correction=1)), ref) < 1e-9


def test_silent_returns_and_errors():
    import eqvio_b200 as eb

    stream = make_stream(N=8, frames=3, coord=0)
    cam = eb.Camera.fromPod(stream["cam"].pod())
    # ctor #1: not initialised, no IMU -> processVisionData returns silently
    f = eb.VIOFilter(eb.Settings(fastRiccati=1), capacity=16)
    assert not f.isInitialised() and f.getTime() == -1.0
    fr = stream["frames"][1]
    assert f.processVisionArrays(fr.stamp, fr.ids, fr.y, cam) is False
    assert f.numLandmarks() == 0
    # first IMU sample initialises attitude from gravity (VIOFilter.cpp:65-78)
    f.processIMUArray(fr.imu[:1])
    assert f.isInitialised() and f.getTime() == fr.imu[0, 0]
    # time not advanced -> silent return
    assert f.processVisionArrays(fr.imu[0, 0], fr.ids, fr.y, cam) is False
    # advancing with a measurement: landmarks are created from bearings at initialSceneDepth
    f.processIMUArray(fr.imu[1:])
    assert f.processVisionArrays(fr.stamp, fr.ids, fr.y, cam) is True
    assert f.numLandmarks() == len(fr.ids)
    # empty measurement: propagates, removes every landmark (removeLostLandmarks), no correction
    fr2 = stream["frames"][2]
    f.processIMUArray(fr2.imu)
    assert f.processVisionArrays(fr2.stamp, np.zeros(0, dtype=np.int32), np.zeros((0, 2)), cam) is False
    assert f.numLandmarks() == 0 and f.getTime() == fr2.stamp
    # capacity
    with pytest.raises(eb.EqvioError) as ei:
        big = np.arange(17, dtype=np.int32)
        f.processIMUArray(np.concatenate([[fr2.stamp + 0.01], np.zeros(3), [0, 0, 9.81], np.zeros(6)])[None])
        f.processVisionArrays(fr2.stamp + 0.05, big, np.full((17, 2), 100.0) + big[:, None], cam)
    assert ei.value.code == eb._capi.EQVIO_ERR_CAPACITY
    # unsorted ids
    with pytest.raises(eb.EqvioError) as ei:
        f.processVisionArrays(fr2.stamp + 1.0, np.array([3, 2], dtype=np.int32), np.zeros((2, 2)), cam)
    assert ei.value.code == eb._capi.EQVIO_ERR_INVALID_ARG
    f.close()
    # switches without a CUDA path fail loudly
    with pytest.raises(eb.EqvioError) as ei:
        eb.VIOFilter(eb.Settings(coordinateChoice=3), capacity=4)
    assert ei.value.code == eb._capi.EQVIO_ERR_UNSUPPORTED


def test_from_imu_initialisation_matches_oracle():
    """ctor #1 + initialiseFromIMUData + addNewLandmarks at initialSceneDepth, against the oracle."""
    import eqvio_b200 as eb
    from oracle import eqf
    from parity_utils import snapshot_oracle

    stream = make_stream(N=16, frames=5, coord=1, settings_overrides=dict(initialSceneDepth=3.0))
    ost = stream["settings"]
    o = eqf.VIOFilter(ost)
    g = eb.VIOFilter(eb.Settings.fromObject(ost), capacity=32)
    cam = eb.Camera.fromPod(stream["cam"].pod())
    for fr in stream["frames"][1:]:
        for row in fr.imu:
            o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
        g.processIMUArray(fr.imu)
        o.processVisionData(eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, stream["cam"]))
        g.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        e = compare_states(snapshot_gpu(g), snapshot_oracle(o))
        assert e["ids_equal"] and e["sigma"] < TOL and e["state"] < TOL
    g.close()


def test_set_state_and_set_landmarks():
    import eqvio_b200 as eb
    from oracle import eqf
    from parity_utils import snapshot_oracle

    stream = make_stream(N=12, frames=3, coord=0, settings_overrides=dict(initialPointDepthVariance=0.25))
    o = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
    g, cam = gpu_filter(stream)
    init = stream["init"]
    o.setState(init)
    g.setState(eb.VIOState(eb.VIOSensorState.fromFlat(init.sensor.flat()), init.p, init.ids))
    e = compare_states(snapshot_gpu(g), snapshot_oracle(o))
    assert e["ids_equal"] and e["sigma"] == 0.0 and e["state"] < 1e-15
    o.setLandmarks(init.p * 1.1, init.ids)
    g.setLandmarks(init.p * 1.1, init.ids)
    e = compare_states(snapshot_gpu(g), snapshot_oracle(o))
    assert e["ids_equal"] and e["sigma"] == 0.0 and e["state"] < 1e-15
    blocks = g.landmarkCovBlocks()
    assert np.allclose(blocks[:, 2, 2], 0.25) and np.allclose(blocks[:, 0, 0], stream["settings"].initialPointVariance)
    g.close()


def test_batch_replicas_match_single():
    """eqvio_batch_process_vision over independent Monte-Carlo replicas == one-by-one processing."""
    import eqvio_b200 as eb

    streams = [make_stream(N=16, frames=5, coord=0, seed=s) for s in range(3)]
    singles = [run_gpu(s) for s in streams]
    pairs = [gpu_filter(s) for s in streams]
    filters = [p[0] for p in pairs]
    cam = pairs[0][1]
    for k in range(5):
        for f, s in zip(filters, streams):
            fr = s["frames"][k]
            f.processIMUArray(fr.imu)
            f.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        eb.batchProcessVision(filters, [s["frames"][k].stamp for s in streams], [s["frames"][k].ids for s in streams],
                              [s["frames"][k].y for s in streams], cam)
        for f, single in zip(filters, singles):
            e = compare_states(snapshot_gpu(f), single[k])
            assert e["ids_equal"] and e["sigma"] == 0.0 and e["state"] == 0.0
    for f in filters:
        f.close()


def test_full_size_properties_n1024():
    """BASELINE's largest size (N=1024, dim 3093): size-independent properties of one propagate +
    correct step -- Sigma stays symmetric, positive definite on a probe, and the correction never
    increases the trace; the state matches the oracle's structured evaluation on one update."""
    stream = make_stream(N=1024, frames=3, coord=0)
    flt, cam = gpu_filter(stream)
    traces = []

    def cb(k, f):
        s = snapshot_gpu(f)
        S = s["Sigma"]
        assert np.isfinite(S).all()
        assert np.abs(S - S.T).max() <= 1e-12 * np.abs(S).max()
        traces.append(np.trace(S))
        rng = np.random.default_rng(k)
        v = rng.standard_normal((S.shape[0], 8))
        assert (np.einsum("ij,ij->j", v, S @ v) > 0).all()

    replay_gpu(flt, cam, stream["frames"], cb)
    assert flt.numLandmarks() == 1024
    assert traces[1] < traces[0] and traces[2] < traces[1] * 1.01
    flt.close()


def test_feature_predictions_match_oracle():
    """getFeaturePredictions / predictState (VIO_eqf.cpp:139-151) with useFeaturePredictions on."""
    import eqvio_b200 as eb
    from oracle import eqf
    from parity_utils import snapshot_oracle

    stream = make_stream(N=24, frames=4, coord=0, settings_overrides=dict(useFeaturePredictions=True))
    o = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
    g, cam = gpu_filter(stream)
    for k, fr in enumerate(stream["frames"]):
        for row in fr.imu:
            o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
        g.processIMUArray(fr.imu)
        if k > 0:
            # the tracker asks for predictions at the new image stamp before the update (main_opt.cpp:205-206)
            po = o.getFeaturePredictions(stream["cam"], fr.stamp)
            pg = g.getFeaturePredictions(cam, fr.stamp)
            assert sorted(po.camCoordinates) == sorted(pg.camCoordinates) and len(pg.camCoordinates) > 0
            for i, px in po.camCoordinates.items():
                assert np.abs(pg.camCoordinates[i] - px).max() < 1e-9 * max(1.0, np.abs(px).max())
        o.augmentLandmarkStates(list(fr.ids), eqf.VIOState(None, fr.provided_p, fr.ids))
        g.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        o.processVisionData(eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, stream["cam"]))
        g.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        e = compare_states(snapshot_gpu(g), snapshot_oracle(o))
        assert e["ids_equal"] and e["sigma"] < TOL and e["state"] < TOL
    # switched off: an empty measurement, like the reference
    g2, cam2 = gpu_filter(make_stream(N=8, frames=2, coord=0))
    assert g2.getFeaturePredictions(cam2, 0.1).camCoordinates == {}
    g.close()
    g2.close()


@pytest.mark.parametrize("coord", [0, 1, 2])
def test_nees_matches_oracle(coord):
    """computeNEES (VIO_eqf.cpp:153-170): device Cholesky solve vs the oracle's dense inverse, along a sequence."""
    import eqvio_b200 as eb
    from oracle import eqf
    from oracle.simulator import SimulationDataServer, benchmarkSim

    stream = make_stream(N=20, frames=6, coord=coord)
    server = SimulationDataServer(benchmarkSim(20, 0), stream["settings"])
    o = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
    g, cam = gpu_filter(stream)
    for fr in stream["frames"]:
        for row in fr.imu:
            o.processIMUData(eqf.IMUVelocity(row[0], row[1:4], row[4:7], row[7:10], row[10:13]))
        g.processIMUArray(fr.imu)
        o.augmentLandmarkStates(list(fr.ids), eqf.VIOState(None, fr.provided_p, fr.ids))
        g.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
        o.processVisionData(eqf.VisionMeasurement.fromArrays(fr.stamp, fr.ids, fr.y, stream["cam"]))
        g.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        true = server.getTrueState(o.getTime())
        # perturb the truth a little so that the error vector is not ~0
        true.sensor.inputBias = true.sensor.inputBias + 0.01
        true.p = true.p * 1.01
        no = o.viewEqFState().computeNEES(true)
        ng = g.computeNEES(eb.VIOState(eb.VIOSensorState.fromFlat(true.sensor.flat()), true.p, true.ids))
        assert np.isfinite(ng) and abs(ng - no) <= (1e-6 if coord == 2 else 1e-8) * max(1.0, abs(no)), (ng, no)
    g.close()


@pytest.mark.parametrize("coord", [0, 1])
def test_accurate_riccati_default_settings(coord):
    """fastRiccati = false, the struct default: integrateRiccatiStateAccurate per IMU sample (VIO_eqf.cpp:74-91,
    VIOFilter.cpp:160-178) -- dense matrix exponential on the device vs the oracle's scipy expm."""
    stream = make_stream(N=12, frames=4, coord=coord, settings_overrides=dict(fastRiccati=False))
    _check(run_gpu(stream), run_oracle(stream), tol=1e-8)


@pytest.mark.parametrize("coord,lift", [(0, True), (1, True), (0, False), (2, True), (2, False)])
def test_discrete_state_matrix(coord, lift):
    """fastRiccati = false with useDiscreteStateMatrix: integrateRiccatiStateDiscrete per IMU sample (VIO_eqf.cpp:93-103)
    with the numerically differentiated stateMatrixADiscrete (EqFMatrices.cpp:24-41; central differences, h = cbrt(eps):
    rounding in the function values is amplified by 1 / 2h ~ 8e4, hence the looser tolerance).  The discrete state matrix
    always differentiates liftVelocityDiscrete, also when the observer integrates with the continuous lift (lift=False), and
    it is taken in the chart of the coordinate suite (coord 2 = Normal: B still goes through M)."""
    stream = make_stream(N=12, frames=4, coord=coord, settings_overrides=dict(fastRiccati=False, useDiscreteStateMatrix=True,
                                                                             useDiscreteVelocityLift=lift))
    _check(run_gpu(stream), run_oracle(stream), tol=1e-7)


@pytest.mark.parametrize("overrides", [dict(), dict(useDiscreteInnovationLift=False), dict(useDiscreteVelocityLift=False),
                                       dict(fastRiccati=False), dict(useEquivariantOutput=False)])
def test_normal_coordinates(overrides):
    """coordinateChoice = Normal (coordinateSuite/normal.cpp): A = M A_euclid M^-1, B = M B_euclid with the numerically
    differentiated chart change M (VIOState.cpp:391-401), its own output block, innovation lifts through M^-1 / the chart
    maps.  Dense propagation on the device vs the oracle; M carries ~1e-10 of differentiation noise in both."""
    stream = make_stream(N=10, frames=5, coord=2, settings_overrides=overrides)
    _check(run_gpu(stream), run_oracle(stream), tol=1e-7)


def test_writer_consistency_files(tmp_path):
    """VIOWriter mirror on a live filter: states + NEES / consistency files (VIOWriter.cpp:33-81,142-228); the pose NEES written to
    nees.csv must equal the one recomputed from the oracle's filter state."""
    import eqvio_b200 as eb
    from oracle import liegroups as lg

    stream = make_stream(N=12, frames=4, coord=0)
    ref = run_oracle(stream)
    g, cam = gpu_filter(stream)
    with eb.VIOWriter(str(tmp_path)) as w:
        for fr in stream["frames"]:
            g.processIMUArray(fr.imu)
            g.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
            g.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
            est = g.stateEstimate()
            w.writeStates(fr.stamp, est)
            w.writeFeatures(fr.stamp, fr.ids, fr.y)
            true = eb.VIOState(est.sensor, est.p * 1.01, est.ids)  # a "truth" 1% away in the landmarks, same sensor state
            w.writeConsistency(fr.stamp, true, g)
    rows = np.loadtxt(str(tmp_path / "nees.csv"), delimiter=",", skiprows=1, ndmin=2)
    assert rows.shape == (len(stream["frames"]), 5) and np.isfinite(rows).all() and (rows[:, 1] > 0).all()
    assert rows[-1, 2] == 21 + 3 * len(ref[-1]["ids"])
    imu = np.loadtxt(str(tmp_path / "IMUState.csv"), delimiter=",", skiprows=1, ndmin=2)
    np.testing.assert_allclose(imu[-1, 1:4], ref[-1]["sensor"][10:13], rtol=2e-5, atol=1e-6)  # six significant digits
    tru = np.loadtxt(str(tmp_path / "trueState.csv"), delimiter=",", skiprows=1, ndmin=2)  # VIOWriter.cpp:144-156
    n = len(ref[-1]["ids"])
    assert tru.shape == (len(stream["frames"]), 1 + 23 + 1 + 4 * n) and tru[-1, 24] == n
    np.testing.assert_array_equal(tru[-1, 25::4].astype(int), np.asarray(ref[-1]["ids"]))
    g.close()


def test_replay_batch_matches_single_replay():
    """eqvio_replay_batch (one C++ host thread per filter, kernels of the replicas overlapping on the GPU) gives every filter
    exactly the estimates of its own eqvio_replay -- and those match the Python-driven calls."""
    import eqvio_b200 as eb
    from simdata import SimConfig, record_stream

    R, K = 4, 10
    streams = [record_stream(SimConfig.benchmark(40, seed), K) for seed in range(R)]
    cam = eb.Camera(**streams[0].camera)

    def make(sm):
        return eb.VIOFilter(eb.Settings(fastRiccati=1), eb.VIOState(eb.VIOSensorState.fromFlat(sm.init_sensor), sm.init_p, sm.init_ids), 0.0,
                            capacity=48)

    batch = [make(sm) for sm in streams]
    _, est_b, wall = eb.replayBatch(batch, [sm.frames for sm in streams], cam)
    assert wall > 0
    for k, sm in enumerate(streams):
        single = make(sm)
        _, est_s = single.replay(sm.frames, cam)
        assert np.array_equal(est_b[k], est_s)
        py = make(sm)
        for fr in sm.frames:
            py.processIMUArray(fr.imu)
            py.augmentLandmarkStates(fr.ids, eb.VIOState(eb.VIOSensorState(), fr.provided_p, fr.ids))
            py.processVisionArrays(fr.stamp, fr.ids, fr.y, cam)
        assert np.array_equal(py.stateEstimate().sensor.flat(), est_s[-1])
        assert np.array_equal(batch[k].stateEstimate().p, py.stateEstimate().p)
        for f_ in (single, py):
            f_.close()
    for f_ in batch:
        f_.close()


def test_tcgen05_probe():
    """Stand-alone tcgen05 / TMEM / UMMA-descriptor probe (tests/csrc/tc_probe.cu): 128x128x64 bf16 GEMM, exact vs CPU."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "lib", "tc_probe")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-o", exe,
                    os.path.join(root, "tests", "csrc", "tc_probe.cu")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("N", [40, 100])
def test_tensor_core_downdate_close_to_fp64(N):
    """BASELINE configs[2] arithmetic: tcgen05 downdate with split-bf16 operands (24-bit mantissa) and fp32 accumulation
    in TMEM.  Not an fp64 path: the downdate Sigma - Y^T Y cancels (the first updates shrink the landmark covariance
    by ~1e3), so the fp32-accumulated product leaves ~1e-7 * |Sigma_old| / |Sigma_new| relative error in Sigma --
    observed ~5e-4 right after initialisation, smaller once the filter has converged.  Landmark ids stay identical and
    the state estimate stays within 1e-4 of the oracle."""
    stream = make_stream(N=N, frames=8, coord=0)
    ref = run_oracle(stream)
    got = run_gpu(stream, tuning=dict(downdate=1))
    worst_sigma = worst_state = 0.0
    for k, (g, r) in enumerate(zip(got, ref)):
        e = compare_states(g, r)
        assert e["ids_equal"], f"update {k}: landmark ids differ"
        worst_sigma, worst_state = max(worst_sigma, e["sigma"]), max(worst_state, e["state"])
    print(f"tcgen05 downdate N={N}: worst rel-Frobenius vs oracle: Sigma {worst_sigma:.3e}, state {worst_state:.3e}")
    assert worst_sigma < 5e-3 and worst_state < 1e-4
    S = got[-1]["Sigma"]
    assert np.abs(S - S.T).max() <= 1e-6 * np.abs(S).max() and np.all(np.linalg.eigvalsh(0.5 * (S + S.T)) > 0)


def test_long_sequence_no_drift():
    """150 consecutive updates (7.5 s of the simulated lap, landmarks entering and leaving all the time) through the
    default path -- CUDA graphs, speculation, compaction -- against the oracle: BASELINE's 1e-6 bound holds with a wide
    margin along the whole sequence, landmark ids identical after every update."""
    stream = make_stream(N=32, frames=150, coord=0)
    got, ref = run_gpu(stream), run_oracle(stream)
    worst = _check(got, ref, tol=1e-6)
    changed = sum(1 for a, b in zip(ref[:-1], ref[1:]) if not np.array_equal(a["ids"], b["ids"]))
    print(f"long sequence: worst rel-Frobenius {worst:.3e} over {len(ref)} updates, landmark set changed {changed} times")
    assert worst < 1e-9 and changed > 10
