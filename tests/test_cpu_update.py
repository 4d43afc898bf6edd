"""oracle/cpu_update.c (the C restatement of the reference's dense linear algebra that bench.py times as the CPU baseline) against
the numpy oracle: both evaluation orders, per update, on recorded operands of a simulated sequence (no GPU needed)."""
import numpy as np
import pytest

from parity_utils import make_stream


@pytest.mark.parametrize("N,coord", [(24, 0), (40, 1)])
def test_c_baseline_matches_oracle(N, coord):
    from oracle import cpu_baseline as cb
    from oracle import eqf

    stream = make_stream(N=N, frames=6, coord=coord)
    flt = eqf.VIOFilter(stream["settings"], stream["init"], 0.0)
    ups = cb.record_updates(flt, stream["frames"], stream["cam"])
    assert len(ups) == 5  # the t = 0 image only augments
    for structured in (False, True):
        for threads in (1, 2):
            r = cb.run_updates(ups, structured=structured, threads=threads)
            assert r["updates"] == 5 and r["threads"] == threads and r["updates_per_s"] > 0
            assert r["worst_rel_error_vs_oracle"] < 1e-10, r
            assert set(r["stage_ms"]) == {"propagation", "preprocessing", "correction"}


def test_recorder_restores_the_oracle():
    from oracle import cpu_baseline as cb
    from oracle import eqf

    before = (eqf.VIO_eqf.integrateRiccatiStateFast, eqf.VIO_eqf.performVisionUpdate)
    with cb.Recorder():
        assert eqf.VIO_eqf.integrateRiccatiStateFast is not before[0]
    assert (eqf.VIO_eqf.integrateRiccatiStateFast, eqf.VIO_eqf.performVisionUpdate) == before
    assert np.isfinite(1.0)
