"""Launch plan of the lazy trailing updates (N > 384: sequential chunks with look-ahead, EQVIO_TUNE_LAZY_DOWNDATE = M), checked on the
CPU through eqvio_plan_lazy_downdates -- the plan eqvio_process_vision executes (csrc/filter.cu: plan_lazy_downdates).

Model of the execution: the factor of chunk c, the wait it may carry and its urgent launch run in order on the filter's stream; a
deferred launch is issued behind the urgent launch of its chunk on the side stream (in order among themselves) and may still be running
until some later launch on the filter's stream waits for it.  Invariants:
  * a launch that may overlap a deferred launch shares no tile with it -- neither an urgent launch (writes) nor a factor (reads the
    tile rows / columns of its landmarks);
  * when a factor gathers, every tile it reads carries all earlier chunks; when the last launch has run, every tile carries all chunks;
  * tile counts match the kernel's decode (grid sizes)."""
import ctypes as C

import numpy as np
import pytest


def _plan(T, blo, bhi, M):
    from eqvio_b200._capi import lib

    n = len(blo)
    out = (C.c_int * (8 * n))()
    lo = (C.c_int * n)(*[int(v) for v in blo])
    hi = (C.c_int * n)(*[int(v) for v in bhi])
    rc = lib.eqvio_plan_lazy_downdates(T, n, lo, hi, M, out, n)
    assert rc == n
    return np.array(out[:], dtype=np.int64).reshape(n, 8)


def _band_tiles(T, lo, hi):
    return {(ti, tj) for ti in range(T) for tj in range(ti + 1) if lo <= ti <= hi or (ti > hi and lo <= tj <= hi)}


def _rest_tiles(T, lo, hi):
    return {(ti, tj) for ti in range(T) for tj in range(ti + 1) if not (lo <= ti <= hi) and not (lo <= tj <= hi)}


def _bands(rng, T, nchunks, sparse):
    """Tile rows of consecutive landmark chunks (96 state rows each, ascending; with gaps when not every landmark is measured)."""
    rows_total = 64 * T
    starts = np.sort(rng.choice(np.arange(21, rows_total - 96), size=nchunks, replace=False)) if sparse else 21 + 96 * np.arange(nchunks)
    lo, hi = [], []
    for c in range(nchunks - 1):  # band c = rows of chunk c + 1
        s = int(starts[c + 1])
        lo.append(s // 64)
        hi.append(min(T - 1, (s + 95) // 64))
    return lo + [0], hi + [0]


def _check(T, blo, bhi, M):
    n = len(blo)
    plan = _plan(T, blo, bhi, M)
    level = {(ti, tj): 0 for ti in range(T) for tj in range(ti + 1)}
    open_rest = {}  # chunk -> tile set of a deferred launch nobody has waited for yet
    for c in range(n):
        lo, hi, n_band, wait, has_rest, xlo, xhi, n_rest = (int(v) for v in plan[c])
        if c >= 1:  # factor(c) gathers the tile rows / columns of band c - 1
            reads = _band_tiles(T, int(plan[c - 1][0]), int(plan[c - 1][1]))
            for r, tiles in open_rest.items():
                assert not (reads & tiles), f"factor({c}) reads tiles the deferred launch of chunk {r} may be writing"
            assert all(level[t] == c for t in reads), f"factor({c}) would gather tiles that miss earlier chunks"
        if wait >= 0:
            assert wait in open_rest or not open_rest or wait < min(open_rest), f"chunk {c} waits for a deferred launch never issued"
            for r in [r for r in open_rest if r <= wait]:  # the side stream is in order
                del open_rest[r]
        if c == n - 1:
            assert not open_rest, "the final launch overlaps a deferred launch"
            assert n_band == T * (T + 1) // 2
            for t in level:
                level[t] = n
            break
        band = _band_tiles(T, lo, hi)
        assert len(band) == n_band
        for r, tiles in open_rest.items():
            assert not (band & tiles), f"urgent launch of chunk {c} shares tiles with the deferred launch of chunk {r}"
        for t in band:
            level[t] = c + 1
        if has_rest:
            rest = _rest_tiles(T, xlo, xhi)
            assert len(rest) == n_rest and xlo == lo and not (rest & band)
            for t in rest:
                level[t] = c + 1
            open_rest[c] = rest
    assert all(v == n for v in level.values())
    return plan


@pytest.mark.parametrize("M", [1, 2, 3, 5, 8])
def test_lazy_plan_contiguous_chunks(M):
    """BASELINE sizes: N = 1024 (T = 49, 32 chunks) and N = 512 (T = 25, 16 chunks), every landmark measured."""
    for T, n in ((49, 32), (25, 16), (6, 3)):
        lo, hi = _bands(None, T, n, sparse=False)
        plan = _check(T, lo, hi, M)
        assert plan[:, 4].sum() == (n - 1) // M  # one deferred launch per M chunks


@pytest.mark.parametrize("seed", range(12))
def test_lazy_plan_random_sparse_bands(seed):
    """Unmeasured landmarks leave gaps between the bands of consecutive chunks; any M, any size."""
    rng = np.random.default_rng(seed)
    T = int(rng.integers(8, 40))
    n = int(rng.integers(3, max(4, (64 * T - 120) // 100)))
    lo, hi = _bands(rng, T, n, sparse=True)
    _check(T, lo, hi, int(rng.integers(1, 7)))


def test_lazy_plan_rejects_bad_arguments():
    from eqvio_b200._capi import lib

    out = (C.c_int * 16)()
    z = (C.c_int * 2)(0, 0)
    assert lib.eqvio_plan_lazy_downdates(4, 2, z, z, 0, out, 2) < 0
    assert lib.eqvio_plan_lazy_downdates(4, 2, z, z, 1, out, 1) < 0
    assert lib.eqvio_plan_lazy_downdates(4, 2, None, z, 1, out, 2) < 0
