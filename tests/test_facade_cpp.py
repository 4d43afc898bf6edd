"""The C++ facade (include/eqvio_b200_facade.hpp): it compiles with plain g++ against the C ABI
(CPU check) and, on the GPU box, a C++ host program replays a golden stream through it the way
main_sim.cpp drives the reference VIOFilter and reproduces the golden outputs."""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry
from golden_utils import load
from parity_utils import rel_fro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "csrc", "facade_replay.cpp")
EXE = os.path.join(ROOT, "tests", "lib", "facade_replay")


def build_exe():
    entry.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(entry.LIB)
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE, "-L", libdir, "-leqvio_b200",
           f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True)
    return EXE


def test_facade_compiles_and_links():
    exe = build_exe()
    assert os.path.exists(exe)
    # no arguments: usage exit code, proves the loader resolves libeqvio_b200.so
    assert subprocess.run([exe]).returncode == 2


def _write_stream(path, stream, coord):
    st = stream["settings"]
    v = [coord, st.measurementNoise, st.outlierThresholdAbs, st.outlierThresholdProb, st.featureRetention,
         len(stream["frames"]), len(stream["init"].ids)]
    v += list(stream["init"].sensor.flat()) + list(stream["init"].ids.astype(float)) + list(stream["init"].p.reshape(-1))
    c = stream["cam"]
    v += [c.width, c.height, c.fx, c.fy, c.cx, c.cy]
    for fr in stream["frames"]:
        v += [fr.stamp, len(fr.ids), fr.imu.shape[0]]
        v += list(fr.ids.astype(float)) + list(fr.y.reshape(-1)) + list(fr.provided_p.reshape(-1)) + list(fr.imu.reshape(-1))
    np.array(v, dtype=np.float64).tofile(path)


@pytest.mark.gpu
@pytest.mark.parametrize("name,coord", [("euclid_n16", 0), ("invdepth_n16", 1), ("euclid_n24_gated_noisy", 0)])
def test_cpp_host_reproduces_golden(tmp_path, name, coord):
    exe = build_exe()
    stream, outs = load(name)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_stream(fin, stream, coord)
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rec = np.fromfile(fout, dtype=np.float64)
    c = 0
    for k, o in enumerate(outs):
        N = int(rec[c]); t = rec[c + 1]; c += 2
        sensor = rec[c:c + 23]; c += 23
        ids = rec[c:c + N].astype(np.int64); c += N
        p = rec[c:c + 3 * N].reshape(N, 3); c += 3 * N
        dim = 21 + 3 * N
        Sigma = rec[c:c + dim * dim].reshape(dim, dim).T; c += dim * dim
        assert np.array_equal(ids, o["ids"]), f"update {k}"
        assert t == o["time"]
        assert rel_fro(Sigma, o["Sigma"]) < 1e-9
        assert rel_fro(np.concatenate([sensor, p.reshape(-1)]), np.concatenate([o["sensor"], o["p"].reshape(-1)])) < 1e-9
    assert c == rec.shape[0]
