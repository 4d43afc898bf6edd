"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/eqvio_b200.h declares, its POD defaults are the reference's, and host-only helpers agree
with the oracle.  No compute entry point is exercised here (no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    entry.build()
    from eqvio_b200 import _capi
    return _capi


def _header_functions():
    src = open(os.path.join(ROOT, "include", "eqvio_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eqvio_[a-z_0-9]+)\s*\(", src)))


def test_simulator_header_symbols_exported(capi):
    src = open(os.path.join(ROOT, "include", "eqvio_b200_sim.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(eqvio_sim_[a-z_0-9]+)\s*\(", src)))
    assert names == ["eqvio_sim_create", "eqvio_sim_destroy", "eqvio_sim_imu", "eqvio_sim_imu_instances", "eqvio_sim_last_error",
                     "eqvio_sim_set_noise", "eqvio_sim_vision"]
    for n in names:
        assert hasattr(capi.lib, n), f"{n} declared in include/eqvio_b200_sim.h but not exported"


def test_header_symbols_exported(capi):
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(capi.lib, n), f"{n} declared in include/eqvio_b200.h but not exported"
        assert n in capi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(capi.SIGNATURES) == names


def test_tuning_keys_match_header(capi):
    """The keyword -> key table of VIOFilter.setTuning is the EQVIO_TUNE_* list of the header."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "eqvio_b200.h")).read()
    header = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+EQVIO_TUNE_(\w+)\s+(\d+)", text)}
    assert header, "no EQVIO_TUNE_* definitions found"
    mine = {capi.TUNE_HEADER_NAMES[k]: v for k, v in capi.TUNE.items()}
    assert mine == header


def test_build_info(capi):
    assert b"sm_100a" in capi.lib.eqvio_build_info()


def test_settings_defaults_match_reference(capi):
    from oracle.eqf import Settings as OracleSettings

    s = capi.Settings()
    capi.lib.eqvio_settings_default(C.byref(s))
    o = OracleSettings()
    for name, _ in capi.Settings._fields_:
        if name == "cameraOffset":
            assert list(s.cameraOffset) == [1.0, 0, 0, 0, 0, 0, 0]
        else:
            assert getattr(s, name) == getattr(o, name), name


def test_settings_wrapper_roundtrip(capi):
    import eqvio_b200 as eb
    from oracle.simulator import benchmarkSettings

    o = benchmarkSettings(1, measurementNoise=0.7, featureRetention=0.11)
    s = eb.Settings.fromObject(o)
    assert s.fastRiccati == 1 and s.coordinateChoice == 1
    assert s.measurementNoise == 0.7 and s.featureRetention == 0.11
    with pytest.raises(AttributeError):
        s.noSuchField = 1


def test_inverse_distortion_fit_matches_oracle(capi):
    """StandardCamera::computeInverseDistortion (host-side helper of the ABI) vs the oracle's lstsq."""
    from oracle.camera import StandardCamera

    dist = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0]  # EuRoC cam0
    oc = StandardCamera(752, 480, 458.654, 457.296, 367.215, 248.375, dist)
    cam = capi.Camera()
    cam.model, cam.width, cam.height, cam.ndist = 1, 752, 480, 5
    cam.fx, cam.fy, cam.cx, cam.cy = 458.654, 457.296, 367.215, 248.375
    for i in range(5):
        cam.dist[i] = dist[i]
    assert capi.lib.eqvio_camera_fit_inverse_distortion(C.byref(cam)) == 0
    got = np.array(list(cam.inv_dist))
    # the last monomial (r^6) is nearly collinear with the others on this grid: compare the fitted
    # inverse map rather than raw coefficients
    from oracle.camera import _distort_homogeneous
    pts = np.stack(np.meshgrid(np.linspace(-0.7, 0.7, 9), np.linspace(-0.5, 0.5, 7)), -1).reshape(-1, 2)
    a = _distort_homogeneous(pts, list(got))
    b = _distort_homogeneous(pts, oc.invDist)
    assert np.abs(a - b).max() < 1e-9
    assert np.allclose(got, oc.invDist, rtol=1e-6, atol=1e-9)


def test_null_handles_are_rejected(capi):
    lib = capi.lib
    assert lib.eqvio_process_imu(None, 0.0, None, None, None, None) == capi.EQVIO_ERR_INVALID_ARG
    assert lib.eqvio_num_landmarks(None) == 0
    assert lib.eqvio_get_time(None) == -1.0
    assert lib.eqvio_camera_fit_inverse_distortion(None) == capi.EQVIO_ERR_INVALID_ARG


def test_create_fails_loudly_without_gpu(capi):
    """No CPU fallback: without a CUDA device creation returns EQVIO_ERR_CUDA and a message."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    s = capi.Settings()
    capi.lib.eqvio_settings_default(C.byref(s))
    h = capi._H()
    rc = capi.lib.eqvio_create(C.byref(s), 0, 16, None, C.byref(h))
    assert rc == capi.EQVIO_ERR_CUDA
    assert b"CUDA" in capi.lib.eqvio_last_error(None)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "eqvio_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_library_has_no_blas_or_solver_dependency():
    """Every Settings combination runs on the library's own kernels: no cuBLAS / cuSOLVER name anywhere in the shared object
    (neither a link-time dependency nor a dlopen string), and libcudart is the only CUDA library it needs."""
    import subprocess

    from eqvio_b200 import _capi

    blob = open(_capi.LIB_PATH, "rb").read()
    for name in (b"cublas", b"cusolver", b"cusparse"):
        assert name not in blob.lower(), name
    needed = subprocess.run(["readelf", "-d", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "cublas" not in needed and "cusolver" not in needed
