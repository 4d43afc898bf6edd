"""ctypes binding of the C ABI in include/eqvio_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) at
``eqvio_b200/lib/libeqvio_b200.so``.  There is no CPU fallback: if the library is missing
the import of this module raises, and every compute entry point returns EQVIO_ERR_CUDA when
no device is present.
"""
import ctypes as C
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libeqvio_b200.so")
LIB_PATH = os.environ.get("EQVIO_B200_LIB", LIB_PATH)  # debug builds (-DEQVIO_TIMELINE ...); still no CPU fallback

EQVIO_OK = 0
# eqvio_set_tuning keys (include/eqvio_b200.h: EQVIO_TUNE_*), keyed by the keyword VIOFilter.setTuning takes
TUNE = dict(correction=0, chunkLandmarks=1, speculate=2, graph=3, downdate=5, lookahead=6, fuseObserver=7, pdl=8,
            fuseSmall=10, speculateNew=11, stageS=12, zeroCopy=13, propFusion=14, lazyDowndate=15)
TUNE_HEADER_NAMES = dict(correction="CORRECTION", chunkLandmarks="CHUNK_LANDMARKS", speculate="SPECULATE", graph="GRAPH",
                         downdate="DOWNDATE", lookahead="LOOKAHEAD", fuseObserver="FUSE_OBSERVER", pdl="PDL",
                         fuseSmall="FUSE_SMALL", speculateNew="SPECULATE_NEW", stageS="STAGE_S", zeroCopy="ZERO_COPY",
                         propFusion="PROP_FUSION", lazyDowndate="LAZY_DOWNDATE")
EQVIO_ERR_INVALID_ARG = -1
EQVIO_ERR_CUDA = -2
EQVIO_ERR_NUMERIC = -3
EQVIO_ERR_CAPACITY = -4
EQVIO_ERR_UNSUPPORTED = -5
ERROR_NAMES = {
    EQVIO_ERR_INVALID_ARG: "EQVIO_ERR_INVALID_ARG",
    EQVIO_ERR_CUDA: "EQVIO_ERR_CUDA",
    EQVIO_ERR_NUMERIC: "EQVIO_ERR_NUMERIC",
    EQVIO_ERR_CAPACITY: "EQVIO_ERR_CAPACITY",
    EQVIO_ERR_UNSUPPORTED: "EQVIO_ERR_UNSUPPORTED",
}
PROF_CLASSES = 7
PROF_NAMES = ("prop_ll", "chunk_factor", "chol_trail", "downdate", "bc_diag", "bc_panel", "bc_trail")

_D = C.c_double
_I = C.c_int


class Settings(C.Structure):
    """eqvio_settings -- POD mirror of VIOFilter::Settings (VIOFilterSettings.h:58-99)."""

    _fields_ = [(n, _D) for n in (
        "biasOmegaProcessVariance", "biasAccelProcessVariance", "attitudeProcessVariance", "positionProcessVariance",
        "velocityProcessVariance", "cameraAttitudeProcessVariance", "cameraPositionProcessVariance",
        "pointProcessVariance", "velGyrNoise", "velAccNoise", "velGyrBiasWalk", "velAccBiasWalk", "measurementNoise",
        "outlierThresholdAbs", "outlierThresholdProb", "featureRetention", "initialAttitudeVariance",
        "initialPositionVariance", "initialVelocityVariance", "initialCameraAttitudeVariance",
        "initialCameraPositionVariance", "initialPointVariance", "initialPointDepthVariance",
        "initialBiasOmegaVariance", "initialBiasAccelVariance", "initialSceneDepth")] + [(n, _I) for n in (
            "useDiscreteInnovationLift", "useDiscreteVelocityLift", "useDiscreteStateMatrix", "fastRiccati",
            "useMedianDepth", "useFeaturePredictions", "useEquivariantOutput", "removeLostLandmarks",
            "coordinateChoice")] + [("cameraOffset", _D * 7)]


class Camera(C.Structure):
    """eqvio_camera -- flattened GIFT::GICamera."""

    _fields_ = [("model", _I), ("width", _I), ("height", _I), ("ndist", _I), ("fx", _D), ("fy", _D), ("cx", _D),
                ("cy", _D), ("dist", _D * 5), ("inv_dist", _D * 5)]


class FilterHandle(C.Structure):
    pass


_H = C.POINTER(FilterHandle)
_PD = C.POINTER(_D)
_PI = C.POINTER(_I)

# name -> (restype, argtypes); kept in step with include/eqvio_b200.h (tests/test_capi.py checks it)
SIGNATURES = {
    "eqvio_settings_default": (None, [C.POINTER(Settings)]),
    "eqvio_camera_fit_inverse_distortion": (_I, [C.POINTER(Camera)]),
    "eqvio_create": (_I, [C.POINTER(Settings), _I, _I, C.c_void_p, C.POINTER(_H)]),
    "eqvio_create_from_state": (_I, [C.POINTER(Settings), _I, _I, C.c_void_p, _PD, _I, _PI, _PD, _D, C.POINTER(_H)]),
    "eqvio_destroy": (None, [_H]),
    "eqvio_last_error": (C.c_char_p, [_H]),
    "eqvio_initialise_from_imu": (_I, [_H, _D, _PD, _PD]),
    "eqvio_set_state": (_I, [_H, _PD, _I, _PI, _PD]),
    "eqvio_set_landmarks": (_I, [_H, _I, _PI, _PD]),
    "eqvio_augment_landmark_states": (_I, [_H, _I, _PI, _I, _PI, _PD]),
    "eqvio_process_imu": (_I, [_H, _D, _PD, _PD, _PD, _PD]),
    "eqvio_process_imu_rows": (_I, [_H, _I, _PD]),
    "eqvio_process_vision": (_I, [_H, _D, _I, _PI, _PD, C.POINTER(Camera), _PI]),
    "eqvio_batch_process_vision": (_I, [C.POINTER(_H), _I, _PD, _PI, C.POINTER(_PI), C.POINTER(_PD),
                                        C.POINTER(Camera), _PI]),
    "eqvio_get_time": (_D, [_H]),
    "eqvio_is_initialised": (_I, [_H]),
    "eqvio_num_landmarks": (_I, [_H]),
    "eqvio_state_dim": (_I, [_H]),
    "eqvio_capacity": (_I, [_H]),
    "eqvio_get_state_estimate": (_I, [_H, _PD, _PI, _PD, _PI]),
    "eqvio_get_eqf_state": (_I, [_H, _PD, _PI, _PD, _PD, _PD, _PD, _I]),
    "eqvio_get_landmark_cov_blocks": (_I, [_H, _PD]),
    "eqvio_get_feature_predictions": (_I, [_H, C.POINTER(Camera), _D, _PI, _PD, _PI]),
    "eqvio_compute_nees": (_I, [_H, _PD, _I, _PI, _PD, _PD]),
    "eqvio_get_last_outliers": (_I, [_H, _PI, _I, _PI]),
    "eqvio_get_stage_ms": (_I, [_H, _PD]),
    "eqvio_enable_stage_timing": (_I, [_H, _I]),
    "eqvio_get_launch_count": (C.c_longlong, [_H]),
    "eqvio_get_graph_stats": (_I, [_H, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "eqvio_enable_kernel_profile": (_I, [_H, _I]),
    "eqvio_get_kernel_profile": (_I, [_H, _I, _PD, C.POINTER(C.c_longlong)]),
    "eqvio_replay_batch": (_I, [C.c_void_p, _I, _I, C.c_void_p, C.c_void_p, _PD, _PD, _PD]),
    "eqvio_replay": (_I, [_H, _I, C.c_void_p, C.c_void_p, C.c_size_t, _PD, _PD]),
    "eqvio_get_host_profile": (_I, [_H, _I, _PD, C.POINTER(C.c_longlong)]),
    "eqvio_set_tuning": (_I, [_H, _I, _I]),
    "eqvio_plan_lazy_downdates": (_I, [_I, _I, _PI, _PI, _I, _PI, _I]),
    "eqvio_build_info": (C.c_char_p, []),
}


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "eqvio_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class EqvioError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code
