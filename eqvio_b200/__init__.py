"""eqvio_b200 -- B200-native EqF vision-update path of EqVIO behind the VIOFilter surface.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
``include/eqvio_b200.h``), ``lib/`` (the built shared library) and ``filter.py`` (host-side mirror
of the reference's ``VIOFilter`` over that ABI).  Importing the package loads the shared library
and raises ImportError when it has not been built -- there is no CPU fallback.
"""
from ._capi import EqvioError, LIB_PATH  # noqa: F401
from .filter import (  # noqa: F401
    COORD_EUCLIDEAN, COORD_INVDEPTH, COORD_NORMAL, Camera, EqFState, IMUVelocity, Settings, VIOFilter, VIOSensorState,
    VIOState, VisionMeasurement, batchProcessVision, replayBatch)
from .writer import VIOWriter, trajectory_errors  # noqa: F401
from .stream import FeatureStream, run_stream, settings_from_yaml  # noqa: F401


def build_info():
    from ._capi import lib
    return lib.eqvio_build_info().decode()
