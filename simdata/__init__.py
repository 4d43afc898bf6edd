"""Synthetic input streams for the benchmark and the tests: a self-contained numpy restatement of
the reference's VIOSimulator (src/VIOSimulator.cpp, src/dataserver/SimulationDataServer.cpp).
It produces plain arrays and imports neither the product package nor the oracle."""
from .vio_simulator import Frame, SimConfig, SimStream, record_stream  # noqa: F401
