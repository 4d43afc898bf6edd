"""Loads eqvio_b200/replicas.py on its own (the package __init__ needs the built CUDA library; this module is
pure host logic and must be testable in spawned CPU workers regardless)."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eqvio_b200", "replicas.py")
_spec = importlib.util.spec_from_file_location("eqvio_b200_replicas", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
gather_trajectories = _m.gather_trajectories
shard_instances = _m.shard_instances
trajectory_row = _m.trajectory_row
