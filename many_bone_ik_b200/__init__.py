"""many_bone_ik_b200 -- B200-native batched solver for ManyBoneIK's iterative constrained IK solve loop.

Product = libmbik.so (hand-written sm_100a CUDA + C++ host runtime) behind the C ABI in include/mbik.h.
This package holds the sources (csrc/), the ctypes binding (solver.py, _capi.py) and the benchmark rig
generators (rigs.py).  There is no CPU or Python fallback for the solve.
"""
from . import rigs  # noqa: F401
from .solver import BatchedIKRig, IKStream, MbikError, device_count  # noqa: F401
