"""starflood_b200 - B200-native force-evaluation engine for Starflood (PP / Barnes-Hut / SPH hot path).

csrc/   hand-written sm_100a CUDA kernels + the C ABI (include/sfengine.h) -> libsfengine.so
shim/   C shims that slot into the reference build in place of gravity.c / tree.c / sph.c
engine  ctypes host binding mirroring the reference's solver interface (tests, bench)
"""
from . import engine  # noqa: F401
