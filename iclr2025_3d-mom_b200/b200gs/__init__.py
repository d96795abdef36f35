"""b200gs — B200-native hot path of the 3D-MOM 4D Gaussian-splatting trainer/renderer.

Host-side mirror of the reference's operator interface for this path; the arithmetic lives
in libb200gs.so (hand-written sm_100a CUDA behind the C ABI in include/b200gs.h).
"""
from ._lib import B200GSError, LIB_PATH, lib  # noqa: F401
