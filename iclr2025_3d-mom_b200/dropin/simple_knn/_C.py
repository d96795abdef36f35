"""Stand-in for the compiled module simple_knn._C (KNN/ext.cpp)."""
import os
import sys

_pkg_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _pkg_root not in sys.path:
    sys.path.insert(0, _pkg_root)

from b200gs.knn import distCUDA2  # noqa: E402,F401
