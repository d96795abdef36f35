rcParams = {}
from . import pyplot  # noqa: E402,F401
