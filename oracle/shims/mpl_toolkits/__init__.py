"""Stand-in for mpl_toolkits (see oracle/shims/matplotlib)."""
from . import mplot3d  # noqa: F401
