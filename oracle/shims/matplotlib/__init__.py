"""Headless stand-in for matplotlib (not installed in this image) so the UNMODIFIED reference can be imported and
its scripts run without a display: ds_mppi/functions/fk_num.py:4 star-imports plots.py:1, which imports pyplot.
Test / baseline infrastructure only (oracle/ref_harness.py puts oracle/shims on sys.path); it draws nothing and
touches no arithmetic."""
from . import pyplot  # noqa: F401
