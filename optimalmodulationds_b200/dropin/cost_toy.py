"""Shadow of the reference's ds_mppi/functions/cost_toy.py: put this directory FIRST on PYTHONPATH and the
reference's scripts pick up the B200 path."""
import _bootstrap  # noqa: F401
from optimalmodulationds_b200.cost_toy import *  # noqa: F401,F403
