"""Shadow of the reference's ds_mppi/functions/MPPI_toy.py: put this directory FIRST on PYTHONPATH and the
reference's scripts (`sys.path.append('../functions/'); from MPPI_toy import *`) pick up the B200 path."""
import _bootstrap  # noqa: F401
from optimalmodulationds_b200.MPPI_toy import *  # noqa: F401,F403
