"""Shadow of the reference's ds_mppi/functions/LinDS.py: put this directory FIRST on PYTHONPATH and the
reference's scripts (`sys.path.append('../functions/'); from MPPI import *`) pick up the B200 path."""
import _bootstrap  # noqa: F401
from optimalmodulationds_b200.LinDS import *  # noqa: F401,F403
