"""Shadow of the reference's ds_mppi/functions/SEDS.py (`from SEDS import *` in the Franka scripts)."""
import _bootstrap  # noqa: F401
import copy  # noqa: F401  (the reference module star-exports these)
import time  # noqa: F401

import numpy as np  # noqa: F401
import torch  # noqa: F401
from scipy.io import loadmat  # noqa: F401

from optimalmodulationds_b200.SEDS import SEDS  # noqa: F401
