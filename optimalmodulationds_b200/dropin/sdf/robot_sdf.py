"""Shadow of the reference's mlp_learn/sdf/robot_sdf.py (see ../MPPI.py)."""
import _bootstrap  # noqa: F401
from optimalmodulationds_b200.sdf.robot_sdf import *  # noqa: F401,F403
from optimalmodulationds_b200.sdf.robot_sdf import RobotSdfCollisionNet  # noqa: F401
