"""The reference-import harness lives in oracle/ref_harness.py (shared with bench.py's CPU legs); this module keeps
the name the golden generators import."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_harness import *  # noqa: E402,F401,F403
from oracle.ref_harness import REF, REF_FUNCS, REF_MLP, REF_MODELS, available, load_reference, make_ref_net  # noqa: E402,F401
