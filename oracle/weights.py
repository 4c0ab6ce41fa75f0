"""Seeded parameters, model configurations and synthetic backbones: re-exported from the product package's workload
definition (``rmpnn_b200/synthetic.py``, pure data generation) so that tests and golden generators keep one import."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rna-mpnn_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from rmpnn_b200.synthetic import *  # noqa: F401,F403,E402
from rmpnn_b200.synthetic import (NODE_IN, EDGE_IN, RDesignConfig, RNAMPNNConfig, make_rnampnn_state_dict,  # noqa: F401,E402
                                  make_state_dict, synthetic_backbones)
