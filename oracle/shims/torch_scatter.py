"""ORACLE stand-in for `torch_scatter` (pytorch-scatter 2.1.2, absent from the image and from
/root/reference).  Test infrastructure: lets the reference's Python files import in the build
container.  Semantics restated in oracle/update.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.update import scatter_sum, scatter_softmax, scatter_max  # noqa: E402,F401
