"""ORACLE import stub (test-only) for pypose, imported by dpvo/loop_closure/optim_utils.py:4 (loop closure, out of scope)."""


class _Unavailable:
    """names used only in type annotations at import time (optim_utils.py:15); calling them is an error"""

    def __init__(self, *a, **k):
        raise NotImplementedError("pypose is not installed (loop closure is out of scope)")


SE3 = Sim3 = _Unavailable
