"""ORACLE import stub (test-only) for pypose, imported by dpvo/loop_closure/optim_utils.py:4 (loop closure, out of scope)."""
