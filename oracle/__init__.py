"""ORACLE -- test infrastructure only.

CPU restatements of the reference algorithms on the DPVO update hot path, used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker.
Nothing under dpvo_b200/ imports this package.
"""
