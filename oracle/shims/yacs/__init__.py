"""ORACLE import stub (test-only) for yacs (dpvo/config.py:1)."""
