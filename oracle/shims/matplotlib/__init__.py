"""ORACLE import stub (test-only): the reference imports matplotlib at module import time (dpvo/net.py:23)."""
