"""ORACLE stand-in for the reference's native module `lietorch_backends` (test infrastructure).

Same 19 functions as dpvo/lietorch/src/lietorch.cpp:286-316, computed on the CPU by oracle/lie.py.
Two uses, both test-only:
  * put this directory on sys.path to import the reference's *Python* lietorch package
    (dpvo/lietorch/groups.py, group_ops.py) in a container where the Eigen-based native module
    cannot be built, e.g. to run the reference's own run_tests.py identities against the oracle;
  * monkeypatch it into dpvo_b200.lietorch to exercise the host-side mirror without a GPU.
Only SO3 (1) and SE3 (3) are restated.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import lie as _L  # noqa: E402


def _G(g):
    if g not in _L.GROUPS:
        raise NotImplementedError("oracle lietorch_backends: group %d not restated" % g)
    return _L.GROUPS[g]


def expm(g, a): return _G(g).exp(a)
def expm_backward(g, grad, a): return [_L.expm_backward(g, grad, a)]
def logm(g, X): return _G(g).log(X)
def logm_backward(g, grad, X): return [_L.logm_backward(g, grad, X)]
def inv(g, X): return _G(g).inv(X)
def inv_backward(g, grad, X): return [_L.inv_backward(g, grad, X)]
def mul(g, X, Y): return _G(g).mul(X, Y)
def mul_backward(g, grad, X, Y): return list(_L.mul_backward(g, grad, X, Y))
def adj(g, X, a): return _G(g).adj(X, a)
def adj_backward(g, grad, X, a): return list(_L.adj_backward(g, grad, X, a))
def adjT(g, X, a): return _G(g).adjT(X, a)
def adjT_backward(g, grad, X, a): return list(_L.adjT_backward(g, grad, X, a))
def act(g, X, p): return _G(g).act(X, p)
def act_backward(g, grad, X, p): return list(_L.act_backward(g, grad, X, p))
def act4(g, X, p): return _G(g).act4(X, p)
def act4_backward(g, grad, X, p): return list(_L.act4_backward(g, grad, X, p))
def as_matrix(g, X): return _G(g).matrix(X)
def projector(g, X): return _G(g).projector(X)
def Jinv(g, X, a): return _L.jinv(g, X, a)
