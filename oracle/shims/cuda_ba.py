"""ORACLE stand-in for the reference's native `cuda_ba` (CPU, from oracle/ba.py, oracle/graph.py).
Test-only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ba as _B, graph as _G  # noqa: E402


def forward(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, PPF, t0, t1, iterations, eff_impl):
    P = patches.shape[-1]
    p, q = _B.fastba_forward(poses.view(-1, 7), patches.view(-1, 3, P, P), intrinsics.view(-1, 4), target.view(-1, 2),
                             weight.view(-1, 2), lmbda, ii, jj, kk, t0, t1, iterations)
    poses.view(-1, 7).copy_(p)
    patches.view(-1, 3, P, P).copy_(q)
    return []


def neighbors(ii, jj):
    ix, jx = _G.neighbors(ii, jj)
    return [ix.to(ii.device), jx.to(ii.device)]


def reproject(poses, patches, intrinsics, ii, jj, kk):
    P = patches.shape[-1]
    return _B.fastba_reproject(poses.view(-1, 7), patches.view(-1, 3, P, P), intrinsics.view(-1, 4), ii, jj, kk)[None]


def solve_system(*a, **k):
    raise NotImplementedError("cuda_ba.solve_system has no oracle (Eigen sparse solve, loop closure only)")
