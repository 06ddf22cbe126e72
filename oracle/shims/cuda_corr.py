"""ORACLE stand-in for the reference's native `cuda_corr` (CPU, from oracle/corr.py).  Test-only:
lets dpvo/altcorr/correlation.py import and run in the build container."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import corr as _C  # noqa: E402


def forward(fmap1, fmap2, coords, ii, jj, radius):
    return [_C.corr_forward(fmap1, fmap2, coords, ii, jj, radius)]


def backward(fmap1, fmap2, coords, ii, jj, grad, radius):
    f1 = fmap1.detach().clone().requires_grad_(True)
    f2 = fmap2.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        out = _C.corr_forward(f1, f2, coords, ii, jj, radius)
        g1, g2 = torch.autograd.grad(out, [f1, f2], grad.to(out.dtype))
    return [g1, g2]


def patchify_forward(net, coords, radius):
    return [_C.patchify_raw(net, coords, radius)]


def patchify_backward(net, coords, gradient, radius):
    n = net.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        out = _C.patchify_raw(n, coords, radius)
        g, = torch.autograd.grad(out, [n], gradient.to(out.dtype))
    return [g]
