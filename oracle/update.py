"""ORACLE (test infrastructure -- never imported by the product path).

CPU restatement in plain PyTorch of the reference's update operator:
    Update            dpvo/net.py:27-92   (same sub-module names => same state_dict keys)
    GatedResidual     dpvo/blocks.py:15-29
    SoftAgg           dpvo/blocks.py:31-48
    GradientClip      dpvo/blocks.py:74-89
torch_scatter (pytorch-scatter 2.1.2, environment.yml:12) is absent from /root/reference; its two
functions used here are restated from their published definition:
    scatter_softmax(x, idx, dim=1): per group, exp(x - max_group) / sum_group exp(x - max_group)
    scatter_sum(x, idx, dim=1):     index_add
fastba.neighbors is taken from oracle/graph.py.

Pinning: the reference has no test or fixture for this module ("parity unpinned" for the
torch_scatter part).  tests/golden/update_*.pt are generated in the build container by importing the
reference's own dpvo/net.py:Update (oracle/make_golden_update.py) with these scatter restatements
injected as the `torch_scatter` module, and the restatement must reproduce them.
"""
import torch
import torch.nn as nn

from . import graph

DIM = 384


def scatter_sum(src, index, dim=1, dim_size=None):
    assert dim == 1
    n = int(index.max().item()) + 1 if dim_size is None else dim_size
    out = torch.zeros(src.shape[:1] + (n,) + src.shape[2:], dtype=src.dtype, device=src.device)
    return out.index_add_(1, index, src)


def scatter_max(src, index, dim=1, dim_size=None):
    assert dim == 1
    n = int(index.max().item()) + 1 if dim_size is None else dim_size
    out = torch.full(src.shape[:1] + (n,) + src.shape[2:], float("-inf"), dtype=src.dtype, device=src.device)
    idx = index.view(1, -1, *([1] * (src.dim() - 2))).expand_as(src)
    return out.scatter_reduce(1, idx, src, reduce="amax", include_self=True), None


def scatter_softmax(src, index, dim=1):
    assert dim == 1
    mx, _ = scatter_max(src, index, dim=1)
    ex = torch.exp(src - mx[:, index])
    den = scatter_sum(ex, index, dim=1)
    return ex / den[:, index]


class GradClip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        g = torch.where(torch.isnan(g), torch.zeros_like(g), g)
        return g.clamp(min=-0.01, max=0.01)


class GradientClip(nn.Module):
    def forward(self, x):
        return GradClip.apply(x)


class GatedResidual(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gate = nn.Sequential(nn.Linear(dim, dim), nn.Sigmoid())
        self.res = nn.Sequential(nn.Linear(dim, dim), nn.ReLU(inplace=True), nn.Linear(dim, dim))

    def forward(self, x):
        return x + self.gate(x) * self.res(x)


class SoftAgg(nn.Module):
    def __init__(self, dim=512, expand=True):
        super().__init__()
        self.dim, self.expand = dim, expand
        self.f = nn.Linear(dim, dim)
        self.g = nn.Linear(dim, dim)
        self.h = nn.Linear(dim, dim)

    def forward(self, x, ix):
        _, jx = torch.unique(ix, return_inverse=True)
        w = scatter_softmax(self.g(x), jx, dim=1)
        y = scatter_sum(self.f(x) * w, jx, dim=1)
        if self.expand:
            return self.h(y)[:, jx]
        return self.h(y)


class Update(nn.Module):
    def __init__(self, p=3):
        super().__init__()
        self.c1 = nn.Sequential(nn.Linear(DIM, DIM), nn.ReLU(inplace=True), nn.Linear(DIM, DIM))
        self.c2 = nn.Sequential(nn.Linear(DIM, DIM), nn.ReLU(inplace=True), nn.Linear(DIM, DIM))
        self.norm = nn.LayerNorm(DIM, eps=1e-3)
        self.agg_kk = SoftAgg(DIM)
        self.agg_ij = SoftAgg(DIM)
        self.gru = nn.Sequential(nn.LayerNorm(DIM, eps=1e-3), GatedResidual(DIM),
                                 nn.LayerNorm(DIM, eps=1e-3), GatedResidual(DIM))
        self.corr = nn.Sequential(nn.Linear(2 * 49 * p * p, DIM), nn.ReLU(inplace=True), nn.Linear(DIM, DIM),
                                  nn.LayerNorm(DIM, eps=1e-3), nn.ReLU(inplace=True), nn.Linear(DIM, DIM))
        self.d = nn.Sequential(nn.ReLU(inplace=False), nn.Linear(DIM, 2), GradientClip())
        self.w = nn.Sequential(nn.ReLU(inplace=False), nn.Linear(DIM, 2), GradientClip(), nn.Sigmoid())

    def forward(self, net, inp, corr, flow, ii, jj, kk):
        net = net + inp + self.corr(corr)
        net = self.norm(net)
        ix, jx = graph.neighbors(kk, jj)
        ix, jx = ix.to(net.device), jx.to(net.device)
        mask_ix = (ix >= 0).float().reshape(1, -1, 1)
        mask_jx = (jx >= 0).float().reshape(1, -1, 1)
        net = net + self.c1(mask_ix * net[:, ix])
        net = net + self.c2(mask_jx * net[:, jx])
        net = net + self.agg_kk(net, kk)
        net = net + self.agg_ij(net, ii * 12345 + jj)
        net = self.gru(net)
        return net, (self.d(net), self.w(net), None)
