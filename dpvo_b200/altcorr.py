"""altcorr host interface over cuda_corr: same call surface as the reference's
dpvo/altcorr/correlation.py:4-74 (`corr`, `patchify`) plus the fused two-level form DPVO.corr needs.

Autograd: CorrFn / PatchFn call cuda_corr.backward / patchify_backward.  As in the reference
(correlation.py:20-25) the correlation backward keeps a random `dropout` fraction of the edges and
draws that mask with torch.rand on the CUDA generator *inside* backward, so RNG consumption stays
aligned with the reference; coords receive no gradient (correlation.py:30)."""
import torch

from . import extensions


class CorrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap1, fmap2, coords, ii, jj, radius, dropout):
        ctx.save_for_backward(fmap1, fmap2, coords, ii, jj)
        ctx.radius, ctx.dropout = radius, dropout
        out, = extensions()[0].forward(fmap1, fmap2, coords, ii, jj, radius)
        return out

    @staticmethod
    def backward(ctx, grad):
        fmap1, fmap2, coords, ii, jj = ctx.saved_tensors
        if ctx.dropout < 1:
            keep = torch.rand(len(ii), device=grad.device) < ctx.dropout
            coords, grad, ii, jj = coords[:, keep], grad[:, keep], ii[keep], jj[keep]
        g1, g2 = extensions()[0].backward(fmap1, fmap2, coords, ii, jj, grad.float(), ctx.radius)
        return g1, g2, None, None, None, None, None


class PatchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, coords, radius):
        ctx.radius = radius
        ctx.save_for_backward(net, coords)
        out, = extensions()[0].patchify_forward(net, coords, radius)
        return out

    @staticmethod
    def backward(ctx, grad):
        net, coords = ctx.saved_tensors
        g, = extensions()[0].patchify_backward(net, coords, grad, ctx.radius)
        return g, None, None


def corr(fmap1, fmap2, coords, ii, jj, radius=1, dropout=1):
    """[B, M, 2R+1 (x), 2R+1 (y), P, P] local correlation volume."""
    return CorrFn.apply(fmap1, fmap2, coords, ii, jj, radius, dropout)


def corr_pyramid(fmap1, pyramid, coords, ii, jj, radius=3, lvl1_div=4.0, pad_to=0, out=None):
    """DPVO.corr (dpvo.py:200-207) in one launch: pyramid = (level0, level1); returns [B, M, 882]
    for radius 3, P 3 (feature order x-off, y-off, pi, pj, level).  pad_to=896 returns rows padded
    with zeros to the k-block of the first dense layer of the update operator.  Inference only."""
    out = extensions()[3].corr_pyramid2(fmap1, pyramid[0], pyramid[1], coords, ii, jj, radius, lvl1_div, pad_to, out)
    return out.view(out.shape[0], out.shape[1], -1)


def patchify(net, coords, radius, mode="bilinear"):
    """(2R+1)^2 window of `net` around `coords` (bilinear) or the raw (2R+2)^2 window (mode=None)."""
    patches = PatchFn.apply(net, coords, radius)
    if mode != "bilinear":
        return patches
    frac = coords - coords.floor()
    dx, dy = frac[:, :, None, None, None].unbind(dim=-1)
    d = 2 * radius + 1
    return ((1 - dy) * (1 - dx) * patches[..., :d, :d] + (1 - dy) * dx * patches[..., :d, 1:] +
            dy * (1 - dx) * patches[..., 1:, :d] + dy * dx * patches[..., 1:, 1:])
