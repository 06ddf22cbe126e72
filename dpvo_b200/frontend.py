"""Front end of the path: feature / context encoders and patch extraction -- the role of dpvo/net.py:95-157
(`Patchifier`) and dpvo/extractor.py:200-264 (`BasicEncoder4`).  SURVEY 8(f) rank 3: the step before the hot
path; kept on library convolutions (cuDNN through torch -- plain library ops are allowed for non-hot-path dense
work, the hand-written kernels are the patch gathers of cuda_corr.patchify) and written so that the reference's
checkpoint keys ("patchify.fnet.layer1.0.conv1.weight", ...) load unchanged.

The encoder is a 7x7/2 stem, two residual stages (the second strided: 1/4 resolution in total) and a 1x1 output
convolution.  fnet uses instance normalisation, inet none.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import altcorr
from .net import DIM as CTX_DIM

ENC_DIM = 32


def _norm(kind, ch):
    return nn.InstanceNorm2d(ch) if kind == "instance" else nn.Sequential()


class _Res(nn.Module):
    """two 3x3 convolutions with a (possibly strided, 1x1-projected) skip; attribute names follow the checkpoint"""

    def __init__(self, cin, cout, norm, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.norm1, self.norm2 = _norm(norm, cout), _norm(norm, cout)
        self.downsample = None
        if stride != 1:
            self.norm3 = _norm(norm, cout)
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride), self.norm3)

    def forward(self, x):
        y = F.relu(self.norm1(self.conv1(x)))
        y = F.relu(self.norm2(self.conv2(y)))
        return F.relu((x if self.downsample is None else self.downsample(x)) + y)


class Encoder4(nn.Module):
    """[b,n,3,H,W] -> [b,n,out,H/4,W/4]"""

    def __init__(self, output_dim=128, norm_fn="instance"):
        super().__init__()
        self.conv1 = nn.Conv2d(3, ENC_DIM, 7, stride=2, padding=3)
        self.norm1 = _norm(norm_fn, ENC_DIM)
        self.layer1 = nn.Sequential(_Res(ENC_DIM, ENC_DIM, norm_fn, 1), _Res(ENC_DIM, ENC_DIM, norm_fn, 1))
        self.layer2 = nn.Sequential(_Res(ENC_DIM, 2 * ENC_DIM, norm_fn, 2), _Res(2 * ENC_DIM, 2 * ENC_DIM, norm_fn, 1))
        self.conv2 = nn.Conv2d(2 * ENC_DIM, output_dim, 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        b, n = x.shape[:2]
        x = x.flatten(0, 1)
        x = F.relu(self.norm1(self.conv1(x)))
        x = self.conv2(self.layer2(self.layer1(x)))
        return x.view(b, n, *x.shape[1:])


def pixel_grid_with_depth(disps):
    """[b,n,h,w] inverse depths -> [b,n,3,h,w] (x, y, d) per pixel"""
    b, n, h, w = disps.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float, device=disps.device),
                            torch.arange(w, dtype=torch.float, device=disps.device), indexing="ij")
    return torch.stack([xs.expand(b, n, h, w), ys.expand(b, n, h, w), disps], dim=2)


class Patchifier(nn.Module):
    def __init__(self, patch_size=3):
        super().__init__()
        self.patch_size = patch_size
        self.fnet = Encoder4(128, "instance")
        self.inet = Encoder4(CTX_DIM, "none")

    def forward(self, images, patches_per_image=80, disps=None, centroid_sel_strat="RANDOM", return_color=False):
        """features, patch features, context, patches [b, n*M, 3, P, P] and their frame index (net.py:110-157).
        RNG: torch.randint x then y on the CUDA generator, in the reference's order."""
        fmap = self.fnet(images) / 4.0
        imap = self.inet(images) / 4.0
        b, n, _, h, w = fmap.shape
        P = self.patch_size
        dev = fmap.device
        if centroid_sel_strat == "GRADIENT_BIAS":
            gray = ((images + 0.5) * (255.0 / 2)).sum(dim=2)
            dx = gray[..., :-1, 1:] - gray[..., :-1, :-1]
            dy = gray[..., 1:, :-1] - gray[..., :-1, :-1]
            g = F.avg_pool2d(torch.sqrt(dx ** 2 + dy ** 2), 4, 4)
            x = torch.randint(1, w - 1, size=[n, 3 * patches_per_image], device=dev)
            y = torch.randint(1, h - 1, size=[n, 3 * patches_per_image], device=dev)
            score = altcorr.patchify(g[0, :, None], torch.stack([x, y], -1).float(), 0).view(n, 3 * patches_per_image)
            keep = torch.argsort(score, dim=1)[:, -patches_per_image:]
            x, y = torch.gather(x, 1, keep), torch.gather(y, 1, keep)
        elif centroid_sel_strat == "RANDOM":
            x = torch.randint(1, w - 1, size=[n, patches_per_image], device=dev)
            y = torch.randint(1, h - 1, size=[n, patches_per_image], device=dev)
        else:
            raise NotImplementedError("Patch centroid selection not implemented: %s" % centroid_sel_strat)
        coords = torch.stack([x, y], dim=-1).float()
        imap = altcorr.patchify(imap[0], coords, 0).view(b, -1, CTX_DIM, 1, 1)
        gmap = altcorr.patchify(fmap[0], coords, P // 2).view(b, -1, 128, P, P)
        if disps is None:
            disps = torch.ones(b, n, h, w, device=dev)
        patches = altcorr.patchify(pixel_grid_with_depth(disps)[0], coords, P // 2).view(b, -1, 3, P, P)
        index = torch.arange(n, device=dev).repeat_interleave(patches_per_image)
        if return_color:
            clr = altcorr.patchify(images[0], 4 * (coords + 0.5), 0).view(b, -1, 3)
            return fmap, gmap, imap, patches, index, clr
        return fmap, gmap, imap, patches, index
