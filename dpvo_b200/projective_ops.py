"""Reprojection of patches between frames.  Interface of dpvo/projective_ops.py:19-130 on top of
dpvo_b200.lietorch; `transform_fused` is the single-kernel form for the inference path
(dpvo.py:209-213: pops.transform + permute + contiguous == ~25 launches in the reference)."""
import torch

from . import extensions
from .lietorch import SE3, Sim3

MIN_DEPTH = 0.2


def iproj(patches, intrinsics):
    x, y, d = patches.unbind(dim=2)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    return torch.stack([(x - cx) / fx, (y - cy) / fy, torch.ones_like(d), d], dim=-1)


def proj(X, intrinsics, depth=False):
    X, Y, Z, W = X.unbind(dim=-1)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    d = 1.0 / Z.clamp(min=0.1)
    x = fx * (d * X) + cx
    y = fy * (d * Y) + cy
    return torch.stack([x, y, d] if depth else [x, y], dim=-1)


def transform_fused(poses, patches, intrinsics, ii, jj, kk):
    """[1, E, 2, P, P] coordinates of patch kk (frame ii) in frame jj -- already in the layout corr
    consumes.  poses: SE3 or [1,N,7] tensor.  Inference only (no autograd)."""
    data = poses.data if isinstance(poses, SE3) else poses
    return extensions()[3].reproject_clamped(data, patches, intrinsics, ii, jj, kk)


def transform(poses, patches, intrinsics, ii, jj, kk, depth=False, valid=False, jacobian=False, tonly=False):
    """Differentiable general form (training / keyframing)."""
    X0 = iproj(patches[:, kk], intrinsics[:, ii])
    Gij = poses[:, jj] * poses[:, ii].inv()
    if tonly:
        Gij[..., 3:] = torch.as_tensor([0, 0, 0, 1], device=Gij.device)
    X1 = Gij[:, :, None, None] * X0
    x1 = proj(X1, intrinsics[:, jj], depth)
    if jacobian:
        p = X1.shape[2]
        X, Y, Z, H = X1[..., p // 2, p // 2, :].unbind(dim=-1)
        o = torch.zeros_like(H)
        fx, fy, cx, cy = intrinsics[:, jj].unbind(dim=-1)
        d = torch.zeros_like(Z)
        d[Z.abs() > 0.2] = 1.0 / Z[Z.abs() > 0.2]
        if isinstance(Gij, Sim3):
            raise NotImplementedError("Sim3 poses are not backed by kernels in dpvo_b200")
        Ja = torch.stack([H, o, o, o, Z, -Y, o, H, o, -Z, o, X, o, o, H, Y, -X, o, o, o, o, o, o, o], dim=-1).view(1, len(ii), 4, 6)
        Jp = torch.stack([fx * d, o, -fx * X * d * d, o, o, fy * d, -fy * Y * d * d, o], dim=-1).view(1, len(ii), 2, 4)
        Jj = torch.matmul(Jp, Ja)
        Ji = -Gij[:, :, None].adjT(Jj)
        Jz = torch.matmul(Jp, Gij.matrix()[..., :, 3:])
        return x1, (Z > 0.2).float(), (Ji, Jj, Jz)
    if valid:
        return x1, (X1[..., 2] > 0.2).float()
    return x1


def point_cloud(poses, patches, intrinsics, ix):
    return poses[:, ix, None, None].inv() * iproj(patches, intrinsics[:, ix])


def flow_mag(poses, patches, intrinsics, ii, jj, kk, beta=0.3):
    coords0 = transform(poses, patches, intrinsics, ii, ii, kk)
    coords1, val = transform(poses, patches, intrinsics, ii, jj, kk, tonly=False, valid=True)
    coords2 = transform(poses, patches, intrinsics, ii, jj, kk, tonly=True)
    flow1 = (coords1 - coords0).norm(dim=-1)
    flow2 = (coords2 - coords0).norm(dim=-1)
    return beta * flow1 + (1 - beta) * flow2, (val > 0.5)
