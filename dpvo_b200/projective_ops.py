"""Reprojection of patches between frames -- the call surface of dpvo/projective_ops.py:19-130 (`iproj`,
`proj`, `transform`, `point_cloud`, `flow_mag`) over dpvo_b200.lietorch, plus `transform_fused`, the
single-kernel form the inference loop uses (dpvo.py:209-213: pops.transform + permute + contiguous is ~25
launches in the reference).

Written from the camera model, not from the reference file:
  * a patch pixel (x, y, d) back-projects to the homogeneous point ((x-cx)/fx, (y-cy)/fy, 1, d);
  * frame-i points move to frame j by G_ij = G_j * G_i^-1 acting on homogeneous points (lietorch act4);
  * pixel = f * X / max(Z, 0.1) + c.
The Jacobians of the centre pixel are emitted in closed form -- each row of d(pixel)/d(xi_j) is written out
directly (two rows of six / seven entries) instead of multiplying a 2x4 by a 4x6 matrix per edge, and the
depth column uses that the fourth column of a rigid / similarity matrix is (t, 1) -- and `flow_mag` uses
that reprojecting a patch into its own frame returns its own pixel grid, which removes one of the three
transforms of projective_ops.py:120-130.  Same results (tests/test_projective_gpu.py vs the oracle that is
pinned bit-exactly to the reference file), less work.
"""
import torch

from . import extensions
from .lietorch import SE3, Sim3

MIN_DEPTH = 0.2        # validity / Jacobian gate on Z (projective_ops.py:79-80, 110-113)
Z_CLAMP = 0.1          # projection clamp (projective_ops.py:43)


def _cam(intrinsics, ndim):
    """(fx, fy, cx, cy), each shaped to broadcast over `ndim` - 2 trailing patch dims"""
    k = intrinsics.reshape(intrinsics.shape[:2] + (1,) * (ndim - 2) + (4,))
    return k[..., 0], k[..., 1], k[..., 2], k[..., 3]


def iproj(patches, intrinsics):
    """[B,E,3,P,P] pixels + inverse depth -> [B,E,P,P,4] homogeneous points with inverse depth as 4th coordinate"""
    fx, fy, cx, cy = _cam(intrinsics, 4)
    u, v, d = patches[:, :, 0], patches[:, :, 1], patches[:, :, 2]
    return torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(d), d], dim=-1)


def proj(X, intrinsics, depth=False):
    """[B,E,P,P,4] points -> [B,E,P,P,2] pixels (with the inverse range as a third channel if `depth`)"""
    fx, fy, cx, cy = _cam(intrinsics, 4)
    inv_z = 1.0 / X[..., 2].clamp(min=Z_CLAMP)
    out = [fx * (X[..., 0] * inv_z) + cx, fy * (X[..., 1] * inv_z) + cy]
    if depth:
        out.append(inv_z)
    return torch.stack(out, dim=-1)


def transform_fused(poses, patches, intrinsics, ii, jj, kk):
    """[1, E, 2, P, P] coordinates of patch kk (frame ii) in frame jj -- already in the layout corr
    consumes.  poses: SE3 or [1,N,7] tensor.  Inference only (no autograd)."""
    data = poses.data if isinstance(poses, SE3) else poses
    return extensions()[3].reproject_clamped(data, patches, intrinsics, ii, jj, kk)


def _centre_jacobians(Gij, Xc, intrinsics_j):
    """d(centre pixel)/d(xi_i), d(.)/d(xi_j) and d(.)/d(inverse depth) for every edge.
    Xc [B,E,4] is the centre point in frame j.  Left perturbations, translation first (lietorch order)."""
    fx, fy = intrinsics_j[..., 0], intrinsics_j[..., 1]
    X, Y, Z, H = Xc.unbind(dim=-1)
    zero = torch.zeros_like(Z)
    d = torch.where(Z.abs() > MIN_DEPTH, 1.0 / Z, zero)
    ax, ay = fx * d, fy * d                       # d(px)/dX, d(py)/dY
    bx, by = -ax * X * d, -ay * Y * d             # d(px)/dZ, d(py)/dZ
    # a point moves by (tau*H + phi x p [+ s*p]); chain through the two projection rows
    row_x = [ax * H, zero, bx * H, bx * Y, ax * Z - bx * X, -ax * Y]
    row_y = [zero, ay * H, by * H, -ay * Z + by * Y, -by * X, ay * X]
    if isinstance(Gij, Sim3):                     # scale generator moves the point along itself
        row_x.append(ax * X + bx * Z)
        row_y.append(ay * Y + by * Z)
    Jj = torch.stack([torch.stack(row_x, -1), torch.stack(row_y, -1)], dim=-2)          # [B,E,2,6|7]
    Ji = -Gij[:, :, None].adjT(Jj)
    # inverse depth is the homogeneous coordinate: the point moves along the 4th column of the matrix = (t, 1)
    t = Gij.translation()
    Jz = torch.stack([ax * t[..., 0] + bx * t[..., 2], ay * t[..., 1] + by * t[..., 2]], dim=-1)[..., None]
    return Ji, Jj, Jz


def transform(poses, patches, intrinsics, ii, jj, kk, depth=False, valid=False, jacobian=False, tonly=False):
    """Differentiable general form (training / keyframing): pixels of patch kk (living in frame ii) seen from
    frame jj, [B,E,P,P,2(+1)].  `jacobian` -> (coords, valid, (Ji, Jj, Jz)); `valid` -> (coords, valid);
    `tonly` drops the rotation of the relative pose (translation-only flow)."""
    if not isinstance(poses, (SE3, Sim3)):
        raise TypeError("transform: poses must be a dpvo_b200.lietorch SE3 or Sim3")
    pts = iproj(patches[:, kk], intrinsics[:, ii])
    Gij = poses[:, jj] * poses[:, ii].inv()
    if tonly:
        ident = torch.zeros_like(Gij.data[..., 3:])
        ident[..., 3] = 1.0                       # unit quaternion; Sim3 carries the scale as an 8th entry, set to 1
        if isinstance(Gij, Sim3):
            ident[..., 4] = 1.0
        Gij = type(Gij)(torch.cat([Gij.data[..., :3], ident], dim=-1))
    moved = Gij[:, :, None, None] * pts
    coords = proj(moved, intrinsics[:, jj], depth)
    if jacobian:
        c = moved.shape[2] // 2
        Xc = moved[:, :, c, c]
        return coords, (Xc[..., 2] > MIN_DEPTH).float(), _centre_jacobians(Gij, Xc, intrinsics[:, jj])
    if valid:
        return coords, (moved[..., 2] > MIN_DEPTH).float()
    return coords


def point_cloud(poses, patches, intrinsics, ix):
    """world-frame homogeneous points of every patch pixel: G_ix^-1 applied to the back-projection"""
    return poses[:, ix, None, None].inv() * iproj(patches, intrinsics[:, ix])


def flow_mag(poses, patches, intrinsics, ii, jj, kk, beta=0.3):
    """blend of full and translation-only flow magnitude per pixel (keyframe test, dpvo.py:259-264), + validity.
    The reference reprojects each patch into its own frame to get the flow origin; that is the patch's own
    pixel grid (G_ii = identity, Z = 1), taken here directly."""
    origin = patches[:, kk, :2].permute(0, 1, 3, 4, 2)
    full, val = transform(poses, patches, intrinsics, ii, jj, kk, valid=True)
    trans = transform(poses, patches, intrinsics, ii, jj, kk, tonly=True)
    mag = beta * (full - origin).norm(dim=-1) + (1 - beta) * (trans - origin).norm(dim=-1)
    return mag, val > 0.5
