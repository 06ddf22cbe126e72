"""Differentiable bundle adjustment for training -- the role of dpvo/ba.py:86-182 (`BA`) with its
`CholeskySolver` (dpvo/ba.py:12-37), on dpvo_b200.projective_ops / dpvo_b200.lietorch (device kernels with
backward operators).  One Gauss-Newton step on the poses (all but the first `fixedp`) and on the inverse
depth of every patch that appears in `kk`; gradients flow to `targets`, `weights`, `poses`, `patches`.

Formulation (own derivation, same normal equations as the reference):
every edge e contributes a 2x13 Jacobian [Ji | Jj | Jz] of its centre pixel.  The two pose blocks are
stacked into one 2x12 matrix J_e, so that  H_e = J_e^T W_e J_e  (12x12) holds the four 6x6 blocks
(ii,ii), (ii,jj), (jj,ii), (jj,jj) at once and ONE index_add scatters all of them into the dense pose
Hessian B; likewise E (pose x depth coupling), C (depth diagonal) and the two gradient vectors.  The depth
block is diagonal, so the Schur complement  S = B - E diag(Q) E^T,  Q = 1/(C + lmbda)  is a plain matrix
product on [6n, m] operands; damping is the reference's  S + (ep + 1e-4 S) o I.
Gates: validity Z > 0.2, |residual| < 250 px, centre inside `bounds` (dpvo/ba.py:96-106).
Retraction: depths clamped to [1e-3, 10] (dpvo/ba.py:176), poses by the left update Exp(dx) * G.
"""
import torch

from . import projective_ops as pops
from .net import EdgeGroups

RESIDUAL_GATE = 250.0


class _SPDSolve(torch.autograd.Function):
    """x = H^-1 b through a Cholesky factorisation.  A failed factorisation yields a zero step (and zero
    gradients) instead of an exception, as the reference chooses for training robustness (ba.py:15-20).
    backward: dz = H^-1 g,  dH = -x dz^T,  db = dz."""

    @staticmethod
    def forward(ctx, H, b):
        L, info = torch.linalg.cholesky_ex(H)
        ctx.failed = bool(torch.any(info))              # one host sync per solve, as the reference
        if ctx.failed:
            return torch.zeros_like(b)
        x = torch.cholesky_solve(b, L)
        ctx.save_for_backward(L, x)
        return x

    @staticmethod
    def backward(ctx, g):
        if ctx.failed:
            return None, None
        L, x = ctx.saved_tensors
        dz = torch.cholesky_solve(g, L)
        return -x @ dz.transpose(-1, -2), dz


def _scatter_rows(values, index, ok, size):
    """sum rows of `values` [B,E,...] into `size` slots; rows with ok == False contribute nothing"""
    idx = torch.where(ok, index, torch.zeros_like(index))
    vals = values * ok.view(1, -1, *([1] * (values.dim() - 2))).to(values.dtype)
    out = torch.zeros(values.shape[:1] + (size,) + values.shape[2:], dtype=values.dtype, device=values.device)
    return out.index_add_(1, idx, vals)


def BA(poses, patches, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds, ep=100.0, PRINT=False, fixedp=1,
       structure_only=False, n_frames=None):
    """poses: SE3 [1,N,7]; patches [1,K,3,P,P]; targets / weights [1,E,2]; ii / jj / kk int64 [E].
    Returns (poses, patches) after one damped Gauss-Newton step.  `n_frames` (= max(ii, jj) + 1) may be passed
    to avoid the device->host read of it."""
    n = int(torch.maximum(ii.max(), jj.max())) + 1 if n_frames is None else int(n_frames)
    coords, valid, (Ji, Jj, Jz) = pops.transform(poses, patches, intrinsics, ii, jj, kk, jacobian=True)
    c = coords.shape[3] // 2
    centre = coords[:, :, c, c]
    resid = targets - centre
    inside = (centre[..., 0] > bounds[0]) & (centre[..., 1] > bounds[1]) & (centre[..., 0] < bounds[2]) & (centre[..., 1] < bounds[3])
    gate = (valid * (resid.norm(dim=-1) < RESIDUAL_GATE).float() * inside.float())[..., None]
    if PRINT:
        print((resid * gate).norm(dim=-1).mean().item())
    resid = gate * resid                                   # [1,E,2]
    w = gate * weights                                     # [1,E,2]

    # ---- per-edge normal-equation pieces
    J = torch.cat([Ji, Jj], dim=-1)                        # [1,E,2,12]
    wJ = w[..., None] * J
    jz = Jz[..., 0]                                        # [1,E,2]
    H_e = torch.einsum("bekp,bekq->bepq", wJ, J)           # [1,E,12,12]
    g_e = torch.einsum("bekp,bek->bep", wJ, resid)         # [1,E,12]
    E_e = torch.einsum("bekp,bek->bep", wJ, jz)            # [1,E,12]
    C_e = (w * jz * jz).sum(-1)                            # [1,E]
    u_e = (w * jz * resid).sum(-1)                         # [1,E]

    # ---- variables: poses fixedp..n-1, one depth per distinct patch
    nv = n - fixedp
    a, b = ii - fixedp, jj - fixedp
    ok_a, ok_b = (a >= 0) & (a < nv), (b >= 0) & (b < nv)
    grp = EdgeGroups(kk)                                   # device grouping; group ids ascend with the patch id
    m = grp.max_groups
    kx = grp.key_a[:m]
    kid = grp.group_of.long()
    C = torch.zeros(1, m, dtype=C_e.dtype, device=C_e.device).index_add_(1, kid, C_e)
    u = torch.zeros(1, m, dtype=u_e.dtype, device=u_e.device).index_add_(1, kid, u_e)
    lm = lmbda.reshape(*C.shape) if isinstance(lmbda, torch.Tensor) and lmbda.numel() > 1 else lmbda
    Q = 1.0 / (C + lm)

    disps = patches[:, :, 2]
    if structure_only or nv <= 0:
        dZ = Q * u
    else:
        # dense pose Hessian: the four 6x6 blocks of every edge in one scatter
        blocks = H_e.view(1, -1, 2, 6, 2, 6).permute(0, 1, 2, 4, 3, 5)                       # [1,E,2(row pose),2(col pose),6,6]
        rows = torch.stack([a, b], 1)                                                        # [E,2]
        oks = torch.stack([ok_a, ok_b], 1)
        slot = (rows[:, :, None] * nv + rows[:, None, :]).reshape(-1)                        # [E*4]
        slot_ok = (oks[:, :, None] & oks[:, None, :]).reshape(-1)
        Bm = _scatter_rows(blocks.reshape(1, -1, 6, 6), slot, slot_ok, nv * nv)
        Bm = Bm.view(1, nv, nv, 6, 6).permute(0, 1, 3, 2, 4).reshape(1, nv * 6, nv * 6)
        # gradient and pose-depth coupling, again both endpoints at once
        vv = _scatter_rows(g_e.view(1, -1, 6), rows.reshape(-1), oks.reshape(-1), nv).reshape(1, nv * 6, 1)
        eslot = (rows * m + kid[:, None]).reshape(-1)
        Em = _scatter_rows(E_e.view(1, -1, 6), eslot, oks.reshape(-1), nv * m)
        Em = Em.view(1, nv, m, 6).permute(0, 1, 3, 2).reshape(1, nv * 6, m)                   # [1,6nv,m]
        EQ = Em * Q[:, None, :]
        S = Bm - EQ @ Em.transpose(1, 2)
        y = vv - EQ @ u[..., None]
        S = S + torch.diag_embed(ep + 1e-4 * torch.diagonal(S, dim1=-2, dim2=-1))
        dX = _SPDSolve.apply(S, y)                                                           # [1,6nv,1]
        dZ = Q * (u - (Em.transpose(1, 2) @ dX)[..., 0])
        upd = torch.zeros(1, poses.shape[1], 6, dtype=dX.dtype, device=dX.device)
        upd = upd.index_add(1, fixedp + torch.arange(nv, device=dX.device), dX.view(1, nv, 6))
        poses = poses.retr(upd)

    step = torch.zeros_like(disps[:, :, 0, 0]).index_add(1, kx, dZ)
    disps = (disps + step[..., None, None]).clamp(min=1e-3, max=10.0)
    patches = torch.stack([patches[:, :, 0], patches[:, :, 1], disps], dim=2)
    return poses, patches
