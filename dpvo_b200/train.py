"""Training step of the path (BASELINE.json configs[3]): the unrolled update iterations of dpvo/net.py:188-272
(`VONet.forward`), the sequence loss of train.py:85-120 and the optimiser / gradient all-reduce around them
(train.py:62-66, 121-126), on dpvo_b200's operators -- altcorr (kernel forward + backward), the differentiable
update operator (net.Update.forward_train), projective_ops / lietorch (kernel forward + backward) and the
differentiable bundle adjustment (ba.BA).

Data-parallel layout: one clip per rank per step, NCCL all-reduce (sum, then mean) of the fp32
gradients in buckets launched from autograd's post-accumulate hooks (multigpu.GradReducer), so the reduction of
the encoders' gradients overlaps what is left of the backward pass; gradient clipping and the AdamW / OneCycle
step follow on every rank identically.
"""
import numpy as np
import torch
import torch.nn as nn

from . import altcorr
from . import projective_ops as pops
from .ba import BA
from .frontend import Patchifier
from .lietorch import SE3
from .net import Update, DIM


def _mesh(a, b):
    """all pairs (a_i, b_j), a-major, flattened"""
    g = torch.meshgrid(a, b, indexing="ij")
    return g[0].reshape(-1), g[1].reshape(-1)


class CorrBlock:
    """two-level correlation lookup with autograd (net.py:160-174): level 1 = 4x4 average pool of the features"""

    def __init__(self, fmap, gmap, radius=3, dropout=0.2, levels=(1, 4)):
        b, n, c, h, w = fmap.shape
        self.gmap, self.radius, self.dropout, self.levels = gmap, radius, dropout, levels
        flat = fmap.view(b * n, c, h, w)
        self.pyramid = [flat.view(b, n, c, h, w) if l == 1 else
                        nn.functional.avg_pool2d(flat, l, stride=l).view(b, n, c, h // l, w // l) for l in levels]

    def __call__(self, ii, jj, coords):
        out = [altcorr.corr(self.gmap, f, coords / l, ii, jj, self.radius, self.dropout) for f, l in zip(self.pyramid, self.levels)]
        return torch.stack(out, -1).view(1, len(ii), -1)


class VONet(nn.Module):
    """`patchify` + `update` with the reference's attribute names, so a DPVO checkpoint loads unchanged."""

    def __init__(self):
        super().__init__()
        self.P = 3
        self.patchify = Patchifier(self.P)
        self.update = Update(self.P)
        self.DIM, self.RES = DIM, 4

    def forward(self, images, poses, disps, intrinsics, M=1024, STEPS=12, P=1, structure_only=False, rescale=False):
        """images [1,N,3,H,W] in 0..255, poses SE3 [1,N,7] (ground truth, camera-from-world), disps [1,N,H,W].
        Returns the list of per-iteration tuples (valid, coords, coords_gt, G[:, :n], P[:, :n], kl) of net.py:270."""
        images = 2 * (images / 255.0) - 0.5
        intrinsics = intrinsics / 4.0
        disps = disps[:, :, 1::4, 1::4].float()
        fmap, gmap, imap, patches, ix = self.patchify(images, disps=disps)
        return unrolled_updates(self.update, fmap, gmap, imap, patches, ix, poses, intrinsics, STEPS=STEPS, structure_only=structure_only)


def unrolled_updates(update, fmap, gmap, imap, patches, ix, poses_gt, intrinsics, STEPS=12, structure_only=False, depth_init=None):
    """The recurrent part of VONet.forward.  Starts from the first 8 frames fully connected, adds one frame per
    iteration from the 9th on (edges new-frame <- all earlier patches, new patches -> all frames), occasionally
    drops the edges of frame n-4 (np.random, p = 0.1), and after every update runs two differentiable BA steps.
    RNG consumption mirrors the reference: one torch.rand_like for the depth initialisation (skipped when
    `depth_init` is given), one np.random.rand per added frame."""
    dev = fmap.device
    b, N, c, h, w = fmap.shape
    p = patches.shape[-1]
    corr_fn = CorrBlock(fmap, gmap)
    patches_gt = patches.clone()
    d0 = patches[..., 2, p // 2, p // 2]
    if depth_init is None:
        depth_init = torch.rand_like(d0)
    patches = torch.cat([patches[:, :, :2], depth_init[..., None, None, None].expand(-1, -1, 1, p, p)], dim=2)

    first = torch.arange(0, 8, device=dev)
    kk, jj = _mesh(torch.where(ix < 8)[0], first)
    ii = ix[kk]
    imap = imap.view(b, -1, DIM)
    net = torch.zeros(b, len(kk), DIM, device=dev, dtype=torch.float)
    Gs = SE3.IdentityLike(poses_gt)
    if structure_only:
        Gs.data[:] = poses_gt.data[:]
    bounds = [-64, -64, w + 64, h + 64]
    traj = []
    n = 8
    while len(traj) < STEPS:
        Gs = Gs.detach()
        patches = patches.detach()
        if len(traj) >= 8 and n < N:
            if not structure_only:
                Gs.data[:, n] = Gs.data[:, n - 1]
            new = torch.arange(n, n + 1, device=dev)
            kk1, jj1 = _mesh(torch.where(ix < n)[0], new)                                  # old patches seen in the new frame
            kk2, jj2 = _mesh(torch.where(ix == n)[0], torch.arange(0, n + 1, device=dev))  # new patches in every frame
            ii = torch.cat([ix[kk1], ix[kk2], ii])
            jj = torch.cat([jj1, jj2, jj])
            kk = torch.cat([kk1, kk2, kk])
            net = torch.cat([torch.zeros(b, len(kk1) + len(kk2), DIM, device=dev), net], dim=1)
            if np.random.rand() < 0.1:
                keep = (ii != (n - 4)) & (jj != (n - 4))
                ii, jj, kk, net = ii[keep], jj[keep], kk[keep], net[:, keep]
            patches = patches.clone()
            patches[:, ix == n, 2] = torch.median(patches[:, (ix == n - 1) | (ix == n - 2), 2])
            n += 1

        coords = pops.transform(Gs, patches, intrinsics, ii, jj, kk)
        corr = corr_fn(kk, jj, coords.permute(0, 1, 4, 2, 3).contiguous())
        net, (delta, weight, _) = update(net, imap[:, kk], corr, None, ii, jj, kk)
        target = coords[..., p // 2, p // 2, :] + delta
        for _ in range(2):
            Gs, patches = BA(Gs, patches, intrinsics, target, weight, 1e-4, ii, jj, kk, bounds, ep=10, fixedp=1,
                             structure_only=structure_only, n_frames=n)

        near = (ii - jj).abs()
        near = (near > 0) & (near <= 2)
        coords = pops.transform(Gs, patches, intrinsics, ii[near], jj[near], kk[near])
        coords_gt, valid, _ = pops.transform(poses_gt, patches_gt, intrinsics, ii[near], jj[near], kk[near], jacobian=True)
        traj.append((valid, coords, coords_gt, Gs[:, :n], poses_gt[:, :n], torch.as_tensor(0)))
    return traj


def scale_alignment(A, B):
    """least-squares scale between two point sets (train.py:30-41): Var(A) / trace(singular values of cov(A,B))"""
    ca, cb = A - A.mean(0), B - B.mean(0)
    var_a = (ca.norm(dim=1) ** 2).mean()
    sv = torch.linalg.svdvals(ca.T @ cb / A.shape[0])
    return var_a / sv.sum()


def sequence_loss(traj, P=3, flow_weight=0.1, pose_weight=10.0, structure_only=False):
    """train.py:85-120: per iteration, the smallest reprojection error over each valid patch's pixels (flow term)
    and, from the third iteration on, translation + rotation error of every relative pose pair after aligning the
    scale of the estimated trajectory (pose term).  Returns (loss, metrics of the last iteration)."""
    loss = 0.0
    for i, (v, x, y, P1, P2, kl) in enumerate(traj):
        e = (x - y).norm(dim=-1)
        e = e.reshape(-1, P ** 2)[(v > 0.5).reshape(-1)].min(dim=-1).values
        N = P1.shape[1]
        ii, jj = _mesh(torch.arange(N, device=x.device), torch.arange(N, device=x.device))
        off = ii != jj
        ii, jj = ii[off], jj[off]
        P1, P2 = P1.inv(), P2.inv()
        s = scale_alignment(P2.translation()[0, :, :3], P1.translation()[0, :, :3]).detach().clamp(max=10.0)
        P1 = P1.scale(s.view(1, 1))
        dP = P1[:, ii].inv() * P1[:, jj]
        dG = P2[:, ii].inv() * P2[:, jj]
        e1 = (dP * dG.inv()).log()
        tr, ro = e1[..., 0:3].norm(dim=-1), e1[..., 3:6].norm(dim=-1)
        loss = loss + flow_weight * e.mean()
        if not structure_only and i >= 2:
            loss = loss + pose_weight * (tr.mean() + ro.mean())
    loss = loss + kl.to(x.device)                  # "kl is 0 (not longer used)", train.py:117-118
    metrics = {"px1": (e < .25).float().mean(), "ro": ro.float().mean(), "tr": tr.float().mean()}
    return loss, metrics


class TrainStep:
    """one optimisation step: forward (STEPS unrolled updates), loss, backward with bucketed NCCL all-reduce,
    clip, AdamW + OneCycle (train.py:62-66, 75-126)"""

    def __init__(self, net, lr=8e-5, total_steps=240000, clip=10.0, steps_unrolled=18, flow_weight=0.1, pose_weight=10.0,
                 reducer=None):
        self.net = net
        self.opt = torch.optim.AdamW(net.parameters(), lr=lr, weight_decay=1e-6)
        self.sched = torch.optim.lr_scheduler.OneCycleLR(self.opt, lr, total_steps, pct_start=0.01, cycle_momentum=False,
                                                         anneal_strategy="linear")
        self.clip, self.steps_unrolled = clip, steps_unrolled
        self.flow_weight, self.pose_weight = flow_weight, pose_weight
        self.reducer = reducer
        self.total_steps = 0

    def __call__(self, images, poses, disps, intrinsics, structure_only=None):
        return self.step_clips([(images, poses, disps, intrinsics)], structure_only)

    def step_clips(self, clips, structure_only=None):
        """one optimiser step over this rank's share of the global batch: the clips are run one after the other with
        gradient accumulation (the reference batch is structurally 1, ba.py:89); the all-reduce buckets are armed for
        the last clip's backward only, so every gradient crosses NVLink once per step"""
        self.opt.zero_grad(set_to_none=False)
        so = (self.total_steps < 1000) if structure_only is None else structure_only
        world = self.reducer.world if self.reducer is not None else 1
        total = 0.0
        for c, (images, poses, disps, intrinsics) in enumerate(clips):
            traj = self.net(images, SE3(poses).inv(), disps, intrinsics, M=1024, STEPS=self.steps_unrolled, structure_only=so)
            loss, metrics = sequence_loss(traj, self.net.P, self.flow_weight, self.pose_weight, so)
            if self.reducer is not None and c == len(clips) - 1:
                self.reducer.begin()
            (loss / len(clips)).backward()
            total = total + loss.detach() / len(clips)
        if self.reducer is not None:
            self.reducer.finish()                   # gradients are now the mean over ranks (and over each rank's clips)
        torch.nn.utils.clip_grad_norm_(self.net.parameters(), self.clip)
        self.opt.step()
        self.sched.step()
        self.total_steps += 1
        return total, metrics
