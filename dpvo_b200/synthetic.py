"""Synthetic DPVO state for benchmarks and parity tests (no datasets, no checkpoint).

Replays the reference's edge rules (dpvo/dpvo.py:362-375 __edges_forw/__edges_back, :457-459
append order, :305-310 removal of edges whose patch left the REMOVAL_WINDOW) to obtain the patch
graph exactly as `DPVO.update()` sees it in steady state, and fills the state tensors with the shapes
and dtypes of dpvo/dpvo.py:58-73 and dpvo/patchgraph.py:26-35:

    default.yaml (config/default.yaml:4-7): M=96, lifetime 13, removal 22, optimisation 10
        -> E = 47,712 edges over 2,208 live patches inside update()
    fast.yaml: M=48, lifetime 11, removal 16, optimisation 7 -> E = 14,496

Everything is seeded; tensors are created on `device`.
"""
import math
from dataclasses import dataclass, field

import torch

CONFIGS = {
    "default": dict(M=96, lifetime=13, removal=22, opt_window=10, ht=480, wd=640,
                    intrinsics=(320.0, 320.0, 320.0, 240.0)),
    "fast": dict(M=48, lifetime=11, removal=16, opt_window=7, ht=480, wd=752,
                 intrinsics=(458.654, 457.296, 367.215, 248.375)),
}


def replay_edges(n_frames, M, lifetime, removal):
    """Edge lists (ii, jj, kk) at the entry of update() after `n_frames` frames were added, assuming
    no keyframe is dropped (the upper bound of SURVEY 8).  CPU int64 tensors, reference order."""
    ii = torch.zeros(0, dtype=torch.long)
    jj = torch.zeros(0, dtype=torch.long)
    kk = torch.zeros(0, dtype=torch.long)

    def append(k, j):
        nonlocal ii, jj, kk
        kg, jg = torch.meshgrid(k, j, indexing="ij")
        kg, jg = kg.reshape(-1), jg.reshape(-1)
        jj = torch.cat([jj, jg])
        kk = torch.cat([kk, kg])
        ii = torch.cat([ii, kg // M])

    n = 0
    for _ in range(n_frames):
        if n > 0 or True:
            n += 1
            # forward: live patches of the previous `lifetime` frames -> new frame (n-1)
            append(torch.arange(M * max(n - lifetime, 0), M * max(n - 1, 0)), torch.arange(n - 1, n))
            # backward: patches of the new frame -> previous `lifetime` frames (and itself)
            append(torch.arange(M * max(n - 1, 0), M * n), torch.arange(max(n - lifetime, 0), n))
        if _ < n_frames - 1:
            keep = (kk // M) >= n - removal        # keyframe(): drop edges outside the removal window
            ii, jj, kk = ii[keep], jj[keep], kk[keep]
    return ii, jj, kk


@dataclass
class SyntheticState:
    cfg: dict
    n: int                      # number of frames (t1 of the BA window)
    t0: int                     # first free pose of the BA window
    ii: torch.Tensor            # [E] int64 source frame of each edge
    jj: torch.Tensor            # [E] int64 target frame
    kk: torch.Tensor            # [E] int64 patch id (frame * M + slot)
    poses: torch.Tensor         # [N,7] fp32
    patches: torch.Tensor       # [N*M,3,3,3] fp32
    intrinsics: torch.Tensor    # [N,4] fp32 (already /4)
    fmap1: torch.Tensor = None  # [1,mem,128,h,w]   logical NCHW, memory per `channels_last`
    fmap2: torch.Tensor = None  # [1,mem,128,h/4,w/4]
    gmap: torch.Tensor = None   # [1,mem*M,128,3,3]
    imap: torch.Tensor = None   # [1,mem*M,384]
    net: torch.Tensor = None    # [1,E,384]
    extras: dict = field(default_factory=dict)

    @property
    def E(self):
        return self.ii.numel()


def _trajectory(n, device, dtype=torch.float32):
    """Smooth camera path: 0.05 units/frame along x, 1 degree/frame yaw.  [n,7]"""
    out = torch.zeros(n, 7, dtype=torch.float64)
    for t in range(n):
        yaw = math.radians(1.0) * t
        out[t, 0] = -0.05 * t
        out[t, 2] = 0.01 * t
        out[t, 4] = math.sin(0.5 * yaw)      # rotation about y
        out[t, 6] = math.cos(0.5 * yaw)
    return out.to(device=device, dtype=dtype)


def make_state(config="default", n_frames=30, device="cuda", seed=1234, features=True, dtype=torch.half,
               channels_last=True, buffer=64, mem=36, noise=0.0):
    """Build a steady-state synthetic DPVO state.  `buffer` is the number of pose slots allocated
    (the reference allocates 4096); `mem` the feature ring size (dpvo.py:58)."""
    cfg = dict(CONFIGS[config]) if isinstance(config, str) else dict(config)
    M = cfg["M"]
    g = torch.Generator(device="cpu").manual_seed(seed)
    ii, jj, kk = replay_edges(n_frames, M, cfg["lifetime"], cfg["removal"])
    n = n_frames
    assert n < buffer and n <= mem, "synthetic state keeps all live frames inside the ring buffers"
    h, w = cfg["ht"] // 4, cfg["wd"] // 4
    fx, fy, cx, cy = [v / 4.0 for v in cfg["intrinsics"]]

    poses = torch.zeros(buffer, 7)
    poses[:, 6] = 1.0
    poses[:n] = _trajectory(n, "cpu")
    if noise > 0:
        poses[1:n, :3] += noise * torch.randn(n - 1, 3, generator=g)
    px = torch.randint(1, w - 1, (buffer * M,), generator=g).float()
    py = torch.randint(1, h - 1, (buffer * M,), generator=g).float()
    offs = torch.tensor([-1.0, 0.0, 1.0])
    patches = torch.zeros(buffer * M, 3, 3, 3)
    patches[:, 0] = px[:, None, None] + offs[None, None, :]
    patches[:, 1] = py[:, None, None] + offs[None, :, None]
    patches[:, 2] = (0.25 + 0.75 * torch.rand(buffer * M, generator=g))[:, None, None]
    intr = torch.tensor([fx, fy, cx, cy]).repeat(buffer, 1)

    st = SyntheticState(cfg=cfg, n=n, t0=max(n - cfg["opt_window"], 1),
                        ii=ii.to(device), jj=jj.to(device), kk=kk.to(device),
                        poses=poses.to(device), patches=patches.to(device), intrinsics=intr.to(device))
    if features:
        gd = torch.Generator(device=device).manual_seed(seed + 1)
        # features are generated directly on the device (177 MB for the default level-0 pyramid)
        if channels_last:
            f1 = (torch.randn(1, mem, h, w, 128, generator=gd, device=device, dtype=torch.float32) / 4).to(dtype).permute(0, 1, 4, 2, 3)
        else:
            f1 = (torch.randn(1, mem, 128, h, w, generator=gd, device=device, dtype=torch.float32) / 4).to(dtype)
        f2 = torch.nn.functional.avg_pool2d(f1[0].float(), 4, 4).to(dtype)[None]
        if channels_last:
            f2 = f2.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
            gm = (torch.randn(1, mem * M, 3, 3, 128, generator=gd, device=device) / 4).to(dtype).permute(0, 1, 4, 2, 3)
        else:
            gm = (torch.randn(1, mem * M, 128, 3, 3, generator=gd, device=device) / 4).to(dtype)
        st.fmap1, st.fmap2, st.gmap = f1, f2, gm
        st.imap = (torch.randn(1, mem * M, 384, generator=gd, device=device) / 4).to(dtype)
        st.net = torch.zeros(1, st.E, 384, device=device, dtype=dtype)
    return st
