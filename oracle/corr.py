"""ORACLE (test infrastructure -- never imported by the product path).

CPU restatement of the reference's altcorr ops in plain PyTorch (fp64 by default):
    corr_raw / corr_forward   dpvo/altcorr/correlation_kernel.cu:82-136 (8x8 integer taps, zero
                              outside the map) + :221-232 (bilinear blend to 7x7, swap to (x, y))
    corr_grid_sample          the algebraically identical F.grid_sample form named by BASELINE.json
                              config 1 (bilinear, zeros padding, align_corners=True); used as the
                              CPU baseline and as an independent cross-check of corr_forward
    patchify_forward          correlation_kernel.cu:16-47 + the Python blend of correlation.py:51-68
Backward passes are obtained by autograd through these (exact transposes of linear maps), which
restates correlation_kernel.cu:139-190, 252-269 and :49-80.

Pinning: the reference holds no golden vectors for altcorr.  The oracle is pinned by
(i) the two independent formulations agreeing to fp64 round-off (tests/test_oracle_corr.py) and
(ii) on the GPU box, the reference's own kernels compiled from /root/reference into oracle/_ref
(oracle/build_ref.py) agreeing with it on seeded inputs (tests/test_parity_ref_gpu.py), with the
committed fixtures in tests/golden/ generated from those kernels.
"""
import torch
import torch.nn.functional as F


def corr_raw(fmap1, fmap2, coords, ii, jj, radius):
    """[B, M, D(y tap), D(x tap), P, P] raw taps, D = 2R+2  (correlation_kernel.cu:82-136)."""
    B, M, _, P, _ = coords.shape
    C = fmap1.shape[2]
    H2, W2 = fmap2.shape[3], fmap2.shape[4]
    D = 2 * radius + 2
    dt = fmap1.dtype
    f1 = fmap1[:, ii]                                   # [B, M, C, P, P]
    f2 = fmap2[:, jj]                                   # [B, M, C, H2, W2]
    x0 = torch.floor(coords[:, :, 0]).long()            # [B, M, P, P]
    y0 = torch.floor(coords[:, :, 1]).long()
    taps = torch.arange(D, device=coords.device) - radius
    yy = y0[:, :, None, None] + taps[None, None, :, None, None, None]      # [B,M,D,1,P,P]
    xx = x0[:, :, None, None] + taps[None, None, None, :, None, None]      # [B,M,1,D,P,P]
    yy = yy.expand(B, M, D, D, P, P)
    xx = xx.expand(B, M, D, D, P, P)
    inb = (yy >= 0) & (yy < H2) & (xx >= 0) & (xx < W2)
    lin = (yy.clamp(0, H2 - 1) * W2 + xx.clamp(0, W2 - 1)).reshape(B, M, 1, -1).expand(B, M, C, -1)
    g = torch.gather(f2.reshape(B, M, C, H2 * W2), 3, lin).reshape(B, M, C, D, D, P, P)
    out = (g * f1[:, :, :, None, None]).sum(2)
    return torch.where(inb, out, torch.zeros((), dtype=dt, device=out.device))


def corr_forward(fmap1, fmap2, coords, ii, jj, radius, chunk=256):
    """cuda_corr.forward: [B, M, 2R+1 (x off), 2R+1 (y off), P, P]  (:193-233)."""
    M = coords.shape[1]
    if M > chunk:   # bound the [chunk, C, D, D, P, P] gather
        return torch.cat([corr_forward(fmap1, fmap2, coords[:, s:s + chunk], ii[s:s + chunk], jj[s:s + chunk], radius, chunk)
                          for s in range(0, M, chunk)], 1)
    D = 2 * radius + 2
    raw = corr_raw(fmap1, fmap2, coords, ii, jj, radius)
    # fractions are formed in the dtype of coords (fp32 in DPVO) and only then cast (:221-224)
    x, y = coords[:, :, 0, None, None], coords[:, :, 1, None, None]
    dx = (x - torch.floor(x)).to(fmap1.dtype)
    dy = (y - torch.floor(y)).to(fmap1.dtype)
    out = (1 - dx) * (1 - dy) * raw[:, :, 0:D - 1, 0:D - 1]
    out = out + dx * (1 - dy) * raw[:, :, 0:D - 1, 1:D]
    out = out + (1 - dx) * dy * raw[:, :, 1:D, 0:D - 1]
    out = out + dx * dy * raw[:, :, 1:D, 1:D]
    return out.permute(0, 1, 3, 2, 4, 5)


def corr_grid_sample(fmap1, fmap2, coords, ii, jj, radius):
    """Same result through F.grid_sample: correlation is linear in fmap2 and out-of-image taps are
    zero, so sampling fmap2 bilinearly at coords + integer offsets and dotting with fmap1 equals the
    blend of raw integer taps."""
    B, M, _, P, _ = coords.shape
    C = fmap1.shape[2]
    H2, W2 = fmap2.shape[3], fmap2.shape[4]
    O = 2 * radius + 1
    offs = torch.arange(O, device=coords.device, dtype=coords.dtype) - radius
    out = torch.zeros(B, M, O, O, P, P, dtype=fmap1.dtype, device=fmap1.device)
    for b in range(B):
        for j in torch.unique(jj).tolist():                           # one grid_sample per target frame
            sel = torch.nonzero(jj == j).squeeze(1)
            n = sel.numel()
            f1 = fmap1[b, ii[sel]]                                    # [n, C, P, P]
            x = coords[b, sel, 0].to(fmap1.dtype)                     # [n, P, P]
            y = coords[b, sel, 1].to(fmap1.dtype)
            gx = (x[:, None, None] + offs.to(fmap1.dtype)[None, :, None, None, None]).expand(n, O, O, P, P)
            gy = (y[:, None, None] + offs.to(fmap1.dtype)[None, None, :, None, None]).expand(n, O, O, P, P)
            grid = torch.stack([2 * gx / (W2 - 1) - 1, 2 * gy / (H2 - 1) - 1], -1).reshape(1, n * O * O * P, P, 2)
            samp = F.grid_sample(fmap2[b, j][None], grid, mode="bilinear", padding_mode="zeros", align_corners=True)
            samp = samp.reshape(C, n, O, O, P, P).permute(1, 0, 2, 3, 4, 5)
            out[b, sel] = (samp * f1[:, :, None, None]).sum(1)
    return out


def patchify_raw(net, coords, radius):
    """cuda_corr.patchify_forward: [B, M, C, D, D] window at floor(coords), zero outside (:16-47)."""
    B, C, H, W = net.shape
    M = coords.shape[1]
    D = 2 * radius + 2
    taps = torch.arange(D, device=coords.device) - radius
    x0 = torch.floor(coords[..., 0]).long()
    y0 = torch.floor(coords[..., 1]).long()
    yy = (y0[:, :, None, None] + taps[None, None, :, None]).expand(B, M, D, D)
    xx = (x0[:, :, None, None] + taps[None, None, None, :]).expand(B, M, D, D)
    inb = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
    lin = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).reshape(B, 1, M * D * D).expand(B, C, -1)
    g = torch.gather(net.reshape(B, C, H * W), 2, lin).reshape(B, C, M, D, D).permute(0, 2, 1, 3, 4)
    return torch.where(inb[:, :, None], g, torch.zeros((), dtype=net.dtype, device=net.device))


def patchify(net, coords, radius, mode="bilinear"):
    """altcorr.patchify (correlation.py:51-68)."""
    patches = patchify_raw(net, coords, radius)
    if mode == "bilinear":
        offset = coords - coords.floor()
        dx, dy = offset[:, :, None, None, None].unbind(dim=-1)
        d = 2 * radius + 1
        x00 = (1 - dy) * (1 - dx) * patches[..., :d, :d]
        x01 = (1 - dy) * dx * patches[..., :d, 1:]
        x10 = dy * (1 - dx) * patches[..., 1:, :d]
        x11 = dy * dx * patches[..., 1:, 1:]
        return x00 + x01 + x10 + x11
    return patches
