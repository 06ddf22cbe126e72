"""ORACLE (test infrastructure -- never imported by the product path).

CPU restatements, vectorised PyTorch (fp64 by default), of the reference's geometry / optimisation:
    fastba_forward   cuda_ba.forward, eff_impl=False: dpvo/fastba/ba_cuda.cu:232-376 (per-edge
                     residuals, Jacobians, dense B/E/C/v/u) and :433-582 (Schur complement, damping
                     S += I*(1e-4 S + 1), Cholesky solve, retractions :156-229)
    fastba_reproject cuda_ba.reproject: ba_cuda.cu:379-429
    iproj/proj/transform   dpvo/projective_ops.py:19-113
    python_ba        dpvo/ba.py:86-182 (the differentiable BA used in training and named by
                     BASELINE.json as the CPU baseline), with scatter_sum restated as index_add
Pose algebra comes from oracle/lie.py for the lietorch-based parts; the fastba part uses its own
unnormalised-quaternion helpers exactly like ba_cuda.cu:36-174.

Pinning: no golden vectors exist in the reference for these.  Pinned by (i) the algebraic
cross-check fastba_forward(1 iteration) == python_ba with the differing constants aligned
(tests/test_oracle_ba.py), (ii) importing the reference's own dpvo/ba.py and
dpvo/projective_ops.py in the build container (oracle/pin_python_ref.py) and (iii) on the GPU box
the reference's cuda_ba compiled into oracle/_ref (tests/test_parity_ref_gpu.py).
"""
import torch

from . import lie


# --------------------------------------------------------------- ba_cuda.cu:36-174 helpers
def _act_so3(q, X):
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * torch.cross(qv, X, dim=-1)
    return X + qw * uv + torch.cross(qv, uv, dim=-1)


def _rel_se3(ti, qi, tj, qj):
    qi_inv = torch.cat([-qi[..., :3], qi[..., 3:]], -1)
    qij = lie.q_mul_raw(qj, qi_inv)
    tij = tj - _act_so3(qij, ti)
    return tij, qij


def _adj_se3_T(t, q, X):
    """Y = Adj(T)^T X (ba_cuda.cu:57-72)"""
    qinv = torch.cat([-q[..., :3], q[..., 3:]], -1)
    y0 = _act_so3(qinv, X[..., :3])
    y1 = _act_so3(qinv, X[..., 3:])
    u = torch.cross(X[..., :3], t, dim=-1)
    return torch.cat([y0, y1 + _act_so3(qinv, u)], -1)


def _exp_so3(phi):
    th2 = (phi * phi).sum(-1, keepdim=True)
    th = th2.sqrt()
    small = th2 < 1e-8
    ths = torch.where(small, torch.ones_like(th), th)
    imag = torch.where(small, 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th2 * th2, torch.sin(0.5 * ths) / ths)
    real = torch.where(small, 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th2 * th2, torch.cos(0.5 * ths))
    return torch.cat([imag * phi, real], -1)


def _exp_se3(xi):
    tau, phi = xi[..., :3], xi[..., 3:]
    q = _exp_so3(phi)
    th2 = (phi * phi).sum(-1, keepdim=True)
    th = th2.sqrt()
    big = th > 1e-4
    ths = torch.where(big, th, torch.ones_like(th))
    th2s = torch.where(big, th2, torch.ones_like(th2))
    a = (1 - torch.cos(ths)) / th2s
    b = (ths - torch.sin(ths)) / (ths * th2s)
    c1 = torch.cross(phi, tau, dim=-1)
    c2 = torch.cross(phi, c1, dim=-1)
    t = tau + torch.where(big, a * c1 + b * c2, torch.zeros_like(tau))
    return t, q


def _retr_se3(xi, t, q):
    dt, dq = _exp_se3(xi)
    q1 = lie.q_mul_raw(dq, q)
    t1 = _act_so3(dq, t) + dt
    return t1, q1


def fastba_linearize(poses, patches, intrinsics, target, weight, ii, jj, kk):
    """Per-edge quantities of ba_cuda.cu:265-333: returns r[E,2], w[E,2], Ji[E,2,6] (= Adj^T Jj, the
    reference's sign), Jj[E,2,6], Jz[E,2]."""
    fx, fy, cx, cy = intrinsics[0].unbind(-1)
    P = patches.shape[-1]
    c = P // 2
    ti, qi = poses[ii, :3], poses[ii, 3:]
    tj, qj = poses[jj, :3], poses[jj, 3:]
    px, py, pd = patches[kk, 0, c, c], patches[kk, 1, c, c], patches[kk, 2, c, c]
    Xi = torch.stack([(px - cx) / fx, (py - cy) / fy, torch.ones_like(px)], -1)
    tij, qij = _rel_se3(ti, qi, tj, qj)
    Xj = _act_so3(qij, Xi) + pd[:, None] * tij
    X, Y, Z, W = Xj[:, 0], Xj[:, 1], Xj[:, 2], pd
    d = torch.where(Z >= 0.2, 1.0 / Z, torch.zeros_like(Z))
    d2 = d * d
    x1 = fx * (X / Z) + cx
    y1 = fy * (Y / Z) + cy
    rx = target[:, 0] - x1
    ry = target[:, 1] - y1
    inb = ((rx * rx + ry * ry).sqrt() < 128) & (Z > 0.2) & (x1 > -64) & (y1 > -64) & \
          (x1 < 2 * cx + 64) & (y1 < 2 * cy + 64)
    mask = inb.to(poses.dtype)
    w = mask[:, None] * weight
    r = torch.stack([rx, ry], -1)
    o = torch.zeros_like(X)
    Jz = torch.stack([fx * (tij[:, 0] * d - tij[:, 2] * (X * d2)),
                      fy * (tij[:, 1] * d - tij[:, 2] * (Y * d2))], -1)
    Jj0 = torch.stack([fx * W * d, o, fx * -X * W * d2, fx * -X * Y * d2, fx * (1 + X * X * d2), fx * -Y * d], -1)
    Jj1 = torch.stack([o, fy * W * d, fy * -Y * W * d2, fy * (-1 - Y * Y * d2), fy * (X * Y * d2), fy * X * d], -1)
    Jj = torch.stack([Jj0, Jj1], 1)
    Ji = torch.stack([_adj_se3_T(tij, qij, Jj0), _adj_se3_T(tij, qij, Jj1)], 1)
    return r, w, Ji, Jj, Jz


def fastba_system(poses, patches, intrinsics, target, weight, ii, jj, kk, t0, t1):
    """Dense B [6N,6N], E [6N,M], C [M], v [6N], u [M], kx [M] as accumulated by ba_cuda.cu:335-373."""
    N = t1 - t0
    kx, ku = torch.unique(kk, return_inverse=True, sorted=True)
    M = kx.numel()
    dt = poses.dtype
    r, w, Ji, Jj, Jz = fastba_linearize(poses, patches, intrinsics, target, weight, ii, jj, kk)
    ix, jx = ii - t0, jj - t0
    B = torch.zeros(6 * N, 6 * N, dtype=dt)
    Em = torch.zeros(6 * N, M, dtype=dt)
    C = torch.zeros(M, dtype=dt)
    v = torch.zeros(6 * N, dtype=dt)
    u = torch.zeros(M, dtype=dt)
    ar6 = torch.arange(6)
    for row in range(2):
        wr, rr, Jzr, Jir, Jjr = w[:, row], r[:, row], Jz[:, row], Ji[:, row], Jj[:, row]
        vi_ok, vj_ok = ix >= 0, jx >= 0
        def add_block(a_idx, b_idx, A, Bm, ok, sign):
            if ok.sum() == 0:
                return
            blk = sign * wr[ok, None, None] * A[ok, :, None] * Bm[ok, None, :]           # [e,6,6]
            rows = (6 * a_idx[ok])[:, None, None] + ar6[None, :, None]
            cols = (6 * b_idx[ok])[:, None, None] + ar6[None, None, :]
            B.index_put_((rows.expand_as(blk).reshape(-1), cols.expand_as(blk).reshape(-1)), blk.reshape(-1), accumulate=True)
        add_block(ix, ix, Jir, Jir, vi_ok, 1.0)
        add_block(jx, jx, Jjr, Jjr, vj_ok, 1.0)
        both = vi_ok & vj_ok
        add_block(ix, jx, Jir, Jjr, both, -1.0)
        add_block(jx, ix, Jjr, Jir, both, -1.0)
        if vi_ok.any():
            rows = (6 * ix[vi_ok])[:, None] + ar6[None]
            Em.index_put_((rows.reshape(-1), ku[vi_ok][:, None].expand(-1, 6).reshape(-1)),
                          (-wr[vi_ok, None] * Jzr[vi_ok, None] * Jir[vi_ok]).reshape(-1), accumulate=True)
            v.index_put_((rows.reshape(-1),), (-wr[vi_ok, None] * rr[vi_ok, None] * Jir[vi_ok]).reshape(-1), accumulate=True)
        if vj_ok.any():
            rows = (6 * jx[vj_ok])[:, None] + ar6[None]
            Em.index_put_((rows.reshape(-1), ku[vj_ok][:, None].expand(-1, 6).reshape(-1)),
                          (wr[vj_ok, None] * Jzr[vj_ok, None] * Jjr[vj_ok]).reshape(-1), accumulate=True)
            v.index_put_((rows.reshape(-1),), (wr[vj_ok, None] * rr[vj_ok, None] * Jjr[vj_ok]).reshape(-1), accumulate=True)
        C.index_add_(0, ku, wr * Jzr * Jzr)
        u.index_add_(0, ku, wr * rr * Jzr)
    return B, Em, C, v, u, kx


def fastba_forward(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations):
    """cuda_ba.forward (eff_impl=False).  poses [n,7], patches [m,3,P,P], intrinsics [k,4],
    target/weight [E,2], lmbda [1].  Returns updated (poses, patches) copies."""
    poses = poses.clone()
    patches = patches.clone()
    N = t1 - t0
    for _ in range(iterations):
        B, Em, C, v, u, kx = fastba_system(poses, patches, intrinsics, target, weight, ii, jj, kk, t0, t1)
        Q = 1.0 / (C + lmbda.reshape(-1)[0])
        if N == 0:
            dZ = Q * u
        else:
            EQ = Em * Q[None]
            S = B - EQ @ Em.t()
            y = v - EQ @ u
            S = S + torch.eye(6 * N, dtype=S.dtype) * (1e-4 * S + 1.0)
            L = torch.linalg.cholesky(S)
            dX = torch.cholesky_solve(y[:, None], L)[:, 0]
            dZ = Q * (u - Em.t() @ dX)
            t_new, q_new = _retr_se3(dX.view(N, 6), poses[t0:t1, :3], poses[t0:t1, 3:])
            poses[t0:t1, :3] = t_new
            poses[t0:t1, 3:] = q_new
        d = patches[kx, 2, 0, 0] + dZ
        d = torch.where(d > 20, torch.ones_like(d), d)
        d = torch.clamp(d, min=1e-4)
        patches[kx, 2] = d[:, None, None].expand(-1, patches.shape[-2], patches.shape[-1])
    return poses, patches


def fastba_reproject(poses, patches, intrinsics, ii, jj, kk):
    """cuda_ba.reproject: [E, 2, P, P] (ba_cuda.cu:379-429; divides by raw Z, intrinsics row 0)."""
    fx, fy, cx, cy = intrinsics[0].unbind(-1)
    tij, qij = _rel_se3(poses[ii, :3], poses[ii, 3:], poses[jj, :3], poses[jj, 3:])
    pk = patches[kk]                                                      # [E,3,P,P]
    Xi = torch.stack([(pk[:, 0] - cx) / fx, (pk[:, 1] - cy) / fy, torch.ones_like(pk[:, 0])], -1)   # [E,P,P,3]
    Xj = _act_so3(qij[:, None, None], Xi) + pk[:, 2, :, :, None] * tij[:, None, None]
    return torch.stack([fx * (Xj[..., 0] / Xj[..., 2]) + cx, fy * (Xj[..., 1] / Xj[..., 2]) + cy], 1)


# ------------------------------------------------------- dpvo/projective_ops.py restated
def iproj(patches, intrinsics):
    """projective_ops.py:19-29.  patches [1,E,3,P,P], intrinsics [1,E,4] -> [1,E,P,P,4]"""
    x, y, d = patches.unbind(dim=2)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    return torch.stack([(x - cx) / fx, (y - cy) / fy, torch.ones_like(d), d], dim=-1)


def proj(X, intrinsics, depth=False):
    """projective_ops.py:32-50"""
    Xc, Y, Z, W = X.unbind(dim=-1)
    fx, fy, cx, cy = intrinsics[..., None, None].unbind(dim=2)
    d = 1.0 / Z.clamp(min=0.1)
    x = fx * (d * Xc) + cx
    y = fy * (d * Y) + cy
    if depth:
        return torch.stack([x, y, d], dim=-1)
    return torch.stack([x, y], dim=-1)


def transform(poses, patches, intrinsics, ii, jj, kk, jacobian=False, valid=False):
    """projective_ops.py:53-113 for SE3 poses given as data [1,n,7].  Returns coords [1,E,P,P,2]
    (+ valid, (Ji, Jj, Jz) when jacobian=True)."""
    X0 = iproj(patches[:, kk], intrinsics[:, ii])
    Gij = lie.se3_mul(poses[:, jj], lie.se3_inv(poses[:, ii]))
    X1 = lie.se3_act4(Gij[:, :, None, None], X0)
    x1 = proj(X1, intrinsics[:, jj])
    if jacobian:
        p = X1.shape[2]
        X, Y, Z, H = X1[..., p // 2, p // 2, :].unbind(dim=-1)
        o = torch.zeros_like(H)
        fx, fy, cx, cy = intrinsics[:, jj].unbind(dim=-1)
        d = torch.zeros_like(Z)
        d[Z.abs() > 0.2] = 1.0 / Z[Z.abs() > 0.2]
        Ja = torch.stack([H, o, o, o, Z, -Y, o, H, o, -Z, o, X, o, o, H, Y, -X, o, o, o, o, o, o, o], dim=-1).view(1, len(ii), 4, 6)
        Jp = torch.stack([fx * d, o, -fx * X * d * d, o, o, fy * d, -fy * Y * d * d, o], dim=-1).view(1, len(ii), 2, 4)
        Jj = torch.matmul(Jp, Ja)
        Ji = -lie.se3_adjT(Gij[:, :, None], Jj)
        Jz = torch.matmul(Jp, lie.se3_matrix(Gij)[..., :, 3:])
        return x1, (Z > 0.2).to(x1.dtype), (Ji, Jj, Jz)
    if valid:
        return x1, (X1[..., 2] > 0.2).to(x1.dtype)
    return x1


def _scatter_sum(src, index, dim_size):
    """torch_scatter.scatter_sum(src, index, dim=1, dim_size=...) (pytorch-scatter 2.1.2) == index_add"""
    out = torch.zeros(src.shape[:1] + (dim_size,) + src.shape[2:], dtype=src.dtype, device=src.device)
    return out.index_add_(1, index, src)


def python_ba(poses, patches, intrinsics, targets, weights, lmbda, ii, jj, kk, bounds, ep=100.0, fixedp=1,
              structure_only=False, resid_gate=250.0):
    """dpvo/ba.py:86-182.  poses [1,n,7] data, patches [1,m,3,P,P]; returns (poses, patches)."""
    b = 1
    n = int(max(ii.max().item(), jj.max().item())) + 1
    coords, v, (Ji, Jj, Jz) = transform(poses, patches, intrinsics, ii, jj, kk, jacobian=True)
    p = coords.shape[3]
    r = targets - coords[..., p // 2, p // 2, :]
    v = v * (r.norm(dim=-1) < resid_gate).to(r.dtype)
    cxy = coords[..., p // 2, p // 2, :]
    in_bounds = (cxy[..., 0] > bounds[0]) & (cxy[..., 1] > bounds[1]) & (cxy[..., 0] < bounds[2]) & (cxy[..., 1] < bounds[3])
    v = v * in_bounds.to(r.dtype)
    r = (v[..., None] * r).unsqueeze(dim=-1)
    weights = (v[..., None] * weights).unsqueeze(dim=-1)
    wJiT = (weights * Ji).transpose(2, 3)
    wJjT = (weights * Jj).transpose(2, 3)
    wJzT = (weights * Jz).transpose(2, 3)
    Bii, Bij = wJiT @ Ji, wJiT @ Jj
    Bji, Bjj = wJjT @ Ji, wJjT @ Jj
    Eik, Ejk = wJiT @ Jz, wJjT @ Jz
    vi, vj = wJiT @ r, wJjT @ r
    n = n - fixedp
    ii = ii - fixedp
    jj = jj - fixedp
    kx, kk = torch.unique(kk, return_inverse=True, sorted=True)
    m = len(kx)

    def sc_mat(A, a, bcol, na, nb):
        ok = (a >= 0) & (bcol >= 0) & (a < na) & (bcol < nb)
        return _scatter_sum(A[:, ok], a[ok] * nb + bcol[ok], na * nb)

    def sc_vec(A, a, na):
        ok = (a >= 0) & (a < na)
        return _scatter_sum(A[:, ok], a[ok], na)

    B = sc_mat(Bii, ii, ii, n, n).view(b, n, n, 6, 6) + sc_mat(Bij, ii, jj, n, n).view(b, n, n, 6, 6) + \
        sc_mat(Bji, jj, ii, n, n).view(b, n, n, 6, 6) + sc_mat(Bjj, jj, jj, n, n).view(b, n, n, 6, 6)
    E = sc_mat(Eik, ii, kk, n, m).view(b, n, m, 6, 1) + sc_mat(Ejk, jj, kk, n, m).view(b, n, m, 6, 1)
    C = sc_vec(wJzT @ Jz, kk, m)
    vv = sc_vec(vi, ii, n).view(b, n, 1, 6, 1) + sc_vec(vj, jj, n).view(b, n, 1, 6, 1)
    w = sc_vec(wJzT @ r, kk, m)
    if isinstance(lmbda, torch.Tensor):
        lmbda = lmbda.reshape(*C.shape)
    Q = 1.0 / (C + lmbda)
    EQ = E * Q[:, None]

    def bmm(A, Bm):
        b_, n1, m1, p1, q1 = A.shape
        _, n2, m2, p2, q2 = Bm.shape
        A2 = A.permute(0, 1, 3, 2, 4).reshape(b_, n1 * p1, m1 * q1)
        B2 = Bm.permute(0, 1, 3, 2, 4).reshape(b_, n2 * p2, m2 * q2)
        return (A2 @ B2).reshape(b_, n1, p1, m2, q2).permute(0, 1, 3, 2, 4)

    if structure_only or n == 0:
        dZ = (Q * w).view(b, -1, 1, 1)
        dX = None
    else:
        Et = E.permute(0, 2, 1, 4, 3)
        S = B - bmm(EQ, Et)
        y = vv - bmm(EQ, w.unsqueeze(dim=2))
        b_, n1, m1, p1, q1 = S.shape
        A = S.permute(0, 1, 3, 2, 4).reshape(b_, n1 * p1, m1 * q1)
        Y = y.permute(0, 1, 3, 2, 4).reshape(b_, n1 * p1, 1)
        A = A + (ep + 1e-4 * A) * torch.eye(n1 * p1, dtype=A.dtype)
        L, info = torch.linalg.cholesky_ex(A)
        Xs = torch.zeros_like(Y) if bool(info.any()) else torch.cholesky_solve(Y, L)
        dX = Xs.reshape(b_, n1, p1, 1, 1).permute(0, 1, 3, 2, 4)
        dZ = Q * (w - bmm(Et, dX).squeeze(dim=-1))
        dX = dX.view(b, -1, 6)
        dZ = dZ.view(b, -1, 1, 1)
    x, y_, disps = patches.unbind(dim=2)
    disps = (disps + _scatter_sum(dZ, kx, disps.shape[1])).clamp(min=1e-3, max=10.0)
    patches = torch.stack([x, y_, disps], dim=2)
    if dX is not None:
        full = _scatter_sum(dX, fixedp + torch.arange(n), poses.shape[1])
        poses = lie.se3_mul(lie.se3_exp(full), poses)
    return poses, patches


def posegraph_solve(J_i, J_j, ii, jj, res, ep, lm, freen):
    """cuda_ba.solve_system restated densely (dpvo/fastba/ba.cpp:99-180): stack the r x 7 residual rows with their
    two 7x7 Jacobian blocks into J [7r, 7n], A = J^T J, b = -J^T res, A.diag += lm * A.diag + ep, solve the leading
    7*freen block (all of it when freen < 0), zero step for the rest.  fp64 like the reference's Eigen path.
    Parity unpinned: the reference function needs Eigen (absent from the image) and has no test or fixture."""
    r = res.shape[0]
    n = int(max(ii.max(), jj.max())) + 1
    J = torch.zeros(r * 7, n * 7, dtype=torch.float64)
    for x in range(r):
        J[x * 7:(x + 1) * 7, int(ii[x]) * 7:(int(ii[x]) + 1) * 7] += J_i[x].double()
        J[x * 7:(x + 1) * 7, int(jj[x]) * 7:(int(jj[x]) + 1) * 7] += J_j[x].double()
    A = J.t() @ J
    b = -(J.t() @ res.double().reshape(-1))
    d = torch.diagonal(A)
    d += d * lm + ep
    f = n * 7 if freen < 0 else min(freen * 7, n * 7)
    delta = torch.zeros(n * 7, dtype=torch.float64)
    delta[:f] = torch.linalg.solve(A[:f, :f], b[:f])
    return delta.view(n, 7)
