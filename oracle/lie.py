"""ORACLE (test infrastructure -- never imported by the product path).

CPU restatement, in plain PyTorch (any float dtype, fp64 for pinning), of the SO3 / SE3 math of
the reference's lietorch:
    forward formulas   dpvo/lietorch/include/so3.h:31-220, se3.h:36-217, common.h:7 (EPS = 1e-6)
    backward formulas  dpvo/lietorch/src/lietorch_gpu.cu:32-256 (left-tangent-space gradients)

The reference's native lietorch cannot be built here (Eigen 3.4.0 is not vendored and absent from
the image), so this restatement is pinned by the reference's own known-answer tests
(dpvo/lietorch/run_tests.py:16-226, re-expressed in tests/test_oracle_lie.py): exp/log round trip,
X * X^-1 = e, the adjoint identity, act == matrix action, and analytic-vs-numeric Jacobians of the
composed backward operators.

Layouts: SO3 [qx qy qz qw]; SE3 [tx ty tz qx qy qz qw]; tangents SO3 [phi], SE3 [tau, phi].
All functions take tensors of shape [..., dim] and broadcast over leading dims.
"""
import math

import torch

EPS = 1e-6


# ----------------------------------------------------------------------------- small helpers
def hat(p):
    """so3.h:109-117"""
    x, y, z = p.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], -1).reshape(p.shape[:-1] + (3, 3))


def _cross(a, b):
    return torch.cross(a, b, dim=-1)


def q_normalize(q):
    """so3.h:31-37: every construction normalises the quaternion"""
    return q / q.norm(dim=-1, keepdim=True)


def q_mul_raw(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz], -1)


def q_conj(q):
    return torch.cat([-q[..., :3], q[..., 3:]], -1)


def q_rot(q, p):
    """so3.h:55-60 (q assumed normalised)"""
    v, w = q[..., :3], q[..., 3:]
    uv = _cross(v, p)
    uv = uv + uv
    return p + w * uv + _cross(v, uv)


def q_matrix(q):
    x, y, z, w = q.unbind(-1)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return torch.stack([1 - (tyy + tzz), txy - twz, txz + twy,
                        txy + twz, 1 - (txx + tzz), tyz - twx,
                        txz - twy, tyz + twx, 1 - (txx + tyy)], -1).reshape(q.shape[:-1] + (3, 3))


def matv(A, v):
    return (A @ v.unsqueeze(-1)).squeeze(-1)


def rowm(v, A):
    return (v.unsqueeze(-2) @ A).squeeze(-2)


# ------------------------------------------------------------------------------------ SO3
def so3_exp(phi):
    """so3.h:159-176"""
    t2 = (phi * phi).sum(-1, keepdim=True)
    th = t2.sqrt()
    small = th < EPS
    ths = torch.where(small, torch.ones_like(th), th)
    t4 = t2 * t2
    im = torch.where(small, 0.5 - (1.0 / 48.0) * t2 + (1.0 / 3840.0) * t4, torch.sin(0.5 * ths) / ths)
    re = torch.where(small, 1.0 - (1.0 / 8.0) * t2 + (1.0 / 384.0) * t4, torch.cos(0.5 * ths))
    return q_normalize(torch.cat([im * phi, re], -1))


def so3_log(q):
    """so3.h:123-157 (q raw data -> normalised first)"""
    q = q_normalize(q)
    v, w = q[..., :3], q[..., 3:]
    sn = (v * v).sum(-1, keepdim=True)
    n = sn.sqrt()
    small = sn < EPS * EPS
    ns = torch.where(small, torch.ones_like(n), n)
    ws = torch.where(w.abs() < EPS, torch.ones_like(w), w)
    f_small = 2.0 / w - (2.0 / 3.0) * sn / (w * w * w)
    f_wzero = torch.where(w > 0, math.pi / ns, -math.pi / ns)
    f_reg = 2.0 * torch.atan(ns / ws) / ns
    f = torch.where(small, f_small, torch.where(w.abs() < EPS, f_wzero, f_reg))
    return f * v


def so3_left_jacobian(phi):
    """so3.h:178-197"""
    Phi = hat(phi)
    Phi2 = Phi @ Phi
    t2 = (phi * phi).sum(-1)
    th = t2.sqrt()
    small = th < EPS
    ths = torch.where(small, torch.ones_like(th), th)
    t2s = torch.where(small, torch.ones_like(t2), t2)
    c1 = torch.where(small, 0.5 - (1.0 / 24.0) * t2, (1.0 - torch.cos(ths)) / t2s)
    c2 = torch.where(small, (1.0 / 6.0) - (1.0 / 120.0) * t2, (ths - torch.sin(ths)) / (t2s * ths))
    I = torch.eye(3, dtype=phi.dtype, device=phi.device)
    return I + c1[..., None, None] * Phi + c2[..., None, None] * Phi2


def so3_left_jacobian_inverse(phi):
    """so3.h:199-215"""
    Phi = hat(phi)
    Phi2 = Phi @ Phi
    t2 = (phi * phi).sum(-1)
    th = t2.sqrt()
    small = th < EPS
    ths = torch.where(small, torch.ones_like(th), th)
    ht = 0.5 * ths
    c2 = torch.where(small, torch.full_like(th, 1.0 / 12.0),
                     (1.0 - ths * torch.cos(ht) / (2.0 * torch.sin(ht))) / (ths * ths))
    I = torch.eye(3, dtype=phi.dtype, device=phi.device)
    return I - 0.5 * Phi + c2[..., None, None] * Phi2


def so3_projector(q):
    """so3.h:83-93 (4x4)"""
    q = q_normalize(q)
    v, w = q[..., :3], q[..., 3]
    J = torch.zeros(q.shape[:-1] + (4, 4), dtype=q.dtype, device=q.device)
    I = torch.eye(3, dtype=q.dtype, device=q.device)
    J[..., :3, :3] = 0.5 * (w[..., None, None] * I + hat(-v))
    J[..., 3, :3] = 0.5 * (-v)
    return J


# ------------------------------------------------------------------------------------ SE3
def se3_split(X):
    return X[..., :3], q_normalize(X[..., 3:7])


def se3_inv(X):
    """se3.h:36-38"""
    t, q = se3_split(X)
    qi = q_normalize(q_conj(q))
    return torch.cat([-q_rot(qi, t), qi], -1)


def se3_mul(X, Y):
    """se3.h:45-47"""
    tx, qx = se3_split(X)
    ty, qy = se3_split(Y)
    return torch.cat([tx + q_rot(qx, ty), q_normalize(q_mul_raw(qx, qy))], -1)


def se3_act(X, p):
    t, q = se3_split(X)
    return q_rot(q, p) + t


def se3_act4(X, p):
    """se3.h:53-56"""
    t, q = se3_split(X)
    return torch.cat([q_rot(q, p[..., :3]) + t * p[..., 3:], p[..., 3:]], -1)


def se3_adj(X, a):
    """Adj(X) a, se3.h:58-82"""
    t, q = se3_split(X)
    R = q_matrix(q)
    Rphi = matv(R, a[..., 3:])
    return torch.cat([matv(R, a[..., :3]) + _cross(t, Rphi), Rphi], -1)


def se3_adjT(X, a):
    """Adj(X)^T a, se3.h:84-86"""
    t, q = se3_split(X)
    R = q_matrix(q)
    a1, a2 = a[..., :3], a[..., 3:]
    return torch.cat([rowm(a1, R), rowm(_cross(a1, t) + a2, R)], -1)


def se3_Adj_matrix(X):
    """6x6 Adj, se3.h:58-67"""
    t, q = se3_split(X)
    R = q_matrix(q)
    A = torch.zeros(X.shape[:-1] + (6, 6), dtype=X.dtype, device=X.device)
    A[..., :3, :3] = R
    A[..., :3, 3:] = hat(t) @ R
    A[..., 3:, 3:] = R
    return A


def se3_adj_small(b):
    """adj(tau_phi) = [[Phi, Tau],[0, Phi]], se3.h:101-114"""
    A = torch.zeros(b.shape[:-1] + (6, 6), dtype=b.dtype, device=b.device)
    A[..., :3, :3] = hat(b[..., 3:])
    A[..., :3, 3:] = hat(b[..., :3])
    A[..., 3:, 3:] = hat(b[..., 3:])
    return A


def se3_calcQ(a):
    """se3.h:147-176"""
    tau, phi = a[..., :3], a[..., 3:]
    Tau, Phi = hat(tau), hat(phi)
    th = phi.norm(dim=-1)
    t2 = th * th
    t4 = t2 * t2
    small = th < EPS
    ths = torch.where(small, torch.ones_like(th), th)
    t2s, t4s = ths * ths, (ths * ths) ** 2
    c1 = torch.where(small, (1.0 / 6.0) - (1.0 / 120.0) * t2, (ths - torch.sin(ths)) / (t2s * ths))
    c2 = torch.where(small, (1.0 / 24.0) - (1.0 / 720.0) * t2, (t2s + 2 * torch.cos(ths) - 2) / (2 * t4s))
    c3 = torch.where(small, (1.0 / 120.0) - (1.0 / 2520.0) * t2,
                     (2 * ths - 3 * torch.sin(ths) + ths * torch.cos(ths)) / (2 * t4s * ths))
    c1, c2, c3 = c1[..., None, None], c2[..., None, None], c3[..., None, None]
    return 0.5 * Tau + c1 * (Phi @ Tau + Tau @ Phi + Phi @ Tau @ Phi) + \
        c2 * (Phi @ Phi @ Tau + Tau @ Phi @ Phi - 3 * Phi @ Tau @ Phi) + \
        c3 * (Phi @ Tau @ Phi @ Phi + Phi @ Phi @ Tau @ Phi)


def se3_left_jacobian(a):
    J = so3_left_jacobian(a[..., 3:])
    Q = se3_calcQ(a)
    M = torch.zeros(a.shape[:-1] + (6, 6), dtype=a.dtype, device=a.device)
    M[..., :3, :3] = J
    M[..., :3, 3:] = Q
    M[..., 3:, 3:] = J
    return M


def se3_left_jacobian_inverse(a):
    Ji = so3_left_jacobian_inverse(a[..., 3:])
    Q = se3_calcQ(a)
    M = torch.zeros(a.shape[:-1] + (6, 6), dtype=a.dtype, device=a.device)
    M[..., :3, :3] = Ji
    M[..., :3, 3:] = -Ji @ Q @ Ji
    M[..., 3:, 3:] = Ji
    return M


def se3_log(X):
    """se3.h:124-132"""
    t, q = se3_split(X)
    phi = so3_log(q)
    return torch.cat([matv(so3_left_jacobian_inverse(phi), t), phi], -1)


def se3_exp(a):
    """se3.h:134-142"""
    tau, phi = a[..., :3], a[..., 3:]
    return torch.cat([matv(so3_left_jacobian(phi), tau), so3_exp(phi)], -1)


def se3_matrix(X):
    t, q = se3_split(X)
    T = torch.zeros(X.shape[:-1] + (4, 4), dtype=X.dtype, device=X.device)
    T[..., :3, :3] = q_matrix(q)
    T[..., :3, 3] = t
    T[..., 3, 3] = 1
    return T


def se3_projector(X):
    """se3.h:116-122 (7x7)"""
    t, q = se3_split(X)
    J = torch.zeros(X.shape[:-1] + (7, 7), dtype=X.dtype, device=X.device)
    J[..., :3, :3] = torch.eye(3, dtype=X.dtype, device=X.device)
    J[..., :3, 3:6] = hat(-t)
    J[..., 3:, 3:] = so3_projector(q)
    return J


# --------------------------------------------------------------------- group-generic table
class _SO3:
    N, K = 4, 3
    exp = staticmethod(so3_exp)
    log = staticmethod(so3_log)
    inv = staticmethod(lambda X: q_normalize(q_conj(q_normalize(X))))
    mul = staticmethod(lambda X, Y: q_normalize(q_mul_raw(q_normalize(X), q_normalize(Y))))
    act = staticmethod(lambda X, p: q_rot(q_normalize(X), p))
    act4 = staticmethod(lambda X, p: torch.cat([q_rot(q_normalize(X), p[..., :3]), p[..., 3:]], -1))
    adj = staticmethod(lambda X, a: matv(q_matrix(q_normalize(X)), a))
    adjT = staticmethod(lambda X, a: rowm(a, q_matrix(q_normalize(X))))
    Adj_matrix = staticmethod(lambda X: q_matrix(q_normalize(X)))
    adj_small = staticmethod(hat)
    left_jacobian = staticmethod(so3_left_jacobian)
    left_jacobian_inverse = staticmethod(so3_left_jacobian_inverse)
    projector = staticmethod(so3_projector)

    @staticmethod
    def matrix(X):
        T = torch.zeros(X.shape[:-1] + (4, 4), dtype=X.dtype, device=X.device)
        T[..., :3, :3] = q_matrix(q_normalize(X))
        T[..., 3, 3] = 1
        return T

    @staticmethod
    def act_jacobian(p):          # so3.h:217-220
        return hat(-p)

    @staticmethod
    def act4_jacobian(p):         # so3.h:222-227
        J = torch.zeros(p.shape[:-1] + (4, 3), dtype=p.dtype, device=p.device)
        J[..., :3, :] = hat(-p[..., :3])
        return J


class _SE3:
    N, K = 7, 6
    exp = staticmethod(se3_exp)
    log = staticmethod(se3_log)
    inv = staticmethod(se3_inv)
    mul = staticmethod(se3_mul)
    act = staticmethod(se3_act)
    act4 = staticmethod(se3_act4)
    adj = staticmethod(se3_adj)
    adjT = staticmethod(se3_adjT)
    Adj_matrix = staticmethod(se3_Adj_matrix)
    adj_small = staticmethod(se3_adj_small)
    left_jacobian = staticmethod(se3_left_jacobian)
    left_jacobian_inverse = staticmethod(se3_left_jacobian_inverse)
    projector = staticmethod(se3_projector)
    matrix = staticmethod(se3_matrix)

    @staticmethod
    def act_jacobian(p):          # se3.h:207-213
        J = torch.zeros(p.shape[:-1] + (3, 6), dtype=p.dtype, device=p.device)
        J[..., :3, :3] = torch.eye(3, dtype=p.dtype, device=p.device)
        J[..., :3, 3:] = hat(-p)
        return J

    @staticmethod
    def act4_jacobian(p):         # se3.h:215-221
        J = torch.zeros(p.shape[:-1] + (4, 6), dtype=p.dtype, device=p.device)
        J[..., :3, :3] = p[..., 3:, None] * torch.eye(3, dtype=p.dtype, device=p.device)
        J[..., :3, 3:] = hat(-p[..., :3])
        return J


GROUPS = {1: _SO3, 3: _SE3}


def _pad(g, G):
    """gradient w.r.t. a group element: embedding width, last component unused (lietorch_gpu.cu:41-42)"""
    return torch.cat([g, torch.zeros(g.shape[:-1] + (G.N - G.K,), dtype=g.dtype, device=g.device)], -1)


# Backward operators, named as the pybind functions of lietorch.cpp:286-316.  `grad` follows the
# reference's conventions: for group-valued outputs only the first K components are read.
def expm_backward(gid, grad, a):                     # lietorch_gpu.cu:32-44
    G = GROUPS[gid]
    return rowm(grad[..., :G.K], G.left_jacobian(a))


def logm_backward(gid, grad, X):                     # :58-70
    G = GROUPS[gid]
    return _pad(rowm(grad, G.left_jacobian_inverse(G.log(X))), G)


def inv_backward(gid, grad, X):                      # :85-97
    G = GROUPS[gid]
    return _pad(-rowm(grad[..., :G.K], G.Adj_matrix(G.inv(X))), G)


def mul_backward(gid, grad, X, Y):                   # :112-125
    G = GROUPS[gid]
    dZ = grad[..., :G.K]
    return _pad(dZ, G), _pad(rowm(dZ, G.Adj_matrix(X)), G)


def adj_backward(gid, grad, X, a):                   # :140-158
    G = GROUPS[gid]
    A = G.Adj_matrix(X)
    b = matv(A, a)
    return _pad(-rowm(grad, G.adj_small(b)), G), rowm(grad, A)


def adjT_backward(gid, grad, X, a):                  # :174-188
    G = GROUPS[gid]
    Xdb = matv(G.Adj_matrix(X), grad)
    return _pad(-rowm(a, G.adj_small(Xdb)), G), Xdb


def act_backward(gid, grad, X, p):                   # :204-221
    G = GROUPS[gid]
    q = G.act(X, p)
    return _pad(rowm(grad, G.act_jacobian(q)), G), rowm(grad, G.matrix(X)[..., :3, :3])


def act4_backward(gid, grad, X, p):                  # :238-256
    G = GROUPS[gid]
    q = G.act4(X, p)
    return _pad(rowm(grad, G.act4_jacobian(q)), G), rowm(grad, G.matrix(X))


def jinv(gid, X, a):                                 # :282-294
    G = GROUPS[gid]
    return matv(G.left_jacobian_inverse(G.log(X)), a)
