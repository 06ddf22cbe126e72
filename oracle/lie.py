"""ORACLE (test infrastructure -- never imported by the product path).

CPU restatement, in plain PyTorch (any float dtype, fp64 for pinning), of the SO3 / RxSO3 / SE3 / Sim3 math of
the reference's lietorch:
    forward formulas   dpvo/lietorch/include/so3.h:31-220, rxso3.h:11-324, se3.h:36-217, sim3.h:12-217,
                       common.h:7 (EPS = 1e-6)
    backward formulas  dpvo/lietorch/src/lietorch_gpu.cu:32-256 (left-tangent-space gradients)

The reference's native lietorch cannot be built here (Eigen 3.4.0 is not vendored and absent from
the image), so this restatement is pinned by the reference's own known-answer tests
(dpvo/lietorch/run_tests.py:16-226, re-expressed in tests/test_oracle_lie.py): exp/log round trip,
X * X^-1 = e, the adjoint identity, act == matrix action, and analytic-vs-numeric Jacobians of the
composed backward operators.

Layouts: SO3 [qx qy qz qw]; RxSO3 [q, s]; SE3 [tx ty tz qx qy qz qw]; Sim3 [t, q, s];
tangents SO3 [phi], RxSO3 [phi, sigma], SE3 [tau, phi], Sim3 [tau, phi, sigma].
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


# ---------------------------------------------------------------------------------- RxSO3
# data [qx qy qz qw s], tangent [phi, sigma]   (rxso3.h:11-324)
def rxso3_split(X):
    return q_normalize(X[..., :4]), X[..., 4:5]


def rxso3_exp(a):
    """rxso3.h:166-187: rotation as SO3::Exp, scale = exp(sigma)"""
    return torch.cat([so3_exp(a[..., :3]), torch.exp(a[..., 3:4])], -1)


def rxso3_log(X):
    """rxso3.h:133-164"""
    q, s = rxso3_split(X)
    return torch.cat([so3_log(q), torch.log(s)], -1)


def rxso3_inv(X):
    q, s = rxso3_split(X)
    return torch.cat([q_normalize(q_conj(q)), 1.0 / s], -1)


def rxso3_mul(X, Y):
    qx, sx = rxso3_split(X)
    qy, sy = rxso3_split(Y)
    return torch.cat([q_normalize(q_mul_raw(qx, qy)), sx * sy], -1)


def rxso3_act(X, p):
    q, s = rxso3_split(X)
    return s * q_rot(q, p)


def rxso3_Adj_matrix(X):
    """rxso3.h:68-72: the scale commutes, so only the rotation acts on the tangent"""
    q, _ = rxso3_split(X)
    A = torch.zeros(X.shape[:-1] + (4, 4), dtype=X.dtype, device=X.device)
    A[..., :3, :3] = q_matrix(q)
    A[..., 3, 3] = 1
    return A


def rxso3_adj_small(a):
    """rxso3.h:122-131"""
    A = torch.zeros(a.shape[:-1] + (4, 4), dtype=a.dtype, device=a.device)
    A[..., :3, :3] = hat(a[..., :3])
    return A


def rxso3_matrix(X):
    q, s = rxso3_split(X)
    T = torch.zeros(X.shape[:-1] + (4, 4), dtype=X.dtype, device=X.device)
    T[..., :3, :3] = s[..., None] * q_matrix(q)
    T[..., 3, 3] = 1
    return T


def rxso3_projector(X):
    """rxso3.h:85-100 (5x5)"""
    q, s = rxso3_split(X)
    J = torch.zeros(X.shape[:-1] + (5, 5), dtype=X.dtype, device=X.device)
    J[..., :4, :4] = so3_projector(q)
    J[..., 3, 3] = 0                     # the SO3 block only fills rows 0..3 x cols 0..2
    J[..., 4, 3] = s[..., 0]
    return J


def _blockdiag_so3(fn, a):
    J = torch.zeros(a.shape[:-1] + (4, 4), dtype=a.dtype, device=a.device)
    J[..., :3, :3] = fn(a[..., :3])
    J[..., 3, 3] = 1
    return J


def rxso3_calcW(a):
    """rxso3.h:189-234: W(phi, sigma) = A Phi + B Phi^2 + C I, with the small-angle / small-scale branches"""
    phi, sigma = a[..., :3], a[..., 3]
    th = phi.norm(dim=-1)
    Phi = hat(phi)
    Phi2 = Phi @ Phi
    sc = torch.exp(sigma)
    one = torch.ones_like(th)
    s_small, t_small = sigma.abs() < EPS, th.abs() < EPS
    ths = torch.where(t_small, one, th)
    sgs = torch.where(s_small, one, sigma)
    t2 = ths * ths
    # sigma ~ 0
    A0 = torch.where(t_small, 0.5 * one, (1.0 - torch.cos(ths)) / t2)
    B0 = torch.where(t_small, one / 6.0, (ths - torch.sin(ths)) / (t2 * ths))
    # general sigma
    C1 = (sc - 1.0) / sgs
    s2 = sgs * sgs
    A1t = ((sgs - 1.0) * sc + 1.0) / s2
    B1t = (sc * 0.5 * s2 + sc - 1.0 - sgs * sc) / (s2 * sgs)
    sa, sb = sc * torch.sin(ths), sc * torch.cos(ths)
    c = t2 + s2
    A1g = (sa * sgs + (1.0 - sb) * ths) / (ths * c)
    B1g = (C1 - ((sb - 1.0) * sgs + sa * ths) / c) / t2
    A = torch.where(s_small, A0, torch.where(t_small, A1t, A1g))
    B = torch.where(s_small, B0, torch.where(t_small, B1t, B1g))
    C = torch.where(s_small, one, C1)
    I = torch.eye(3, dtype=a.dtype, device=a.device)
    return A[..., None, None] * Phi + B[..., None, None] * Phi2 + C[..., None, None] * I


class _RxSO3:
    N, K = 5, 4
    exp = staticmethod(rxso3_exp)
    log = staticmethod(rxso3_log)
    inv = staticmethod(rxso3_inv)
    mul = staticmethod(rxso3_mul)
    act = staticmethod(rxso3_act)
    act4 = staticmethod(lambda X, p: torch.cat([rxso3_act(X, p[..., :3]), p[..., 3:]], -1))
    adj = staticmethod(lambda X, a: matv(rxso3_Adj_matrix(X), a))
    adjT = staticmethod(lambda X, a: rowm(a, rxso3_Adj_matrix(X)))
    Adj_matrix = staticmethod(rxso3_Adj_matrix)
    adj_small = staticmethod(rxso3_adj_small)
    left_jacobian = staticmethod(lambda a: _blockdiag_so3(so3_left_jacobian, a))            # rxso3.h:293-298
    left_jacobian_inverse = staticmethod(lambda a: _blockdiag_so3(so3_left_jacobian_inverse, a))
    projector = staticmethod(rxso3_projector)
    matrix = staticmethod(rxso3_matrix)

    @staticmethod
    def act_jacobian(p):          # rxso3.h:307-311
        return torch.cat([hat(-p), p[..., None]], -1)

    @staticmethod
    def act4_jacobian(p):         # rxso3.h:313-318
        J = torch.zeros(p.shape[:-1] + (4, 4), dtype=p.dtype, device=p.device)
        J[..., :3, :3] = hat(-p[..., :3])
        J[..., :3, 3] = p[..., :3]
        return J


# ----------------------------------------------------------------------------------- Sim3
# data [tx ty tz qx qy qz qw s], tangent [tau, phi, sigma]   (sim3.h:12-217)
def sim3_split(X):
    return X[..., :3], X[..., 3:8]


def sim3_exp(a):
    """sim3.h:160-167: t = W(phi, sigma) tau"""
    return torch.cat([matv(rxso3_calcW(a[..., 3:7]), a[..., :3]), rxso3_exp(a[..., 3:7])], -1)


def sim3_log(X):
    """sim3.h:151-158: tau = W^-1 t (a general 3x3 inverse, as in the reference)"""
    t, R = sim3_split(X)
    ps = rxso3_log(R)
    return torch.cat([matv(torch.linalg.inv(rxso3_calcW(ps)), t), ps], -1)


def sim3_inv(X):
    t, R = sim3_split(X)
    Ri = rxso3_inv(R)
    return torch.cat([-rxso3_act(Ri, t), Ri], -1)


def sim3_mul(X, Y):
    tx, Rx = sim3_split(X)
    ty, Ry = sim3_split(Y)
    return torch.cat([tx + rxso3_act(Rx, ty), rxso3_mul(Rx, Ry)], -1)


def sim3_act(X, p):
    t, R = sim3_split(X)
    return rxso3_act(R, p) + t


def sim3_act4(X, p):
    t, R = sim3_split(X)
    return torch.cat([rxso3_act(R, p[..., :3]) + p[..., 3:] * t, p[..., 3:]], -1)


def sim3_matrix(X):
    t, R = sim3_split(X)
    T = rxso3_matrix(R)
    T[..., :3, 3] = t
    return T


def sim3_Adj_matrix(X):
    """sim3.h:98-110"""
    t, R = sim3_split(X)
    q, s = rxso3_split(R)
    Rm = q_matrix(q)
    A = torch.zeros(X.shape[:-1] + (7, 7), dtype=X.dtype, device=X.device)
    A[..., :3, :3] = s[..., None] * Rm
    A[..., :3, 3:6] = hat(t) @ Rm
    A[..., :3, 6] = -t
    A[..., 3:6, 3:6] = Rm
    A[..., 6, 6] = 1
    return A


def sim3_adj_small(a):
    """sim3.h:133-149"""
    tau, phi, sigma = a[..., :3], a[..., 3:6], a[..., 6]
    A = torch.zeros(a.shape[:-1] + (7, 7), dtype=a.dtype, device=a.device)
    I = torch.eye(3, dtype=a.dtype, device=a.device)
    A[..., :3, :3] = hat(phi) + sigma[..., None, None] * I
    A[..., :3, 3:6] = hat(tau)
    A[..., :3, 6] = -tau
    A[..., 3:6, 3:6] = hat(phi)
    return A


def sim3_left_jacobian(a):
    """sim3.h:169-180.  The reference's series stops at the Xi^4 / 120 term: the `+ Xi Xi^4 / 720` line follows a
    semicolon and is dead code -- restated as it behaves."""
    Xi = sim3_adj_small(a)
    Xi2 = Xi @ Xi
    Xi4 = Xi2 @ Xi2
    I = torch.eye(7, dtype=a.dtype, device=a.device)
    return I + Xi / 2.0 + Xi2 / 6.0 + (Xi @ Xi2) / 24.0 + Xi4 / 120.0


def sim3_left_jacobian_inverse(a):
    """sim3.h:182-191"""
    Xi = sim3_adj_small(a)
    Xi2 = Xi @ Xi
    Xi4 = Xi2 @ Xi2
    I = torch.eye(7, dtype=a.dtype, device=a.device)
    return I - Xi / 2.0 + Xi2 / 12.0 - Xi4 / 720.0


def sim3_projector(X):
    """sim3.h:88-96 (8x8)"""
    t, R = sim3_split(X)
    J = torch.zeros(X.shape[:-1] + (8, 8), dtype=X.dtype, device=X.device)
    J[..., :3, :3] = torch.eye(3, dtype=X.dtype, device=X.device)
    J[..., :3, 3:6] = hat(-t)
    J[..., :3, 6] = t
    J[..., 3:, 3:] = rxso3_projector(R)
    return J


class _Sim3:
    N, K = 8, 7
    exp = staticmethod(sim3_exp)
    log = staticmethod(sim3_log)
    inv = staticmethod(sim3_inv)
    mul = staticmethod(sim3_mul)
    act = staticmethod(sim3_act)
    act4 = staticmethod(sim3_act4)
    adj = staticmethod(lambda X, a: matv(sim3_Adj_matrix(X), a))
    adjT = staticmethod(lambda X, a: rowm(a, sim3_Adj_matrix(X)))
    Adj_matrix = staticmethod(sim3_Adj_matrix)
    adj_small = staticmethod(sim3_adj_small)
    left_jacobian = staticmethod(sim3_left_jacobian)
    left_jacobian_inverse = staticmethod(sim3_left_jacobian_inverse)
    projector = staticmethod(sim3_projector)
    matrix = staticmethod(sim3_matrix)

    @staticmethod
    def act_jacobian(p):          # sim3.h:193-199
        I = torch.eye(3, dtype=p.dtype, device=p.device).expand(p.shape[:-1] + (3, 3))
        return torch.cat([I, hat(-p), p[..., None]], -1)

    @staticmethod
    def act4_jacobian(p):         # sim3.h:201-207
        J = torch.zeros(p.shape[:-1] + (4, 7), dtype=p.dtype, device=p.device)
        J[..., :3, :3] = p[..., 3:, None] * torch.eye(3, dtype=p.dtype, device=p.device)
        J[..., :3, 3:6] = hat(-p[..., :3])
        J[..., :3, 6] = p[..., :3]
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


GROUPS = {1: _SO3, 2: _RxSO3, 3: _SE3, 4: _Sim3}


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
