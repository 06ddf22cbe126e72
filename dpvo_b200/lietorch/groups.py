"""Lie-group tensor wrappers with autograd over the `lietorch_backends` kernels.

Interface parity: dpvo/lietorch/groups.py:51-322 (LieGroup, SO3, SE3, cat, stack),
group_ops.py:7-102 (autograd functions incl. ToVec / FromVec), broadcasting.py:9-31.
Differences by design: operands are broadcast with expand (no physical repeat unless the kernel
needs a contiguous copy); SO3 / SE3 (the groups on the DPVO hot path) run hand-specialised kernels, RxSO3 / Sim3
the generic small-matrix operators of csrc/lie_scaled.cuh.
"""
import numpy as np
import torch

_B = None   # the lietorch_backends module (tests may monkeypatch an oracle stand-in)


def backend():
    global _B
    if _B is None:
        from .. import extensions
        _B = extensions()[2]
    return _B


def _flatten_pair(x, y):
    """Broadcast the leading dims of x and y; return contiguous [n, dx], [n, dy] and the batch shape."""
    if y is None:
        return (x.reshape(-1, x.shape[-1]).contiguous(),), tuple(x.shape[:-1])
    if x.dim() != y.dim():
        raise ValueError("lie op operands must have the same number of dims (%d vs %d)" % (x.dim(), y.dim()))
    lead = torch.broadcast_shapes(x.shape[:-1], y.shape[:-1])
    xs = x.expand(lead + x.shape[-1:]).reshape(-1, x.shape[-1]).contiguous()
    ys = y.expand(lead + y.shape[-1:]).reshape(-1, y.shape[-1]).contiguous()
    return (xs, ys), tuple(lead)


def _make_op(fwd, bwd):
    """autograd.Function over one backend op; gradients follow lietorch's left-tangent convention."""

    class _Op(torch.autograd.Function):
        @staticmethod
        def forward(ctx, gid, *inputs):
            ctx.gid = gid
            ctx.save_for_backward(*inputs)
            return getattr(backend(), fwd)(gid, *inputs)

        @staticmethod
        def backward(ctx, grad):
            if bwd is None:
                raise RuntimeError("backward of %s is not defined (as in the reference)" % fwd)
            outs = getattr(backend(), bwd)(ctx.gid, grad.contiguous(), *ctx.saved_tensors)
            return (None,) + tuple(outs)

    _Op.__name__ = "Lie_" + fwd
    return _Op


Exp = _make_op("expm", "expm_backward")
Log = _make_op("logm", "logm_backward")
Inv = _make_op("inv", "inv_backward")
Mul = _make_op("mul", "mul_backward")
Adj = _make_op("adj", "adj_backward")
AdjT = _make_op("adjT", "adjT_backward")
Act3 = _make_op("act", "act_backward")
Act4 = _make_op("act4", "act4_backward")
Jinv = _make_op("Jinv", None)
ToMatrix = _make_op("as_matrix", None)


class _VecCast(torch.autograd.Function):
    """group element <-> plain vector (group_ops.py:69-102): identity forward, projector backward."""

    @staticmethod
    def forward(ctx, gid, to_vec, x):
        ctx.gid, ctx.to_vec = gid, to_vec
        ctx.save_for_backward(x)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad):
        x, = ctx.saved_tensors
        J = backend().projector(ctx.gid, x)
        if not ctx.to_vec:
            J = torch.linalg.pinv(J)
        return None, None, torch.matmul(grad.unsqueeze(-2), J).squeeze(-2)


class LieGroup:
    group_name, group_id, manifold_dim, embedded_dim, id_elem = None, None, None, None, None

    def __init__(self, data):
        self.data = data

    def __repr__(self):
        return "%s: size=%s, device=%s, dtype=%s" % (self.group_name, tuple(self.shape), self.device, self.dtype)

    # ---- tensor-like properties
    shape = property(lambda self: self.data.shape[:-1])
    device = property(lambda self: self.data.device)
    dtype = property(lambda self: self.data.dtype)
    tangent_shape = property(lambda self: self.data.shape[:-1] + (self.manifold_dim,))

    # ---- construction
    @classmethod
    def Identity(cls, *batch_shape, **kwargs):
        if len(batch_shape) == 1 and isinstance(batch_shape[0], (tuple, list, torch.Size)):
            batch_shape = tuple(batch_shape[0])
        e = cls.id_elem.to(device=kwargs.get("device", None), dtype=kwargs.get("dtype", cls.id_elem.dtype))
        return cls(e.expand(tuple(batch_shape) + (cls.embedded_dim,)).contiguous())

    @classmethod
    def IdentityLike(cls, G):
        return cls.Identity(G.shape, device=G.data.device, dtype=G.data.dtype)

    @classmethod
    def InitFromVec(cls, data):
        return cls(cls._apply(lambda gid, x: _VecCast.apply(gid, False, x), data))

    @classmethod
    def Random(cls, *batch_shape, sigma=1.0, **kwargs):
        if len(batch_shape) == 1 and isinstance(batch_shape[0], (tuple, list, torch.Size)):
            batch_shape = tuple(batch_shape[0])
        return cls.exp(sigma * torch.randn(tuple(batch_shape) + (cls.manifold_dim,), **kwargs))

    # ---- op plumbing
    @classmethod
    def _apply(cls, fn, x, y=None):
        flat, lead = _flatten_pair(x, y)
        out = fn(cls.group_id, *flat)
        return out.view(lead + out.shape[1:])

    @classmethod
    def apply_op(cls, op, x, y=None):
        return cls._apply(op.apply, x, y)

    @classmethod
    def exp(cls, x):
        return cls(cls.apply_op(Exp, x))

    def log(self):
        return self.apply_op(Log, self.data)

    def inv(self):
        return self.__class__(self.apply_op(Inv, self.data))

    def mul(self, other):
        return self.__class__(self.apply_op(Mul, self.data, other.data))

    def retr(self, a):
        """Exp(a) * X (groups.py:153-156)"""
        return self.__class__(self.apply_op(Mul, self.apply_op(Exp, a), self.data))

    def adj(self, a):
        return self.apply_op(Adj, self.data, a)

    def adjT(self, a):
        return self.apply_op(AdjT, self.data, a)

    def Jinv(self, a):
        return self.apply_op(Jinv, self.data, a)

    def act(self, p):
        if p.shape[-1] == 3:
            return self.apply_op(Act3, self.data, p)
        if p.shape[-1] == 4:
            return self.apply_op(Act4, self.data, p)
        raise ValueError("points must have 3 or 4 components")

    def matrix(self):
        """4x4 matrices through act4 on the identity columns (groups.py:180-184), so it is differentiable."""
        I = torch.eye(4, dtype=self.dtype, device=self.device).view((1,) * (self.data.dim() - 1) + (4, 4))
        return self.__class__(self.data[..., None, :]).act(I).transpose(-1, -2)

    def translation(self):
        p = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=self.dtype, device=self.device)
        return self.apply_op(Act4, self.data, p.view((1,) * (self.data.dim() - 1) + (4,)))

    def vec(self):
        return self._apply(lambda gid, x: _VecCast.apply(gid, True, x), self.data)

    # ---- container behaviour
    def detach(self):
        return self.__class__(self.data.detach())

    def view(self, dims):
        return self.__class__(self.data.view(tuple(dims) + (self.embedded_dim,)))

    def __mul__(self, other):
        if isinstance(other, LieGroup):
            return self.mul(other)
        if isinstance(other, torch.Tensor):
            return self.act(other)
        return NotImplemented

    def __getitem__(self, index):
        return self.__class__(self.data[index])

    def __setitem__(self, index, item):
        self.data[index] = item.data

    def to(self, *args, **kwargs):
        return self.__class__(self.data.to(*args, **kwargs))

    def cpu(self):
        return self.__class__(self.data.cpu())

    def cuda(self):
        return self.__class__(self.data.cuda())

    def float(self, device=None):
        return self.__class__(self.data.float())

    def double(self, device=None):
        return self.__class__(self.data.double())

    def unbind(self, dim=0):
        return [self.__class__(x) for x in self.data.unbind(dim=dim)]


class SO3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "SO3", 1, 3, 4
    id_elem = torch.tensor([0.0, 0.0, 0.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SE3):
            data = data.data[..., 3:7]
        super().__init__(data)


class RxSO3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "RxSO3", 2, 4, 5
    id_elem = torch.tensor([0.0, 0.0, 0.0, 1.0, 1.0])

    def __init__(self, data):
        if isinstance(data, Sim3):                       # rotation + scale part of a similarity (groups.py:259-262)
            data = data.data[..., 3:8]
        super().__init__(data)


class SE3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "SE3", 3, 6, 7
    id_elem = torch.tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])

    def __init__(self, data):
        if isinstance(data, SO3):
            data = torch.cat([torch.zeros_like(data.data[..., :3]), data.data], -1)
        super().__init__(data)

    def scale(self, s):
        t, q = self.data.split([3, 4], -1)
        return SE3(torch.cat([t * s.unsqueeze(-1), q], dim=-1))


class Sim3(LieGroup):
    group_name, group_id, manifold_dim, embedded_dim = "Sim3", 4, 7, 8
    id_elem = torch.tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0])

    def __init__(self, data):
        # embeddings of the smaller groups (groups.py:297-311): unit scale, and zero translation for a pure rotation
        if isinstance(data, SO3):
            q = data.data
            data = torch.cat([torch.zeros_like(q[..., :3]), q, torch.ones_like(q[..., :1])], -1)
        elif isinstance(data, SE3):
            data = torch.cat([data.data, torch.ones_like(data.data[..., :1])], -1)
        elif isinstance(data, Sim3):
            data = data.data
        super().__init__(data)


class LieGroupParameter(torch.Tensor):
    """Tangent-space parameter around a fixed group element (groups.py:10-48)."""
    from torch._C import _disabled_torch_function_impl
    __torch_function__ = _disabled_torch_function_impl

    def __new__(cls, group, requires_grad=True):
        data = torch.zeros(group.tangent_shape, device=group.data.device, dtype=group.data.dtype, requires_grad=True)
        return torch.Tensor._make_subclass(cls, data, requires_grad)

    def __init__(self, group):
        self.group = group

    def retr(self):
        return self.group.retr(self)

    def log(self):
        return self.retr().log()

    def inv(self):
        return self.retr().inv()

    def adj(self, a):
        return self.retr().adj(a)

    def __mul__(self, other):
        if isinstance(other, LieGroupParameter):
            return self.retr() * other.retr()
        return self.retr() * other

    def add_(self, update, alpha):
        self.group = self.group.exp(alpha * update) * self.group

    def __getitem__(self, index):
        return self.retr().__getitem__(index)


def cat(group_list, dim):
    return group_list[0].__class__(torch.cat([X.data for X in group_list], dim=dim))


def stack(group_list, dim):
    return group_list[0].__class__(torch.stack([X.data for X in group_list], dim=dim))
