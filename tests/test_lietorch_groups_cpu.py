"""Constructors that embed one Lie group in another (dpvo/lietorch/groups.py:243-311): pure tensor bookkeeping, checked on
the CPU (no kernel runs: group operations themselves need the CUDA extension and are covered by the GPU tests)."""
import torch


def test_group_embeddings():
    from dpvo_b200.lietorch.groups import SO3, RxSO3, SE3, Sim3
    g = torch.Generator().manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(2, 5, 4, generator=g), dim=-1)
    t = torch.randn(2, 5, 3, generator=g)
    s = torch.rand(2, 5, 1, generator=g) + 0.5
    R, T, S = SO3(q), SE3(torch.cat([t, q], -1)), Sim3(torch.cat([t, q, s], -1))
    assert torch.equal(SE3(R).data, torch.cat([torch.zeros_like(t), q], -1))
    assert torch.equal(SO3(T).data, q)
    assert torch.equal(Sim3(T).data, torch.cat([t, q, torch.ones_like(s)], -1))
    assert torch.equal(Sim3(R).data, torch.cat([torch.zeros_like(t), q, torch.ones_like(s)], -1))
    assert torch.equal(Sim3(S).data, S.data)
    assert torch.equal(RxSO3(S).data, torch.cat([q, s], -1))
    for X, dim in ((R, 4), (T, 7), (S, 8), (RxSO3(S), 5)):
        assert X.data.shape[-1] == dim and X.shape == (2, 5)
