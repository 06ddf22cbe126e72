"""Device edge grouping / neighbors: integer work, bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from oracle import graph as OG
from dpvo_b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check_group(ext, ka, kb, sec):
    out = ext[3].group_edges(ka.to(DEV), None if kb is None else kb.to(DEV), None if sec is None else sec.to(DEV))
    order, gof, gstart, oka, okb, ng = [t.cpu().numpy() for t in out]
    ref = OG.group_edges(ka, kb, sec)
    G = int(ng[0])
    assert G == ref["n"]
    assert np.array_equal(order, ref["order"])
    assert np.array_equal(gof, ref["group_of"])
    assert np.array_equal(gstart[:G + 1], ref["group_start"])
    assert np.array_equal(oka[:G], ref["key_a"])
    if kb is not None:
        assert np.array_equal(okb[:G], ref["key_b"])


@pytest.mark.parametrize("E", [1, 31, 2048, 2049, 50000])
def test_group_edges_random(ext, E):
    g = torch.Generator().manual_seed(E)
    ka = torch.randint(-5, 40, (E,), generator=g)
    kb = torch.randint(100, 160, (E,), generator=g)
    sec = torch.randint(0, 3000, (E,), generator=g)
    _check_group(ext, ka, None, None)
    _check_group(ext, ka, kb, None)
    _check_group(ext, ka, kb, sec)
    _check_group(ext, ka * 100000007, None, sec)      # wide keys: several radix passes


@pytest.mark.parametrize("E", [65536, 65537, 150000])
def test_group_edges_both_kernels_at_the_size_boundary(ext, E):
    """E <= 65,536 runs inside one thread-block cluster (distributed shared memory), larger graphs on the persistent
    multi-CTA kernel: same results on either side of the switch"""
    g = torch.Generator().manual_seed(E)
    ka = torch.randint(0, 3000, (E,), generator=g)
    kb = torch.randint(0, 40, (E,), generator=g)
    sec = torch.randint(0, 40, (E,), generator=g)
    _check_group(ext, ka, None, sec)
    _check_group(ext, kb, sec, None)


@pytest.mark.parametrize("E", [9, 5000])
def test_group_edges_keys_wider_than_64_bits(ext, E):
    """three fields that do not fit one 64-bit composite key: sorted one field at a time"""
    g = torch.Generator().manual_seed(E)
    big = 1 << 40
    ka = torch.randint(-big, big, (E,), generator=g)
    ka[E // 2:] = ka[:E - E // 2].clone()              # duplicates so that groups have more than one member
    kb = torch.randint(-big, big, (E,), generator=g)
    kb[E // 2:] = kb[:E - E // 2].clone()
    sec = torch.randint(-big, big, (E,), generator=g)
    _check_group(ext, ka, kb, sec)
    _check_group(ext, ka, None, sec)


def test_group_edges_dpvo_graph(ext):
    ii, jj, kk = synthetic.replay_edges(36, 96, 13, 22)
    _check_group(ext, kk, None, jj)
    _check_group(ext, ii, jj, None)


@pytest.mark.parametrize("E", [1, 5, 4097, 47712])
def test_neighbors(ext, E):
    if E == 47712:
        ii, jj, kk = synthetic.replay_edges(36, 96, 13, 22)
        a, b = kk, jj
    else:
        g = torch.Generator().manual_seed(E)
        a = torch.randint(0, max(2, E // 7), (E,), generator=g)
        b = torch.randint(0, 9, (E,), generator=g)          # many ties: stable order matters
    ix, jx = ext[1].neighbors(a.to(DEV), b.to(DEV))
    rix, rjx = OG.neighbors(a, b)
    assert ix.dtype == torch.int64 and jx.dtype == torch.int64
    assert torch.equal(ix.cpu(), rix) and torch.equal(jx.cpu(), rjx)


def test_neighbors_empty(ext):
    e = torch.zeros(0, dtype=torch.long, device=DEV)
    ix, jx = ext[1].neighbors(e, e)
    assert ix.numel() == 0 and jx.numel() == 0


@pytest.mark.parametrize("E", [1, 777, 5000, 47712])
def test_group_edges_pair_equals_two_single_launches(ext, E):
    """both groupings of an update in one cooperative launch (grid.y = 2) are bit-identical to two launches"""
    g = torch.Generator().manual_seed(E)
    kk = torch.randint(0, 2300, (E,), generator=g).to(DEV)
    ii = torch.randint(0, 36, (E,), generator=g).to(DEV)
    jj = torch.randint(0, 36, (E,), generator=g).to(DEV)
    a = ext[3].group_edges(kk, None, jj)
    b = ext[3].group_edges(ii, jj, None)
    p = ext[3].group_edges_pair(kk, None, jj, ii, jj, None)
    na, nb = int(a[5]), int(b[5])
    assert int(p[5]) == na and int(p[11]) == nb
    for q, (one, n) in enumerate(((a, na), (b, nb))):
        o = p[6 * q:6 * q + 6]
        assert torch.equal(o[0], one[0]) and torch.equal(o[1], one[1])
        assert torch.equal(o[2][:n + 1], one[2][:n + 1]) and torch.equal(o[3][:n], one[3][:n])
    assert torch.equal(p[10][:nb], b[4][:nb])
