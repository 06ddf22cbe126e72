"""ORACLE (test infrastructure -- never imported by the product path).

CPU restatement of the index bookkeeping on the update path:
    neighbors      cuda_ba.neighbors, dpvo/fastba/ba.cpp:59-97: group edges by ii, std::stable_sort
                   each group by jj, link predecessor / successor (-1 at the ends)
    group_edges    the partition torch.unique(key, return_inverse=True) yields in SoftAgg
                   (dpvo/blocks.py:41) and torch::_unique(kk) in ba_cuda.cu:447, as a CSR
Integer work: results must match the device implementation bit for bit.
Pinning: ba.cpp's neighbors is compiled from /root/reference into oracle/_ref (it is plain C++ with a
CPU sort) and compared on the GPU box; here the restatement is checked against a brute-force
definition (tests/test_oracle_graph.py).
"""
import numpy as np
import torch


def neighbors(ii, jj):
    ii = np.asarray(ii.cpu() if torch.is_tensor(ii) else ii, dtype=np.int64)
    jj = np.asarray(jj.cpu() if torch.is_tensor(jj) else jj, dtype=np.int64)
    E = len(ii)
    ix = np.full(E, -1, dtype=np.int64)
    jx = np.full(E, -1, dtype=np.int64)
    # stable sort by (ii, jj) keeps original order inside equal (ii, jj), as std::stable_sort does
    order = np.lexsort((np.arange(E), jj, ii))
    for a, b in zip(order[:-1], order[1:]):
        if ii[a] == ii[b]:
            jx[a] = b
            ix[b] = a
    return torch.from_numpy(ix), torch.from_numpy(jx)


def group_edges(key_a, key_b=None, sec=None):
    """Returns dict(order, group_of, group_start, key_a, key_b, n) with the semantics of
    include/dpvo_b200.h:dpvo_group_edges."""
    ka = np.asarray(key_a.cpu() if torch.is_tensor(key_a) else key_a, dtype=np.int64)
    E = len(ka)
    kb = np.zeros(E, dtype=np.int64) if key_b is None else np.asarray(key_b.cpu() if torch.is_tensor(key_b) else key_b, dtype=np.int64)
    sc = np.zeros(E, dtype=np.int64) if sec is None else np.asarray(sec.cpu() if torch.is_tensor(sec) else sec, dtype=np.int64)
    order = np.lexsort((np.arange(E), sc, kb, ka)).astype(np.int32)
    if E == 0:
        return dict(order=order, group_of=np.zeros(0, np.int32), group_start=np.zeros(1, np.int32),
                    key_a=np.zeros(0, np.int64), key_b=np.zeros(0, np.int64), n=0)
    sa, sb = ka[order], kb[order]
    head = np.ones(E, dtype=bool)
    head[1:] = (sa[1:] != sa[:-1]) | (sb[1:] != sb[:-1])
    gid_sorted = np.cumsum(head) - 1
    group_of = np.empty(E, dtype=np.int32)
    group_of[order] = gid_sorted
    starts = np.flatnonzero(head).astype(np.int32)
    return dict(order=order, group_of=group_of, group_start=np.concatenate([starts, [E]]).astype(np.int32),
                key_a=sa[head], key_b=sb[head], n=int(head.sum()))
