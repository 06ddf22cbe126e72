"""Device-resident patch-graph bookkeeping (csrc/pgraph.cu, dpvo_b200/patchgraph.py) against a restatement of the
reference's growing / shrinking tensors (dpvo/dpvo.py:215-238 append_factors / remove_factors, :279-286 keyframe
renumbering, :362-375 edge rules), and the update step on the fixed-capacity store: parked slots do not touch the
results of the active edges, and one captured CUDA graph keeps working while the topology changes."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dev_scalar(v, dtype=torch.long):
    return torch.tensor([v], dtype=dtype, device=DEV)


class RefGraph:
    """the reference's bookkeeping on variable-length tensors"""

    def __init__(self, M):
        self.M = M
        z = torch.zeros(0, dtype=torch.long, device=DEV)
        self.ii, self.jj, self.kk = z, z.clone(), z.clone()
        self.net = torch.zeros(0, 384, device=DEV)

    def append(self, ii, jj, kk):                      # dpvo.py:215-222
        self.ii, self.jj, self.kk = torch.cat([self.ii, ii]), torch.cat([self.jj, jj]), torch.cat([self.kk, kk])
        self.net = torch.cat([self.net, torch.zeros(len(ii), 384, device=DEV)])

    def remove(self, m):                               # dpvo.py:224-238
        self.ii, self.jj, self.kk, self.net = self.ii[~m], self.jj[~m], self.kk[~m], self.net[~m]

    def keyframe(self, k):                             # dpvo.py:279-286
        self.remove((self.ii == k) | (self.jj == k))
        self.kk[self.ii > k] -= self.M
        self.ii[self.ii > k] -= 1
        self.jj[self.jj > k] -= 1

    def frame_edges(self, n, r):                       # dpvo.py:362-375
        M = self.M
        kf, jf = torch.meshgrid(torch.arange(M * max(n - r, 0), M * max(n - 1, 0), device=DEV), torch.arange(n - 1, n, device=DEV), indexing="ij")
        kb, jb = torch.meshgrid(torch.arange(M * max(n - 1, 0), M * n, device=DEV), torch.arange(max(n - r, 0), n, device=DEV), indexing="ij")
        kk = torch.cat([kf.reshape(-1), kb.reshape(-1)])
        jj = torch.cat([jf.reshape(-1), jb.reshape(-1)])
        return kk // M, jj, kk


def _sorted_edges(ii, jj, kk, extra=None):
    key = (kk * 4096 + jj) * 4096 + ii
    o = torch.argsort(key)
    return (ii[o], jj[o], kk[o]) if extra is None else (ii[o], jj[o], kk[o], extra[o])


def test_bookkeeping_matches_reference_semantics():
    from dpvo_b200.patchgraph import DevicePatchGraph
    M, r, window = 8, 4, 6
    pg = DevicePatchGraph(capacity=1000, M=M, dummy_frame=63)
    ref = RefGraph(M)
    gen = torch.Generator(device=DEV).manual_seed(0)
    n = 0
    for frame in range(14):
        n += 1
        ii, jj, kk = ref.frame_edges(n, r)
        ref.append(ii, jj, kk)
        if n >= r:                                     # steady state: the device generates the same lists from the device counter
            di, dj, dk = __import__("dpvo_b200").extensions()[3].pgraph_new_edges(_dev_scalar(n), M, r)
            assert torch.equal(di, ii) and torch.equal(dj, jj) and torch.equal(dk, kk)
        slots = pg.append(ii, jj, kk)
        assert (slots >= 0).all()
        # give every active row a recognisable state: value = f(edge) -- stays with the edge through removals
        rows = pg.net_rows()
        tag = (pg.kk * 131 + pg.jj * 7 + 1).float()
        rows[pg.active.bool()] = tag[pg.active.bool()][:, None].expand(-1, 384) * 1e-3
        ref.net = ((ref.kk * 131 + ref.jj * 7 + 1).float() * 1e-3)[:, None].expand(-1, 384).contiguous()
        if frame in (7, 10):                           # keyframe removal of frame k = n - 3 (dpvo.py:279-299)
            k = n - 3
            ref.keyframe(k)
            pg.remove_frame(_dev_scalar(k))
            n -= 1
        ref.remove((ref.kk // M) < n - window)         # dpvo.py:300-306
        pg.remove_old(_dev_scalar(n), window)
        a = _sorted_edges(*pg.edges())
        b = _sorted_edges(ref.ii, ref.jj, ref.kk)
        assert all(torch.equal(x, y) for x, y in zip(a, b)), "active edge set differs from the reference's lists at frame %d" % frame
        assert int(pg.n_active) == ref.ii.numel()
        # the state rows travelled with their edges
        tags = pg.net_rows()[pg.active.bool()][:, 0]
        assert torch.allclose(torch.sort(tags)[0], torch.sort(ref.net[:, 0])[0])
        # parked slots hold the dummy edge
        parked = ~pg.active.bool()
        assert (pg.ii[parked] == 63).all() and (pg.jj[parked] == 63).all() and (pg.kk[parked] == 63 * M).all()
    assert int(pg.overflow) == 0
    # a disabled call is a no-op; a full store reports overflow and drops the surplus
    before = [t.clone() for t in (pg.ii, pg.jj, pg.kk, pg.active)]
    off = _dev_scalar(0, torch.int32)
    pg.remove_old(_dev_scalar(10 ** 6), window, enable=off)
    pg.append(ii, jj, kk, enable=off)
    assert all(torch.equal(x, y) for x, y in zip(before, (pg.ii, pg.jj, pg.kk, pg.active)))
    big = torch.zeros(2000, dtype=torch.long, device=DEV)
    slots = pg.append(big, big, big)
    assert int(pg.overflow) == 1 and int(pg.n_active) == pg.cap and (slots[-1] == -1)


def test_new_rows_start_at_zero_and_kept_rows_are_untouched():
    from dpvo_b200.patchgraph import DevicePatchGraph
    pg = DevicePatchGraph(capacity=512, M=4, dummy_frame=31)
    e = torch.arange(300, device=DEV)
    pg.append(e // 4 // 3, e % 5, e // 3)
    rows = torch.randn(pg.cap, 384, device=DEV)
    pg.net[0].copy_(rows)
    pg.remove_old(_dev_scalar(20), 10)                 # parks the edges of patches in frames < 10
    kept = pg.active.bool().clone()
    slots = pg.append(torch.full((50,), 20, device=DEV), torch.full((50,), 21, device=DEV), torch.arange(80, 130, device=DEV)).long()
    out = pg.net_rows()
    assert torch.equal(out[kept], rows[kept])
    assert (out[slots] == 0).all()
    assert torch.equal(slots, torch.nonzero(~kept).flatten()[:50])          # parked slots are reused in index order


def test_update_on_the_store_ignores_parked_slots_and_survives_topology_changes():
    """One captured graph, three topologies: every replay equals eager launches on the same store bit for bit, and the
    active edges get the same update as a compact runner built from exactly the active edge list."""
    from dpvo_b200 import synthetic
    from dpvo_b200.runner import UpdateRunner
    from dpvo_b200.patchgraph import DevicePatchGraph
    st = synthetic.make_state("fast", 12, device=DEV, seed=3)
    M = st.cfg["M"]
    pg = DevicePatchGraph(capacity=st.E + 700, M=M, dummy_frame=st.poses.shape[0] - 1)
    pg.append(st.ii, st.jj, st.kk)
    run = UpdateRunner(st, graph=pg)
    poses0, patches0 = st.poses.clone(), st.patches.clone()

    def compact_result():
        """the same update on exact-length arrays holding the active edges in slot order, from the same state rows"""
        m = pg.active.bool()
        st2 = synthetic.make_state("fast", 12, device=DEV, seed=3)
        st2.ii, st2.jj, st2.kk = pg.ii[m].clone(), pg.jj[m].clone(), pg.kk[m].clone()
        st2.poses.copy_(poses0); st2.patches.copy_(patches0)
        r2 = UpdateRunner(st2, update=run.update)
        r2.net[0].copy_(state_rows[m])
        tgt, wgt = r2.step()
        torch.cuda.synchronize()
        return st2.poses.clone(), st2.patches.clone(), tgt[0], wgt[0], r2.net[0], m

    run.capture()
    frame = _dev_scalar(st.n)
    for round_ in range(3):
        if round_ == 1:                                # drop the edges of the oldest live frame's patches
            oldest = int((pg.kk[pg.active.bool()] // M).min())
            pg.remove_old(_dev_scalar(oldest + 1 + 16), 16)
        if round_ == 2:                                # and add edges again into the parked slots
            m_old = ~pg.active.bool()
            k_new = torch.arange(M * (st.n - 1), M * st.n, device=DEV).repeat_interleave(3)
            j_new = torch.arange(st.n - 3, st.n, device=DEV).repeat(M)
            pg.append(k_new // M, j_new, k_new)
        state_rows = pg.net_rows().clone()
        # eager launches
        st.poses.copy_(poses0); st.patches.copy_(patches0)
        tgt_e, wgt_e = run.step()
        torch.cuda.synchronize()
        eager = (st.poses.clone(), st.patches.clone(), tgt_e.clone(), wgt_e.clone(), pg.net_rows().clone())
        # graph replay from the same state
        pg.net[0].copy_(state_rows)
        st.poses.copy_(poses0); st.patches.copy_(patches0)
        tgt_g, wgt_g = run.step_graph()
        torch.cuda.synchronize()
        replay = (st.poses.clone(), st.patches.clone(), tgt_g.clone(), wgt_g.clone(), pg.net_rows().clone())
        for a, b in zip(eager, replay):
            assert torch.equal(a, b), "graph replay differs from eager launches in round %d" % round_
        # against the compact runner
        p2, q2, tgt2, wgt2, net2, m = compact_result()
        n = st.n
        assert (eager[0][:n] - p2[:n]).abs().max() <= 1e-4 * p2[:n].abs().max()
        dref = q2[:n * M, 2, 1, 1]
        assert (eager[1][:n * M, 2, 1, 1] - dref).abs().max() <= 1e-4 * dref.abs().max()
        assert torch.allclose(eager[2][0][m], tgt2, atol=2e-3) and torch.allclose(eager[3][0][m], wgt2, atol=1e-3)
        assert (eager[4][m] - net2).abs().max() <= 2e-3 * net2.abs().max()
        assert (eager[3][0][~m] == 0).all(), "parked edges must carry zero confidence"
        # carry on from the replayed state
    assert int(pg.overflow) == 0
