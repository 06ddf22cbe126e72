"""Update operator (dpvo/net.py:27-92) on the device vs the fp32 CPU oracle, plus its row kernels."""
import pytest
import torch
import torch.nn.functional as F

from oracle import update as OU, graph as OG
from dpvo_b200 import synthetic
from dpvo_b200.net import Update, EdgeGroups

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_add_layernorm_variants(ext):
    g = torch.Generator().manual_seed(61)
    a = torch.randn(1, 777, 384, generator=g)
    b = torch.randn(1, 777, 384, generator=g).half()
    c = torch.randn(1, 777, 384, generator=g).half()
    gamma, beta = torch.rand(384, generator=g) + 0.5, torch.randn(384, generator=g)
    ref = F.layer_norm(a.double() + b.double() + c.double(), (384,), gamma.double(), beta.double(), 1e-3)
    y32, y16 = ext[3].add_layernorm(a.to(DEV), b.to(DEV), c.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-3, False, True, True)
    assert (y32.cpu().double() - ref).abs().max().item() < 2e-5
    assert (y16.cpu().double() - ref).abs().max().item() < 4e-3
    _, r16 = ext[3].add_layernorm(a.half().to(DEV), None, None, gamma.to(DEV), beta.to(DEV), 1e-3, True, False, True)
    ref2 = F.relu(F.layer_norm(a.half().double(), (384,), gamma.double(), beta.double(), 1e-3))
    assert (r16.cpu().double() - ref2).abs().max().item() < 4e-3


def test_gather_residual_gate_heads(ext):
    g = torch.Generator().manual_seed(62)
    E = 1000
    x = torch.randn(1, E, 384, generator=g)
    idx = torch.randint(-1, E, (E,), generator=g)
    y = ext[3].gather_rows_masked(x.to(DEV), idx.to(DEV), True).cpu()
    ref = torch.where(idx[None, :, None] >= 0, x[:, idx.clamp(min=0)], torch.zeros(())).half()
    assert torch.equal(y, ref)
    net = torch.randn(1, E, 384, generator=g)
    u = torch.randn(1, 50, 384, generator=g).half()
    gof = torch.randint(0, 50, (E,), generator=g)
    nd = net.clone().to(DEV)
    n16 = ext[3].residual_add_(nd, u.to(DEV), gof.to(DEV).int(), True)
    ref = net + u.float()[:, gof]
    assert (nd.cpu() - ref).abs().max().item() < 1e-6 and torch.equal(n16.cpu(), ref.half())
    gate = torch.randn(1, E, 384, generator=g).half()
    res = torch.randn(1, E, 384, generator=g).half()
    y32, r16 = ext[3].gated_residual(x.to(DEV), gate.to(DEV), res.to(DEV), True)
    ref = x + torch.sigmoid(gate.float()) * res.float()
    assert (y32.cpu() - ref).abs().max().item() < 1e-5
    assert (r16.cpu().float() - F.relu(ref)).abs().max().item() < 4e-3
    W4, b4 = torch.randn(4, 384, generator=g) / 20, torch.randn(4, generator=g)
    d, w = ext[3].update_heads(x.to(DEV), W4.to(DEV), b4.to(DEV))
    o = F.relu(x.double()) @ W4.double().t() + b4.double()
    assert (d.cpu().double() - o[..., :2]).abs().max().item() < 1e-5
    assert (w.cpu().double() - torch.sigmoid(o[..., 2:])).abs().max().item() < 1e-5


def test_gated_residual_folded_into_layernorm_and_heads(ext):
    """x + gate * res applied by its consumer: as a scaled operand of add_layernorm (in place, strided gate view)
    and as the gated input of the heads (which also writes the sum back)"""
    g = torch.Generator().manual_seed(77)
    E, D = 1500, 384
    x = torch.randn(1, E, D, generator=g)
    ga = torch.rand(1, E, 2 * D, generator=g).half()          # [gate | other half], gate is a strided view
    r2 = torch.randn(1, E, D, generator=g).half()
    gamma, beta = torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g) * 0.1
    s = x.double() + ga[..., :D].double() * r2.double()
    ref = F.layer_norm(s, (D,), gamma.double(), beta.double(), 1e-3)
    xd, gad = x.clone().to(DEV), ga.to(DEV)
    y32, y16 = ext[3].add_layernorm(xd, None, r2.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-3, False, True, True, None, True, gad[..., :D])
    assert y32.data_ptr() == xd.data_ptr()
    assert (y32.cpu().double() - ref).abs().max().item() < 1e-4
    assert (y16.cpu().double() - ref).abs().max().item() < 4e-3
    W4, b4 = torch.randn(4, D, generator=g) / 20, torch.randn(4, generator=g)
    xd = x.clone().to(DEV)
    d, w = ext[3].update_heads(xd, W4.to(DEV), b4.to(DEV), None, gad[..., :D], r2.to(DEV))
    assert (xd.cpu().double() - s).abs().max().item() < 1e-5               # the sum was written back
    o = F.relu(s) @ W4.double().t() + b4.double()
    assert (d.cpu().double() - o[..., :2]).abs().max().item() < 1e-4
    assert (w.cpu().double() - torch.sigmoid(o[..., 2:])).abs().max().item() < 1e-4


def test_softagg_reduce_matches_scatter_softmax(ext):
    g = torch.Generator().manual_seed(63)
    E, G = 5000, 137
    fg = torch.randn(1, E, 768, generator=g).half()
    key = torch.randint(0, G, (E,), generator=g)
    grp = EdgeGroups(key.to(DEV))
    y = ext[3].softagg_reduce(fg.to(DEV), grp.order, grp.group_start, grp.n, grp.max_groups).cpu().double()
    f, gl = fg[..., :384].double(), fg[..., 384:].double()
    _, inv = torch.unique(key, return_inverse=True)
    w = OU.scatter_softmax(gl, inv, dim=1)
    ref = OU.scatter_sum(f * w, inv, dim=1)
    assert y.shape[1] == ref.shape[1]
    assert (y - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("config,n_frames", [("fast", 14), ("default", 36)])
def test_update_forward_matches_oracle(ext, config, n_frames):
    """fp16 tensor-core GEMMs with fp32 state vs the all-fp32 oracle.  Tolerance is set by the
    reference's own mixed precision: the oracle run under autocast (what dpvo.py:332 does) is measured
    against the same fp32 result and we must be no worse than 2x that error."""
    st = synthetic.make_state(config, n_frames, device="cpu", features=False)
    E = st.E
    torch.manual_seed(1234)
    ref_mod = OU.Update(3).eval()
    ours = Update(3).eval()
    ours.load_state_dict(ref_mod.state_dict())
    ours = ours.to(DEV)
    g = torch.Generator().manual_seed(64)
    net = (torch.randn(1, E, 384, generator=g) * 0.5)
    inp = (torch.randn(1, E, 384, generator=g) * 0.25).half()
    corr = (torch.randn(1, E, 882, generator=g) * 2).half()
    with torch.no_grad():
        rn, (rd, rw, _) = ref_mod(net, inp.float(), corr.float(), None, st.ii, st.jj, st.kk)
        on, (od, ow, _) = ours(net.to(DEV), inp.to(DEV), corr.to(DEV), None, st.ii.to(DEV), st.jj.to(DEV), st.kk.to(DEV))
        mod16 = OU.Update(3).eval()
        mod16.load_state_dict(ref_mod.state_dict())
        mod16 = mod16.to(DEV)
        with torch.autocast("cuda", dtype=torch.half):
            an, (ad, aw, _) = mod16(net.to(DEV), inp.to(DEV), corr.to(DEV), None, st.ii.to(DEV), st.jj.to(DEV), st.kk.to(DEV))

    def err(a, b):
        return (a.float().cpu() - b).abs().max().item()

    e_net, e_d, e_w = err(on, rn), err(od, rd), err(ow, rw)
    a_net, a_d, a_w = err(an, rn), err(ad, rd), err(aw, rw)
    print("ours  vs fp32:", e_net, e_d, e_w, " autocast vs fp32:", a_net, a_d, a_w, " scale:", rn.abs().max().item())
    assert e_net <= max(2 * a_net, 2e-2) and e_d <= max(2 * a_d, 1e-2) and e_w <= max(2 * a_w, 5e-3)
    assert on.dtype == torch.float32 and od.shape == (1, E, 2) and ow.shape == (1, E, 2)


def test_update_state_dict_keys_match_reference_names():
    keys = set(Update(3).state_dict().keys())
    for k in ("c1.0.weight", "c2.2.bias", "norm.weight", "agg_kk.f.weight", "agg_ij.h.bias", "gru.0.weight",
              "gru.1.gate.0.weight", "gru.3.res.2.bias", "corr.0.weight", "corr.3.weight", "corr.5.bias", "d.1.weight", "w.1.bias"):
        assert k in keys
    assert keys == set(OU.Update(3).state_dict().keys())


def test_runner_graph_replay_equals_eager_steps(ext):
    """UpdateRunner.capture(): replaying the CUDA graph is bit-identical to launching the kernels one by one,
    including the recurrent state carried in place from one update to the next"""
    from dpvo_b200.runner import UpdateRunner
    outs = []
    for use_graph in (False, True):
        st = synthetic.make_state("fast", 14, device=DEV, seed=7)
        run = UpdateRunner(st, seed=3)
        if use_graph:
            run.capture()                       # two warm-up steps + the captured one have advanced the state:
            run.reset()                         # restore poses / patches and the recurrent state
            run.net.zero_()
        for _ in range(3):
            tgt, wgt = run.step_graph() if use_graph else run.step()
        torch.cuda.synchronize()
        outs.append((st.poses.clone(), st.patches.clone(), run.net.clone(), tgt.clone(), wgt.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
