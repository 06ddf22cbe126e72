"""Times the add_layernorm variants of the update operator (register kernel vs bulk-staged kernel)."""
import os, sys, subprocess
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import os, sys, torch
sys.path.insert(0, os.path.dirname(%r))
import dpvo_b200
ex = dpvo_b200.extensions()[3]
E, D = 47712, 384
g = torch.Generator(device="cuda").manual_seed(0)
net = torch.randn(1, E, D, generator=g, device="cuda")
h16 = torch.randn(1, E, D, generator=g, device="cuda").half()
imap = torch.randn(1, 3456, D, generator=g, device="cuda").half()
idx = torch.randint(0, 3456, (E,), generator=g, device="cuda")
ga = torch.rand(1, E, 2 * D, generator=g, device="cuda").half()
gamma, beta = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c) * 1e3)
    return sorted(ts)[5]
print("%%s  LN1 f16->f16 relu %%.1f | LN2 f32+gather+f16 inplace %%.1f | GRU1 f32+gather inplace %%.1f | GRU2 f32+gate*res inplace %%.1f | plain f32 %%.1f" %% (
    "register" if os.environ.get("DPVO_B200_LN_REGISTER") else "bulk    ",
    t(lambda: ex.add_layernorm(h16, None, None, gamma, beta, 1e-3, True, False, True)),
    t(lambda: ex.add_layernorm(net, imap, h16, gamma, beta, 1e-3, False, True, True, idx, True)),
    t(lambda: ex.add_layernorm(net, imap, None, gamma, beta, 1e-3, False, True, True, idx, True)),
    t(lambda: ex.add_layernorm(net, None, h16, gamma, beta, 1e-3, False, True, True, None, True, ga[..., :D])),
    t(lambda: ex.add_layernorm(net, None, None, gamma, beta, 1e-3, False, True, True, None, True))))
''' % here
for reg in (False, True):
    env = dict(os.environ)
    if reg:
        env["DPVO_B200_LN_REGISTER"] = "1"
    subprocess.run([sys.executable, "-c", code], env=env)
