"""Perf attribution of the fused GEMM epilogues: DPVO_B200_GEMM_EXP bit flags switch single operand streams off."""
import os, sys, subprocess
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import os, sys, torch
sys.path.insert(0, os.path.dirname(%r))
import dpvo_b200
ex = dpvo_b200.extensions()[3]
rows, N, K = 47712, 384, 384
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(rows, K, generator=g, device="cuda") * 0.5).half()
w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).half()
b = torch.randn(N, generator=g, device="cuda")
res = torch.randn(1, rows, N, generator=g, device="cuda")
gate = torch.rand(1, rows, N, generator=g, device="cuda").half()
out32 = torch.empty(1, rows, N, device="cuda"); out16 = torch.empty(1, rows, N, device="cuda", dtype=torch.half)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c) * 1e3)
    return sorted(ts)[5]
print("EXP=%%s  resadd %%.1f  resadd-inplace %%.1f  gatedres %%.1f" %% (os.environ.get("DPVO_B200_GEMM_EXP", "0"),
      t(lambda: ex.linear_f16(x, w, b, 3, res=res, out_f32=True, out=out32, out16=out16)),
      t(lambda: ex.linear_f16(x, w, b, 3, res=res, out_f32=True, out=res, out16=out16)),
      t(lambda: ex.linear_f16(x, w, b, 4, res=res, gate=gate, out_f32=True, out=out32))))
''' % here
for flags in (0, 1, 2, 3, 4, 8, 12, 15):
    env = dict(os.environ, DPVO_B200_GEMM_EXP=str(flags))
    subprocess.run([sys.executable, "-c", code], env=env)
env = dict(os.environ, DPVO_B200_GEMM_NOPREFETCH="1")
subprocess.run([sys.executable, "-c", code], env=env)
