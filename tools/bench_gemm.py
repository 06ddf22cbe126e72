"""Micro-benchmark of dpvo_linear_f16 variants vs cuBLAS (torch) on the update operator's shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpvo_b200
ex = dpvo_b200.extensions()[3]
dev = "cuda"
rows, N, K = 47712, 384, 384
g = torch.Generator(device=dev).manual_seed(0)
x = (torch.randn(rows, K, generator=g, device=dev) * 0.5).half()
w = (torch.randn(N, K, generator=g, device=dev) / K ** 0.5).half()
w2 = (torch.randn(768, K, generator=g, device=dev) / K ** 0.5).half()
xk = (torch.randn(rows, 896, generator=g, device=dev) * 0.5).half()
wk = (torch.randn(N, 896, generator=g, device=dev) / 30).half()
b = torch.randn(N, generator=g, device=dev)
b2 = torch.randn(768, generator=g, device=dev)
res = torch.randn(1, rows, N, generator=g, device=dev)
gate = torch.rand(1, rows, N, generator=g, device=dev).half()
idx = torch.randint(-1, rows, (rows,), generator=g, device=dev)
out32 = torch.empty(1, rows, N, device=dev)
out16 = torch.empty(1, rows, N, device=dev, dtype=torch.half)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

cases = {
    "plain f16": lambda: ex.linear_f16(x, w, b, 0),
    "relu f16": lambda: ex.linear_f16(x, w, b, 1),
    "sigmoid f16": lambda: ex.linear_f16(x, w, b, 2),
    "N=768": lambda: ex.linear_f16(x, w2, b2, 0),
    "K=896": lambda: ex.linear_f16(xk, wk, b, 1),
    "gather relu": lambda: ex.linear_f16(x, w, b, 1, gather=idx),
    "resadd f32+f16": lambda: ex.linear_f16(x, w, b, 3, res=res, out_f32=True, out=out32, out16=out16),
    "gatedres f32": lambda: ex.linear_f16(x, w, b, 4, res=res, gate=gate, out_f32=True, out=out32),
    "cublas plain": lambda: torch.nn.functional.linear(x, w, b.half()),
    "cublas N=768": lambda: torch.nn.functional.linear(x, w2, b2.half()),
    "cublas K=896": lambda: torch.nn.functional.linear(xk, wk, b.half()),
}
for style in ("1",):
    os.environ["DPVO_B200_EPI_STYLE"] = style
    for name, fn in cases.items():
        if style == "0" and name.startswith("cublas"):
            continue
        print("style %s  %-16s %8.1f us" % (style, name, timeit(fn)), flush=True)
