"""One GEMM variant in isolation, for ncu captures:  python tools/one_gemm.py <epilogue 0..4>"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpvo_b200
ex = dpvo_b200.extensions()[3]
rows, N, K = 47712, 384, 384
x = (torch.randn(rows, K, device="cuda") * 0.5).half()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
b = torch.randn(N, device="cuda")
res = torch.randn(1, rows, N, device="cuda")
gate = torch.rand(1, rows, N, device="cuda").half()
out32 = torch.empty(1, rows, N, device="cuda")
out16 = torch.empty(1, rows, N, device="cuda", dtype=torch.half)
epi = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(4):
    if epi == 3:
        ex.linear_f16(x, w, b, 3, res=res, out_f32=True, out=out32, out16=out16)
    elif epi == 4:
        ex.linear_f16(x, w, b, 4, res=res, gate=gate, out_f32=True, out=out32, out16=out16)
    else:
        ex.linear_f16(x, w, b, epi)
torch.cuda.synchronize()
