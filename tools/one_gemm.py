import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpvo_b200
ex = dpvo_b200.extensions()[3]
rows, N, K = 47712, 384, 384
x = (torch.randn(rows, K, device="cuda") * 0.5).half()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
b = torch.randn(N, device="cuda")
epi = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(4):
    y = ex.linear_f16(x, w, b, epi)
torch.cuda.synchronize()
