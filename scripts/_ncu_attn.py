import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddle_b200._build import ext
E = ext()
b, s, h = 1, 4096, 40
q, k, v = (torch.randn(b, s, h, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
for _ in range(3):
    out, lse = E.attention_fwd(q, k, v, 128 ** -0.5, True)
g = torch.randn_like(out)
if len(sys.argv) > 1 and sys.argv[1] == "bwd":
    for _ in range(3):
        E.attention_bwd(q, k, v, out, lse, g, 128 ** -0.5, True)
torch.cuda.synchronize()
