import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paddle_b200 as paddle
from paddle_b200._build import ext
from paddle_b200.nn import quant as Q
E = ext()
k, n, m = 5120, 15360, 16
ws = [torch.randn(k, n, device="cuda") * 0.02 for _ in range(3)]
qs = [Q.weight_quantize(w.as_subclass(paddle.Tensor), algo="weight_only_int8") for w in ws]
w8 = [(q[0].as_subclass(torch.Tensor), q[1].as_subclass(torch.Tensor).float()) for q in qs]
x = (torch.randn(m, k, device="cuda") * 0.5).to(torch.bfloat16)
for i in range(9):
    E.weight_only_gemm(x, *w8[i % 3], None, False)
torch.cuda.synchronize()
