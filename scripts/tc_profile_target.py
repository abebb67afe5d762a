"""Short target for `ncu --set full`: one attention launch and one GEMM launch at the decode's shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followmyhold_b200.decoder import tc

dev = "cuda:0"
torch.manual_seed(0)
n = 148 * 128 * 2
q = torch.randn(n, 16, 64, device=dev).half()
kv = torch.randn(3072, 16, 128, device=dev).half()
o = torch.empty(1, n, 1024, dtype=torch.float16, device=dev)
a = torch.randn(n, 1024, device=dev).half()
w = torch.randn(4096, 1024, device=dev).half()
u = torch.empty(n, 4096, dtype=torch.float16, device=dev)
bias = torch.zeros(4096, device=dev)
for _ in range(3):
    tc.attention(q, kv[:, :, :64], kv[:, :, 64:], 1, out=o, q_shared=True)
    tc.gemm(a, w, out=u, bias=bias, act=tc.ACT_GELU)
torch.cuda.synchronize()
