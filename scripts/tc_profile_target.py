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
# the fused attention adjoint at the transformer's shape (3072 tokens, 16 heads)
t = 3072
qb = (torch.randn(t, 16, 64, device=dev) * 0.7).half(); kb = (torch.randn(t, 16, 64, device=dev) * 0.7).half()
vb = torch.randn(t, 16, 64, device=dev).half(); dob = torch.randn(t, 16, 64, device=dev).half()
lse = torch.full((1, 16, t), 12.0, device=dev); dl = torch.zeros(1, 16, t, device=dev)
dq, dk, dv = torch.empty_like(qb), torch.empty_like(kb), torch.empty_like(vb)
for _ in range(3):
    tc.attention_bwd(qb, kb, vb, dob, lse, dl, dq, dk, dv, 1)
torch.cuda.synchronize()
