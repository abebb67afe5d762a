"""GPU probe of the fused attention adjoint (k_attn_bwd) against torch autograd in fp32; then its time at the transformer's shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followmyhold_b200.decoder import tc

LOG2E = 1.4426950408889634
dev = "cuda:0"
torch.manual_seed(0)
res = []


def case(n_img, heads, n_q, n_k, fused_views=False):
    if fused_views:                                   # q | k | v interleaved per head like the fused projection
        qkv = (torch.randn(n_img * n_q, heads, 192, device=dev) * 0.7).half()
        q, k, v = qkv[:, :, :64], qkv[:, :, 64:128], qkv[:, :, 128:]
    else:
        q = (torch.randn(n_img * n_q, heads, 64, device=dev) * 0.7).half()
        k = (torch.randn(n_img * n_k, heads, 64, device=dev) * 0.7).half()
        v = torch.randn(n_img * n_k, heads, 64, device=dev).half()
    do = (torch.randn(n_img * n_q, heads, 64, device=dev) * 0.5).half()
    qf, kf, vf = (t.float().view(n_img, -1, heads, 64).transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * 0.125
    o = torch.softmax(s, -1) @ vf
    dof = do.float().view(n_img, n_q, heads, 64).transpose(1, 2)
    (o * dof).sum().backward()
    lse2 = (torch.logsumexp(s, -1) * LOG2E).detach().contiguous()            # [n_img, heads, n_q]
    delta = (o * dof).sum(-1).detach().contiguous()
    dq = torch.full((n_img * n_q, heads, 64), float("nan"), device=dev, dtype=torch.float16)
    dk = torch.full((n_img * n_k, heads, 64), float("nan"), device=dev, dtype=torch.float16)
    dv = torch.full((n_img * n_k, heads, 64), float("nan"), device=dev, dtype=torch.float16)
    tc.attention_bwd(q, k, v, do, lse2, delta, dq, dk, dv, n_img)
    torch.cuda.synchronize()
    out = {"n_img": n_img, "heads": heads, "n_q": n_q, "n_k": n_k, "fused_views": fused_views}
    for name, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        r = ref.transpose(1, 2).reshape(got.shape)
        err = float((got.float() - r).abs().max() / r.abs().max())
        out["err_" + name] = err
        assert err < 6e-3, (name, err, out)
    # determinism
    dq2, dk2, dv2 = torch.empty_like(dq), torch.empty_like(dk), torch.empty_like(dv)
    tc.attention_bwd(q, k, v, do, lse2, delta, dq2, dk2, dv2, n_img)
    assert torch.equal(dq, dq2) and torch.equal(dk, dk2) and torch.equal(dv, dv2)
    res.append(out)
    print(out, flush=True)


if not os.environ.get("FOHO_ATTN_BWD_DBG"):
    case(1, 1, 128, 128)
    case(1, 2, 256, 384)
    case(2, 4, 300, 256)
    case(1, 16, 1024, 1024, fused_views=True)
# time at the transformer's shape: 3072 tokens, 16 heads
n = 3072
q = (torch.randn(n, 16, 64, device=dev) * 0.7).half(); k = (torch.randn(n, 16, 64, device=dev) * 0.7).half()
v = torch.randn(n, 16, 64, device=dev).half(); do = torch.randn(n, 16, 64, device=dev).half()
lse2 = torch.full((1, 16, n), 12.0, device=dev); delta = torch.zeros(1, 16, n, device=dev)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
for _ in range(3):
    tc.attention_bwd(q, k, v, do, lse2, delta, dq, dk, dv, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    tc.attention_bwd(q, k, v, do, lse2, delta, dq, dk, dv, 1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
flops = 7 * 2 * n * n * 64 * 16
print(json.dumps({"cases": res, "transformer_shape": {"ms": ms, "tflops_executed_7_products": flops / ms / 1e9,
                                                       "tflops_algorithmic_5_products": flops * 5 / 7 / ms / 1e9}}))
print("TIMING", ms, "ms", flops * 5 / 7 / ms / 1e9, "TF/s algorithmic")
print("ALL OK")
