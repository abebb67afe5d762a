"""GPU probe of the tcgen05 attention forward against torch SDPA (fp32), then throughput."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from followmyhold_b200.decoder import tc

torch.manual_seed(0)
dev = "cuda:0"
res = {"cases": []}
all_ok = True


def ref_attn(q, k, v, n_img, q_shared):
    H = q.shape[1]
    kk = k.float().view(n_img, -1, H, 64).transpose(1, 2)
    vv = v.float().view(n_img, -1, H, 64).transpose(1, 2)
    qq = q.float().unsqueeze(0).expand(n_img, -1, -1, -1) if q_shared else q.float().view(n_img, -1, H, 64)
    o = F.scaled_dot_product_attention(qq.transpose(1, 2), kk, vv)
    return o.transpose(1, 2).reshape(n_img, -1, H * 64)


def case(name, n_img, H, n_q, n_k, q_shared, qscale=1.0, fused=False, max_ctas=0):
    for variant in (0, 2, 1):
        _case(f"v{variant} {name}", n_img, H, n_q, n_k, q_shared, qscale, fused, max_ctas, variant)


def _case(name, n_img, H, n_q, n_k, q_shared, qscale, fused, max_ctas, variant):
    global all_ok
    if fused:      # self-attention layout: one [tokens, heads, 192] projection
        qkv = torch.randn(n_img * n_k, H, 192, device=dev).half()
        q, k, v = qkv[:, :, :64], qkv[:, :, 64:128], qkv[:, :, 128:]
        q = q * qscale
    else:
        q = (qscale * torch.randn(n_q if q_shared else n_img * n_q, H, 64, device=dev)).half()
        kv = torch.randn(n_img * n_k, H, 128, device=dev).half()
        k, v = kv[:, :, :64], kv[:, :, 64:]
    out = tc.attention(q, k, v, n_img, q_shared=q_shared, max_ctas=max_ctas, variant=variant)
    torch.cuda.synchronize()
    ref = ref_attn(q, k, v, n_img, q_shared)
    err = (out.float() - ref).abs().max().item()
    sc = ref.abs().max().item()
    ok = err <= 3e-3 * sc + 1e-4
    all_ok &= ok
    res["cases"].append({"name": name, "err": err, "ref_max": sc, "ok": bool(ok)})
    print("ok  " if ok else "FAIL", name, err, sc, flush=True)


case("one tile one block", 1, 1, 128, 128, True)
case("one tile two blocks", 1, 1, 128, 256, True)
case("one tile 24 blocks", 1, 2, 128, 3072, True)
case("ragged queries", 1, 3, 200, 512, True)
case("several items per CTA", 2, 4, 1000, 1024, True, max_ctas=3)
case("peaked softmax (rescale path)", 1, 2, 256, 3072, True, qscale=6.0)
case("self attention fused qkv", 2, 16, 384, 384, False, fused=True)
case("cross attention 16 heads", 2, 16, 5000, 3072, True)
case("odd number of tiles", 1, 2, 128 * 3 + 5, 256, True)

perf = []
for (n_img, n_q) in ((1, 148 * 128), (1, 274625), (4, 65536)):
    H, n_k = 16, 3072
    q = torch.randn(n_q, H, 64, device=dev).half()
    kv = torch.randn(n_img * n_k, H, 128, device=dev).half()
    k, v = kv[:, :, :64], kv[:, :, 64:]
    out = torch.empty(n_img, n_q, H * 64, dtype=torch.float16, device=dev)
    fl = 4.0 * n_img * n_q * n_k * H * 64
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R = 10
    for variant in (0, 2, 1):
        for _ in range(2):
            tc.attention(q, k, v, n_img, out=out, q_shared=True, variant=variant)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(R):
            tc.attention(q, k, v, n_img, out=out, q_shared=True, variant=variant)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / R
        perf.append({"n_img": n_img, "n_q": n_q, "variant": variant, "ms": ms, "tflops": fl / ms / 1e9})
        print(perf[-1], flush=True)
    if n_q <= 65536:
        qq = q.unsqueeze(0).expand(n_img, -1, -1, -1).transpose(1, 2).contiguous()
        kk = k.reshape(n_img, n_k, H, 64).transpose(1, 2).contiguous(); vv = v.reshape(n_img, n_k, H, 64).transpose(1, 2).contiguous()
        for _ in range(2):
            F.scaled_dot_product_attention(qq, kk, vv)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(R):
            F.scaled_dot_product_attention(qq, kk, vv)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / R
        perf.append({"n_img": n_img, "n_q": n_q, "impl": "torch sdpa", "ms": ms, "tflops": fl / ms / 1e9})
        print(perf[-1], flush=True)
res["perf"] = perf
res["all_ok"] = bool(all_ok)
json.dump(res, open("gpurun_out/r02_attn_probe.json", "w"), indent=1)
print("ALL OK" if all_ok else "SOME FAILED")
