"""GPU probe of the tcgen05 GEMM (run under gpurun with a timeout): every operand-major combination, tile
width, ragged shapes, batched strided views, epilogues -- against torch matmul; then throughput."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followmyhold_b200.decoder import tc

torch.manual_seed(0)
dev = "cuda:0"
res = {"cases": []}


def check(name, got, ref, tol):
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item() + 1e-9
    ok = err <= tol * scale
    res["cases"].append({"name": name, "max_err": err, "ref_max": scale, "ok": bool(ok)})
    print(("ok  " if ok else "FAIL"), name, err, scale, flush=True)
    return ok


def run_case(M, N, K, a_mn, b_mn, bn, batch=1):
    A = torch.randn(batch, M, K, device=dev).half()
    B = torch.randn(batch, N, K, device=dev).half()
    ref = torch.matmul(A.float(), B.float().transpose(1, 2))
    a = A.transpose(1, 2).contiguous() if a_mn else A
    b = B.transpose(1, 2).contiguous() if b_mn else B
    out = tc.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, block_n=bn)
    torch.cuda.synchronize()
    return check(f"M{M} N{N} K{K} a_mn{int(a_mn)} b_mn{int(b_mn)} bn{bn} batch{batch}", out, ref, 2e-3)


all_ok = True
# smallest first: a single tile, a single k-block
for (a_mn, b_mn) in ((False, False), (False, True), (True, False), (True, True)):
    for bn in (64, 128, 256):
        all_ok &= run_case(128, bn, 64, a_mn, b_mn, bn)
for (a_mn, b_mn) in ((False, False), (False, True), (True, False), (True, True)):
    all_ok &= run_case(256, 512, 256, a_mn, b_mn, 256)
    all_ok &= run_case(384, 192, 1024, a_mn, b_mn, 64)
    all_ok &= run_case(1000, 1096, 200, a_mn, b_mn, 128)       # ragged M, N, K
    all_ok &= run_case(136 if a_mn else 130, 72, 72, a_mn, b_mn, 0, batch=3)
all_ok &= run_case(5000, 4096, 1024, False, False, 256)
all_ok &= run_case(4096, 1024, 4096, False, False, 256)

# strided per-head views: S_h = Q_h K_h^T with Q [Mq, 16*64], K [3072, 16*64]
Mq, H, hd, T = 512, 16, 64, 3072
Q = torch.randn(Mq, H * hd, device=dev).half(); Kt = torch.randn(T, H * hd, device=dev).half(); V = torch.randn(T, H * hd, device=dev).half()
Qh = Q.view(Mq, H, hd).permute(1, 0, 2); Kh = Kt.view(T, H, hd).permute(1, 0, 2); Vh = V.view(T, H, hd).permute(1, 0, 2)
S = tc.gemm(Qh, Kh, alpha=0.125, out_dtype=torch.float32)
all_ok &= check("per-head QK^T strided", S, 0.125 * torch.matmul(Qh.float(), Kh.float().transpose(1, 2)), 2e-3)
Pm = torch.softmax(S, -1).half()
O = torch.empty(Mq, H * hd, device=dev, dtype=torch.float16)
tc.gemm(Pm, Vh, out=O.view(Mq, H, hd).permute(1, 0, 2), b_mn=True)          # B = V [keys, d] is MN-major
all_ok &= check("per-head PV (V MN-major, strided out)", O.view(Mq, H, hd).permute(1, 0, 2), torch.matmul(Pm.float(), Vh.float()), 2e-3)
dO = torch.randn(Mq, H * hd, device=dev).half(); dOh = dO.view(Mq, H, hd).permute(1, 0, 2)
dV = tc.gemm(Pm, dOh, a_mn=True, b_mn=True, out_dtype=torch.float32)        # P^T dO : both MN-major
all_ok &= check("per-head dV = P^T dO", dV, torch.matmul(Pm.float().transpose(1, 2), dOh.float()), 2e-3)

# epilogues
M, N, K = 300, 1024, 512
A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) / K ** 0.5).half(); bias = torch.randn(N, device=dev)
R = torch.randn(M, N, device=dev).half(); R32 = torch.randn(M, N, device=dev)
lin = A.float() @ W.float().t() + bias
pre = torch.empty(M, N, device=dev, dtype=torch.float16)
g = tc.gemm(A, W, bias=bias, act=tc.ACT_GELU, aux_out=pre)
all_ok &= check("bias+gelu", g, torch.nn.functional.gelu(lin), 2e-3)
all_ok &= check("aux_out pre-activation", pre, lin, 2e-3)
all_ok &= check("bias+residual fp16", tc.gemm(A, W, bias=bias, res=R), lin + R.float(), 2e-3)
all_ok &= check("residual fp32 out fp32", tc.gemm(A, W, res=R32, out_dtype=torch.float32), A.float() @ W.float().t() + R32, 2e-3)
x = pre.float().requires_grad_(True)
torch.nn.functional.gelu(x).backward(torch.ones_like(x))
dg = tc.gemm(A, W, act=tc.ACT_DGELU, aux_in=pre, out_dtype=torch.float32)
all_ok &= check("dgelu epilogue", dg, (A.float() @ W.float().t()) * x.grad, 3e-3)
acc = R32.clone()
tc.gemm(A, W, out=acc, res=acc)
all_ok &= check("in-place accumulate", acc, A.float() @ W.float().t() + R32, 2e-3)

# throughput
perf = []
for (M, N, K) in ((8192, 4096, 1024), (8192, 1024, 4096), (8192, 1024, 1024), (16384, 4096, 1024), (3072, 3072, 1024)):
    A = torch.randn(M, K, device=dev).half(); W = torch.randn(N, K, device=dev).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    for bn in (128, 256):
        for _ in range(3):
            tc.gemm(A, W, out=out, block_n=bn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            tc.gemm(A, W, out=out, block_n=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        perf.append({"M": M, "N": N, "K": K, "bn": bn, "ms": ms, "tflops": 2 * M * N * K / ms / 1e9})
        print(perf[-1], flush=True)
    for _ in range(3):
        torch.matmul(A, W.t(), out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        torch.matmul(A, W.t(), out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    perf.append({"M": M, "N": N, "K": K, "impl": "cublas", "ms": ms, "tflops": 2 * M * N * K / ms / 1e9})
    print(perf[-1], flush=True)
# the 1024 -> 4096 projection with bias + GELU (the per-query MLP of the lattice decode)
M, N, K = 32768, 4096, 1024
A = torch.randn(M, K, device=dev).half(); W = torch.randn(N, K, device=dev).half() * 0.03; bias = torch.randn(N, device=dev)
out = torch.empty(M, N, dtype=torch.float16, device=dev)
for name, kw in (("plain", {}), ("bias", {"bias": bias}), ("bias+gelu", {"bias": bias, "act": tc.ACT_GELU})):
    for _ in range(3):
        tc.gemm(A, W, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tc.gemm(A, W, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    perf.append({"M": M, "N": N, "K": K, "epilogue": name, "ms": ms, "tflops": 2 * M * N * K / ms / 1e9})
    print(perf[-1], flush=True)
# the 1024 -> 1024 projection with bias + half residual (c_proj + x0 of the lattice decode) and the 4096 -> 1024 one (fc2 + x)
for (M, N, K) in ((32768, 1024, 1024), (32768, 1024, 4096)):
    A = torch.randn(M, K, device=dev).half(); W = torch.randn(N, K, device=dev).half() * 0.03; bias = torch.randn(N, device=dev)
    R = torch.randn(M, N, device=dev).half()
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    for name, kw in (("bias", {"bias": bias}), ("bias+residual", {"bias": bias, "res": R})):
        for _ in range(3):
            tc.gemm(A, W, out=out, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            tc.gemm(A, W, out=out, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        perf.append({"M": M, "N": N, "K": K, "epilogue": name, "ms": ms, "tflops": 2 * M * N * K / ms / 1e9})
        print(perf[-1], flush=True)
res["perf"] = perf
res["all_ok"] = bool(all_ok)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r02_tc_probe.json", "w"), indent=1)
print("ALL OK" if all_ok else "SOME FAILED")
