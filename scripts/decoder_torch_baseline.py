"""What the reference's own execution model costs for row f1 on this GPU: ``latent2sdf`` (pipelines.py:292-312) as PyTorch
eager fp16 modules (cuBLAS GEMMs, SDPA flash attention) with autograd through the whole lattice, which is how the reference
back-propagates every inner iteration (pipelines.py:1392-1399 -> loss.backward()).  The modules are the oracle's restatement of
hy3dgen's ShapeVAE (oracle/decoder_oracle.py, architecture unpinned); random weights, synthetic latents; same shapes as
bench.py's decoder leg.  A measurement aid like bench.py's cpu_baseline leg -- not product code."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import decoder_oracle as DO

dev = "cuda:0"
D = int(os.environ.get("D", "65"))
torch.manual_seed(0)
vae = DO.ShapeVAE().half().to(dev)
for p_ in vae.parameters():
    p_.requires_grad_(False)                      # frozen networks, as in the reference's guidance loop
axis = torch.linspace(-1.10, 1.10, D)
xyz = torch.stack(torch.meshgrid(axis, axis, axis, indexing="ij"), -1).reshape(-1, 3).to(dev)
g = torch.Generator().manual_seed(1)
idx = torch.randint(0, D ** 3, (8192,), generator=g).to(dev)
gs = (torch.randn(8192, generator=g) * 1e-2).to(dev)
res = []
for chunk in (8000, 65536):                       # the reference's chunk (pipelines.py:300) and a launch-friendlier one
    lat = torch.randn(1, 3072, 64, device=dev, dtype=torch.float16, requires_grad=True)
    times = []
    for rep in range(4):
        lat.grad = None
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        sdf = DO.latent2sdf(lat, xyz, (D, D, D), vae, num_chunks=chunk)
        e[1].record()
        (sdf.reshape(-1)[idx] * gs).sum().backward()
        e[2].record()
        torch.cuda.synchronize()
        times.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    fwd = min(t[0] for t in times[1:]); bwd = min(t[1] for t in times[1:])
    with torch.no_grad():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        DO.latent2sdf(lat, xyz, (D, D, D), vae, num_chunks=chunk)
        e0.record()
        DO.latent2sdf(lat, xyz, (D, D, D), vae, num_chunks=chunk)
        e1.record(); torch.cuda.synchronize()
        fwd_ng = e0.elapsed_time(e1)
    res.append({"chunk": chunk, "forward_with_graph_ms": fwd, "backward_ms": bwd, "evaluation_ms": fwd + bwd, "forward_no_grad_ms": fwd_ng,
                "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30})
    print(res[-1], flush=True)
    del sdf, lat
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
json.dump({"workload": f"latent2sdf + loss.backward(), 1 image, {D}^3 lattice, 16-layer ShapeVAE decoder, fp16, PyTorch eager (torch " + torch.__version__ + ")",
           "runs": res}, open("gpurun_out/r02_decoder_torch_baseline.json", "w"), indent=1)
