"""GPU probe of the tensor-core latent -> SDF decoder and its adjoint against the torch oracle (fp32)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followmyhold_b200.decoder.shapevae import DecoderWeights, LatentDecoder
from oracle import decoder_oracle as DO

torch.manual_seed(0)
dev = "cuda:0"
L = int(os.environ.get("LAYERS", "2"))
QC = int(os.environ.get("QC", "0"))
D = int(os.environ.get("D", "17"))
B = int(os.environ.get("B", "2"))
vae = DO.ShapeVAE(num_decoder_layers=L).float()
with torch.no_grad():
    for n, p in vae.named_parameters():          # non-trivial LayerNorm affine parameters and biases
        if n.endswith("norm.weight") or ".ln_" in n and n.endswith("weight"):
            p.add_(0.2 * torch.randn_like(p))
        elif n.endswith("bias"):
            p.add_(0.05 * torch.randn_like(p))
    vae.geo_decoder.output_proj.weight.mul_(4.0)
vae_g = vae.to(dev)
axis = torch.linspace(-1.10, 1.10, D)
xyz = torch.stack(torch.meshgrid(axis, axis, axis, indexing="ij"), -1).reshape(-1, 3)
W = DecoderWeights(vae.state_dict(), dev)
dec = LatentDecoder(W, B, query_chunk=QC or 2048, active_chunk=512 if D < 33 else 2048)
dec.set_queries(xyz)
lat = torch.randn(B, 3072, 64, device=dev)
res = {}

t0 = time.time()
sdf = dec.forward(lat)
torch.cuda.synchronize()
print("forward done", time.time() - t0, flush=True)
lat_o = lat.clone().requires_grad_(True)
ref = torch.stack([DO.latent2sdf(lat_o[b:b + 1], xyz.to(dev), (D, D, D), vae_g).reshape(-1) for b in range(B)])
with torch.no_grad():
    data_ref = vae_g(lat / vae.scale_factor)
    err_data = (dec.data.float().view(B, 3072, 1024) - data_ref).abs().max().item() / data_ref.abs().max().item()
    err = (sdf - ref).abs().max().item() / ref.abs().max().item()
print("transformer output rel err", err_data, " sdf rel err", err, " |sdf|max", ref.abs().max().item(), flush=True)
res.update(layers=L, D=D, B=B, err_transformer=err_data, err_sdf=err, sdf_absmax=ref.abs().max().item())

# adjoint: E = sum_m g_m sdf[idx_m]
M = int(os.environ.get("M", "1024"))
g = torch.Generator(device="cpu").manual_seed(1)
idx = torch.randint(0, D ** 3, (B, M), generator=g).to(torch.int32).to(dev)
gs = (torch.randn(B, M, generator=g) * 1e-2).to(dev)
E = sum((ref[b][idx[b].long()] * gs[b]).sum() for b in range(B))
E.backward()
gref = lat_o.grad
got = dec.backward(idx, gs)
torch.cuda.synchronize()
errg = (got - gref).abs().max().item() / gref.abs().max().item()
cos = torch.nn.functional.cosine_similarity(got.reshape(1, -1), gref.reshape(1, -1)).item()
print("adjoint rel err", errg, "cos", cos, "|g|max", gref.abs().max().item(), flush=True)
res.update(err_adjoint=errg, cos_adjoint=cos, grad_absmax=gref.abs().max().item())

# timings
for name, fn in (("forward", lambda: dec.forward(lat, out=sdf)), ("backward", lambda: dec.backward(idx, gs, out=got))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    res[name + "_ms"] = e0.elapsed_time(e1) / 3
    print(name, res[name + "_ms"], "ms", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/r02_decoder_probe_L{L}_D{D}_B{B}.json", "w"), indent=1)
