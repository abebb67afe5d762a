"""Is the decoder's forward bound by the host's launch rate?  Eager launches against a CUDA-graph replay of the same call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followmyhold_b200.decoder.shapevae import DecoderWeights, LatentDecoder, lattice_points, random_state_dict

dev = "cuda:0"
B = int(os.environ.get("B", "1"))
W = DecoderWeights(random_state_dict(seed=0), dev)
dec = LatentDecoder(W, B)
dec.set_queries(lattice_points(65))
lat = torch.randn(B, 3072, 64, device=dev)
out = torch.empty(B, 65 ** 3, dtype=torch.float32, device=dev)
g = torch.Generator().manual_seed(1)
idx = torch.randint(0, 65 ** 3, (B, 8192), generator=g).to(torch.int32).to(dev)
gs = (torch.randn(B, 8192, generator=g) * 1e-2).to(dev)
gout = torch.empty(B, 3072, 64, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        dec.forward(lat, out=out, stream=s); dec.backward(idx, gs, out=gout, stream=s)
    s.synchronize()

    def timed(fn, R=5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(R):
            fn()
        e1.record(s); s.synchronize()
        return e0.elapsed_time(e1) / R
    print("eager forward ms", timed(lambda: dec.forward(lat, out=out, stream=s)), "adjoint ms", timed(lambda: dec.backward(idx, gs, out=gout, stream=s)), flush=True)
    ref = out.clone(); gref = gout.clone()
    gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(gf, stream=s):
        dec.forward(lat, out=out, stream=s)
    with torch.cuda.graph(gb, stream=s):
        dec.backward(idx, gs, out=gout, stream=s)
    out.zero_(); gout.zero_()
    print("graph forward ms", timed(gf.replay), "adjoint ms", timed(gb.replay), flush=True)
    gf.replay(); gb.replay(); s.synchronize()
    print("same results:", torch.equal(out, ref), torch.equal(gout, gref))
