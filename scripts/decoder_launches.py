"""Target for the ncu launch list of the tensor-core decoder: 3 forwards + 3 adjoints at the bench's shapes (B = 1, 65^3 queries,
16 layers, 8192 lattice points carrying a gradient), no oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followmyhold_b200.decoder.shapevae import DecoderWeights, LatentDecoder, lattice_points, random_state_dict

dev = "cuda:0"
D, B = int(os.environ.get("D", "65")), int(os.environ.get("B", "1"))
W = DecoderWeights(random_state_dict(seed=0), dev)
dec = LatentDecoder(W, B)
dec.set_queries(lattice_points(D))
lat = torch.randn(B, 3072, 64, device=dev)
g = torch.Generator().manual_seed(1)
idx = torch.randint(0, D ** 3, (B, 8192), generator=g).to(torch.int32).to(dev)
gs = (torch.randn(B, 8192, generator=g) * 1e-2).to(dev)
for _ in range(3):
    dec.forward(lat)
    dec.backward(idx, gs)
torch.cuda.synchronize()
print("ok")
