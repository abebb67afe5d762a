"""Forward / adjoint time of the tensor-core decoder at the bench's shape as a function of the query chunk."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followmyhold_b200.decoder.shapevae import DecoderWeights, LatentDecoder, lattice_points, random_state_dict

dev = "cuda:0"
D, B = int(os.environ.get("D", "65")), int(os.environ.get("B", "1"))
W = DecoderWeights(random_state_dict(seed=0), dev)
res = []
for qc in [int(x) for x in (sys.argv[1:] or ["32768", "65536", "137344", "274688"])]:
    dec = LatentDecoder(W, B, query_chunk=qc // B // 128 * 128)
    dec.set_queries(lattice_points(D))
    lat = torch.randn(B, 3072, 64, device=dev)
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(0, D ** 3, (B, 8192), generator=g).to(torch.int32).to(dev)
    gs = (torch.randn(B, 8192, generator=g) * 1e-2).to(dev)
    for _ in range(2):
        dec.forward(lat); dec.backward(idx, gs)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    R = 5
    e[0].record()
    for _ in range(R):
        dec.forward(lat)
    e[1].record()
    for _ in range(R):
        dec.backward(idx, gs)
    e[2].record(); torch.cuda.synchronize()
    res.append({"B": B, "query_chunk": dec.query_chunk, "forward_ms": e[0].elapsed_time(e[1]) / R, "adjoint_ms": e[1].elapsed_time(e[2]) / R})
    print(res[-1], flush=True)
    del dec
    torch.cuda.empty_cache()
json.dump(res, open(f"gpurun_out/r02_chunk_probe_B{B}.json", "w"), indent=1)
