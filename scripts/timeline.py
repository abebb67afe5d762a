#!/usr/bin/env python
"""Timeline of one guidance evaluation (BASELINE config 3 shapes) from the desc->trace hook:
per kernel, first CTA start and last CTA end relative to the start of the evaluation."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from followmyhold_b200.guidance.loop import GuidanceLoop
from followmyhold_b200.synthetic import make_guidance_sample, stack_samples

NAMES = ["prep", "stream", "h2c", "c2h", "chamfer_brute", "raster", "compact", "voxdist", "finalize_verts", "assemble", "keypoints"]
ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--D", type=int, default=256)
ap.add_argument("--P", type=int, default=65536)
ap.add_argument("--serial", type=int, default=0)
ap.add_argument("--stages", type=int, default=0)
ap.add_argument("--prefetch", type=int, default=0)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--evals", type=int, default=6)
a = ap.parse_args()
dev = torch.device("cuda:0")
samples = [make_guidance_sample(a.D, a.P, seed=i) for i in range(a.B)]
sdf0, theta0, st = stack_samples(samples, device=dev, cap=True)
loop = GuidanceLoop(a.B, a.D, st, a.P, device=dev, stream_variant=a.variant)
loop.sdf0.copy_(sdf0); loop.sdf.copy_(sdf0); loop.theta.copy_(theta0)
eng = loop.engine
eng.serial, eng.stream_stages, eng.stream_prefetch = a.serial, a.stages, a.prefetch
init = torch.tensor([[2 ** 63 - 1, 0]] * len(NAMES), dtype=torch.int64, device=dev)
trace = init.clone()
desc = eng.make_desc(loop.sdf, loop.theta, st)
desc.trace = trace.data_ptr()
rows = []
# one evaluation captured in a CUDA graph (as the loop runs it): no host launch latency in the timeline
side = torch.cuda.Stream()
eng.launch(desc)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=side):
    eng.launch(desc, side)
for it in range(a.evals):
    trace.copy_(init)
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    t = trace.cpu().numpy()
    live = [i for i in range(len(NAMES)) if t[i, 1] > 0]
    t0 = min(t[i, 0] for i in live)
    rows.append({NAMES[i]: [round((t[i, 0] - t0) / 1e3, 1), round((t[i, 1] - t0) / 1e3, 1)] for i in live})
print(json.dumps({"ncand": eng.terms[:, 14].tolist(), "config": vars(a), "us_start_end_last_eval": rows[-1], "us_start_end_prev_eval": rows[-2]}))
