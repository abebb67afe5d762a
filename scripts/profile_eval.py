#!/usr/bin/env python
"""Short driver for ncu: a few guidance evaluations (+ fused update) on BASELINE config 3
shapes (B=8, D=256, P=65536), no CUDA graph, so every kernel shows up as its own launch."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from followmyhold_b200.guidance.loop import GuidanceLoop
from followmyhold_b200.synthetic import make_guidance_sample, stack_samples

ap = argparse.ArgumentParser()
ap.add_argument("--evals", type=int, default=3)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--D", type=int, default=256)
ap.add_argument("--P", type=int, default=65536)
a = ap.parse_args()
dev = torch.device("cuda:0")
samples = [make_guidance_sample(a.D, a.P, seed=i) for i in range(a.B)]
sdf0, theta0, st = stack_samples(samples, device=dev, cap=True)
loop = GuidanceLoop(a.B, a.D, st, a.P, device=dev, stream_variant=a.variant)
loop.sdf0.copy_(sdf0); loop.sdf.copy_(sdf0); loop.theta.copy_(theta0)
s = torch.cuda.current_stream()
for _ in range(a.evals):
    loop._enqueue_eval(0.5, False, s)
torch.cuda.synchronize()
print(loop.terms[:, 0].tolist())
