#!/usr/bin/env python
"""Probe: does splitting the batch of 8 images into k independent loops (own graphs, own streams) hide
the serial tail of an evaluation (assemble -> decoder adjoint -> update -> decoder forward) behind the
other loops' dense stream?  Prints image-steps/s for k = 1, 2, 4."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from followmyhold_b200.guidance.loop import GuidanceLoop
from followmyhold_b200.synthetic import make_guidance_sample, stack_samples

D, P, STEP, K = 256, 65536, 15, 6
dev = torch.device("cuda:0")
samples = [make_guidance_sample(D, P, seed=i) for i in range(8)]
out = {}
for k, stages, pre in ((1, 0, 0), (2, 0, 0), (2, 5, 4), (2, 4, 3), (4, 0, 0), (4, 4, 3)):
    nb = 8 // k
    loops = []
    for j in range(k):
        sdf0, theta0, st = stack_samples(samples[j * nb:(j + 1) * nb], device=dev, cap=True)
        lp = GuidanceLoop(nb, D, st, P, device=dev)
        lp.engine.stream_stages, lp.engine.stream_prefetch = stages, pre
        lp.sdf0.copy_(sdf0); lp.sdf.copy_(sdf0); lp.theta.copy_(theta0)
        lp.x_t.normal_(); lp.velocity.normal_().mul_(0.1)
        lp.capture(STEP)
        loops.append(lp)
    torch.cuda.synchronize()
    def run_once():
        cur = torch.cuda.current_stream()
        for lp in loops:
            lp.stream.wait_stream(cur)
            with torch.cuda.stream(lp.stream):
                lp._graph.replay()
        for lp in loops:
            cur.wait_stream(lp.stream)
    for _ in range(2):
        run_once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        run_once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out[f"k{k}_s{stages}p{pre}"] = {"ms_per_step": round(ms, 3), "image_steps_per_s": round(8 / ms * 1e3, 1)}
    del loops
    torch.cuda.empty_cache()
print(json.dumps(out))
