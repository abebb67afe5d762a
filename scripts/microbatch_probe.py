#!/usr/bin/env python
"""Probe: image-steps/s of the step graph for m micro-batches (lanes inside one graph) and several
launch shapes of the dense stream kernel (ring slots, bulk loads in flight per CTA)."""
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
sdf0, theta0, st = stack_samples(samples, device=dev, cap=True)
out = {}
for m, stages, pre in ((1, 0, 0), (2, 0, 0), (2, 4, 3), (2, 4, 2), (2, 6, 3), (2, 5, 3), (2, 3, 2), (4, 4, 2), (4, 3, 2)):
    lp = GuidanceLoop(8, D, st, P, device=dev, micro_batches=m)
    for ln in lp.lanes:
        ln.engine.stream_stages, ln.engine.stream_prefetch = stages, pre
    lp.sdf0.copy_(sdf0); lp.sdf.copy_(sdf0); lp.theta.copy_(theta0)
    lp.x_t.normal_(); lp.velocity.normal_().mul_(0.1)
    lp.capture(STEP)
    for _ in range(2):
        lp.run_step_device(STEP)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        lp.run_step_device(STEP)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out[f"m{m}_s{stages}p{pre}"] = {"ms_per_step": round(ms, 3), "image_steps_per_s": round(8 / ms * 1e3, 1)}
    del lp
    torch.cuda.empty_cache()
print(json.dumps(out))
