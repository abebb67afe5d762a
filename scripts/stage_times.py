#!/usr/bin/env python
"""Per-stage timing of one guidance evaluation on BASELINE config 3 shapes (B=8, D=256, P=65536):
each stage_mask bit is launched alone, back to back, between CUDA events on the launching stream."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from followmyhold_b200.guidance.loop import GuidanceLoop
from followmyhold_b200.synthetic import make_guidance_sample, stack_samples

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--D", type=int, default=256)
ap.add_argument("--P", type=int, default=65536)
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--variant", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
samples = [make_guidance_sample(a.D, a.P, seed=i) for i in range(a.B)]
sdf0, theta0, st = stack_samples(samples, device=dev, cap=True)
loop = GuidanceLoop(a.B, a.D, st, a.P, device=dev, stream_variant=a.variant)
loop.sdf0.copy_(sdf0); loop.sdf.copy_(sdf0); loop.theta.copy_(theta0)
eng = loop.engine
desc = eng.make_desc(loop.sdf, loop.theta, st)
desc.stage_mask = 0
eng.launch(desc)
torch.cuda.synchronize()
out = {}
names = {1: "prep", 2: "stream", 4: "chamfer", 8: "raster+compact+voxdist", 16: "finalize", 0: "all"}
for m, n in names.items():
    desc.stage_mask = m
    for _ in range(3):
        eng.launch(desc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        eng.launch(desc)
    e1.record()
    torch.cuda.synchronize()
    out[n] = round(e0.elapsed_time(e1) / a.reps * 1e3, 2)
for name, serial, stages, pre in (("all_serial", 1, 0, 0), ("all_overlap_s4p2", 0, 4, 2), ("all_overlap_s5p3", 0, 5, 3),
                                  ("all_overlap_s6p3", 0, 6, 3), ("all_overlap_s4p3", 0, 4, 3), ("all_overlap_ldg", 0, 0, 0)):
    eng.serial, eng.stream_stages, eng.stream_prefetch = serial, stages, pre
    eng.stream_variant = 1 if name.endswith("ldg") else a.variant
    desc = eng.make_desc(loop.sdf, loop.theta, st)
    for _ in range(3):
        eng.launch(desc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        eng.launch(desc)
    e1.record()
    torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) / a.reps * 1e3, 2)
# stream alone: persistent CTAs per SM x ring depth x loads in flight
for ctas, stages, pre in ((2, 6, 3), (2, 4, 2), (2, 3, 2), (2, 3, 1), (2, 2, 1), (1, 8, 4), (1, 6, 3), (1, 4, 2), (1, 4, 3), (1, 6, 4),
                          (1, 12, 6), (1, 12, 4), (3, 3, 2), (3, 2, 1), (4, 3, 1), (4, 2, 1)):
    eng.serial, eng.stream_stages, eng.stream_prefetch, eng.stream_ctas, eng.stream_variant = 1, stages, pre, ctas, a.variant
    desc = eng.make_desc(loop.sdf, loop.theta, st)
    desc.stage_mask = 2
    for _ in range(3):
        eng.launch(desc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        eng.launch(desc)
    e1.record()
    torch.cuda.synchronize()
    out[f"stream_c{ctas}s{stages}p{pre}"] = round(e0.elapsed_time(e1) / a.reps * 1e3, 2)
print(json.dumps({"us_per_launch": out, "B": a.B, "D": a.D, "P": a.P}))
