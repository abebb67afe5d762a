#!/usr/bin/env python
"""Timeline of ONE inner iteration of BOTH micro-batch lanes in the steady state of the captured denoise step
(BASELINE config 1 shapes): per lane and kernel, first CTA start and last CTA end (desc->trace hook, %globaltimer),
relative to the earlier lane's start.  Shows what runs beside what once the lanes have drifted into their pattern."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from followmyhold_b200.guidance import engine as E
from followmyhold_b200.guidance.loop import GuidanceLoop
from followmyhold_b200.synthetic import make_guidance_sample, stack_samples

NAMES = ["prep", "stream", "h2c", "c2h", "chamfer_brute", "raster", "compact", "voxdist", "finalize_verts", "assemble", "keypoints"]
ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--D", type=int, default=256)
ap.add_argument("--P", type=int, default=65536)
ap.add_argument("--lanes", type=int, default=2)
ap.add_argument("--which", type=int, nargs="+", default=[24, 25], help="inner iterations to trace")
ap.add_argument("--step", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
samples = [make_guidance_sample(a.D, a.P, seed=i) for i in range(a.B)]
sdf0, theta0, st = stack_samples(samples, device=dev, cap=True)
loop = GuidanceLoop(a.B, a.D, st, a.P, device=dev, micro_batches=a.lanes)
g = torch.Generator().manual_seed(5)
loop.sdf0.copy_(sdf0); loop.sdf.copy_(sdf0); loop.theta.copy_(theta0)
loop.x_t.copy_(torch.randn(a.B, loop.L, generator=g)); loop.velocity.copy_(0.1 * torch.randn(a.B, loop.L, generator=g))
init = torch.tensor([[2 ** 63 - 1, 0]] * len(NAMES), dtype=torch.int64, device=dev)
traces = {(ln.engine.lane, k): init.clone() for ln in loop.lanes for k in a.which}
count = {}
orig = E.GuidanceEngine.make_desc


def make_desc(self, *args, **kw):
    d = orig(self, *args, **kw)
    k = count.get(self.lane, 0)
    count[self.lane] = k + 1
    if (self.lane, k) in traces:
        d.trace = traces[(self.lane, k)].data_ptr()
    return d


E.GuidanceEngine.make_desc = make_desc
loop.capture(a.step)
E.GuidanceEngine.make_desc = orig
for _ in range(3):
    loop.run_step_device(a.step)
torch.cuda.synchronize()
for t in traces.values():
    t.copy_(init)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
loop.run_step_device(a.step)
e1.record()
torch.cuda.synchronize()
raw = {k: t.cpu().numpy() for k, t in traces.items()}
t0 = min(int(v[i, 0]) for v in raw.values() for i in range(len(NAMES)) if v[i, 1] > 0)
out = {}
for (lane, k), v in sorted(raw.items()):
    out[f"lane{lane}_iter{k}"] = {NAMES[i]: [round((int(v[i, 0]) - t0) / 1e3, 1), round((int(v[i, 1]) - t0) / 1e3, 1)]
                                   for i in range(len(NAMES)) if v[i, 1] > 0}
print(json.dumps({"config": vars(a), "step_ms": e0.elapsed_time(e1), "inner_iterations": loop.phase_iterations(2),
                  "us_per_iteration_per_lane": e0.elapsed_time(e1) * 1e3 / loop.phase_iterations(2), "us_start_end": out}))
