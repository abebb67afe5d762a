#!/usr/bin/env python
"""Probe of the end-to-end path: H2D bandwidth alone and under a running step graph, and the pipelined
host API, for 1 and 2 micro-batches."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from followmyhold_b200.guidance.loop import GuidanceLoop
from followmyhold_b200.synthetic import make_guidance_sample, stack_samples

D, P, STEP = 256, 65536, 15
dev = torch.device("cuda:0")
samples = [make_guidance_sample(D, P, seed=i) for i in range(8)]
sdf0, theta0, st = stack_samples(samples, device=dev, cap=True)
out = {}
host = torch.empty(sdf0.shape, dtype=sdf0.dtype, pin_memory=True); host.copy_(sdf0)
dst = torch.empty_like(sdf0)
cs = torch.cuda.Stream()
def h2d_ms(n=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for _ in range(n):
            dst.copy_(host, non_blocking=True)
        e1.record(cs)
    cs.synchronize()
    return e0.elapsed_time(e1) / n
out["h2d_alone_ms_512MB"] = h2d_ms()
for m in (1, 2):
    lp = GuidanceLoop(8, D, st, P, device=dev, micro_batches=m)
    lp.sdf0.copy_(sdf0); lp.sdf.copy_(sdf0); lp.theta.copy_(theta0); lp.x_t.normal_(); lp.velocity.normal_().mul_(0.1)
    lp.capture(STEP)
    torch.cuda.synchronize()
    # H2D while 6 step graphs replay back to back
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(lp.stream):
        c0.record(lp.stream)
        for _ in range(6):
            lp._graph.replay()
        c1.record(lp.stream)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for _ in range(4):
            dst.copy_(host, non_blocking=True)
        e1.record(cs)
    torch.cuda.synchronize()
    out[f"m{m}_h2d_under_compute_ms_512MB"] = e0.elapsed_time(e1) / 4
    out[f"m{m}_step_ms_under_h2d"] = c0.elapsed_time(c1) / 6
    x_t_h = torch.randn(8, lp.L).pin_memory(); vel_h = (0.1 * torch.randn(8, lp.L)).pin_memory(); th_h = theta0.cpu().pin_memory()
    batch = (host, x_t_h, vel_h, th_h)
    lp.denoise_steps_host(STEP, [batch] * 8)
    t0 = time.perf_counter()
    lp.denoise_steps_host(STEP, [batch] * 8)
    out[f"m{m}_pipelined_ms_per_batch"] = (time.perf_counter() - t0) / 8 * 1e3
    out[f"m{m}_host_enqueue_ms_per_batch"] = lp._last_enqueue_s / 8 * 1e3
    del lp
    torch.cuda.empty_cache()
print(json.dumps(out))
