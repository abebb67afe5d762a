#!/usr/bin/env python
"""Timing of the trimmed similarity ICP (REF a16, src/foho/alignment/mesh_align.py:56-175) at the two
stage sizes both reference callers use (h2m.py:40-53, mano.py:29-42): coarse 50 iterations with
1 000 source / 5 000 target samples, fine 100 iterations with 5 000 / 10 000, 20 % outliers,
scale clip [0.7, 3].  GPU = foho_icp_run (whole loop on the device, float64); CPU = the oracle
(scipy cKDTree + numpy, one thread, as the reference runs it).  Checks that both return the same
transform before reporting."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ctypes as C

from followmyhold_b200 import _lib
from followmyhold_b200.alignment.mesh_align import icp_points, icp_points_many
from oracle import icp_oracle as IO


def device_ms(src_np, tgt_np, n_iter, n_out, reps=10):
    """foho_icp_run with the point sets resident, CUDA events on the launching stream: target sort + boxes + source sort
    + the persistent loop (what one call costs once the samples are on the device)."""
    lib = _lib.load()
    dev = torch.device("cuda:0")
    src = torch.as_tensor(np.ascontiguousarray(src_np)).to(dev)
    tgt = torch.as_tensor(np.ascontiguousarray(tgt_np)).to(dev)
    nbytes = lib.foho_icp_workspace_bytes(src.shape[0], tgt.shape[0])
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    ws_ptr = ws.data_ptr() + ((-ws.data_ptr()) % 256)
    T = torch.zeros(16, dtype=torch.float64, device=dev)
    cost = torch.zeros(1, dtype=torch.float64, device=dev)
    s = torch.cuda.current_stream(dev)

    def call(k):
        _lib.check("foho_icp_run", lib.foho_icp_run(src.data_ptr(), src.shape[0], tgt.data_ptr(), tgt.shape[0], k, n_out, 0, 0.7, 3.0,
                                                    T.data_ptr(), cost.data_ptr(), None, None, C.c_void_p(ws_ptr), nbytes,
                                                    C.c_void_p(s.cuda_stream)))
    out = []
    for k in (n_iter, 2 * n_iter):
        call(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call(k)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps)
    # two lengths: the difference is the loop alone, the rest is the per-call setup (sorts, boxes)
    per_iter_us = (out[1] - out[0]) / n_iter * 1e3
    return {"call_ms": out[0], "loop_us_per_iter": per_iter_us, "setup_ms": out[0] - per_iter_us * n_iter / 1e3}


rng = np.random.default_rng(0)
out = {}
for name, n_iter, ns, nt in (("coarse", 50, 1000, 5000), ("fine", 100, 5000, 10000)):
    tgt = rng.normal(size=(nt, 3)) * np.array([1.0, 0.6, 0.3])
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    R *= np.sign(np.linalg.det(R))
    src = ((tgt[rng.choice(nt, ns, replace=False)] - 0.05) @ R.T) / 1.15 + 0.002 * rng.normal(size=(ns, 3))
    src = src @ np.linalg.inv(R).T * 0.98            # a few degrees / per cent off, like after the init transform
    n_out = int(0.2 * ns)
    icp_points(src, tgt, 2, n_out, False, 0.7, 3.0)  # warm-up (library load, allocator)
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        T, c = icp_points(src, tgt, n_iter, n_out, False, 0.7, 3.0)
    gpu_s = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    To, co = IO.icp_points(src, tgt, n_iter, n_out, False, 0.7, 3.0)
    cpu_s = time.perf_counter() - t0
    assert np.abs(T - To).max() < 1e-7 and abs(c - co) < 1e-9, (np.abs(T - To).max(), c, co)
    # eight images' loops in one persistent launch (foho_icp_run_batch)
    probs = [(src + 0.001 * k, tgt) for k in range(8)]
    icp_points_many(probs, 2, n_out, False, 0.7, 3.0)
    t0 = time.perf_counter()
    many = icp_points_many(probs, n_iter, n_out, False, 0.7, 3.0)
    many_s = time.perf_counter() - t0
    T0, c0 = icp_points(probs[3][0], probs[3][1], n_iter, n_out, False, 0.7, 3.0)
    assert np.abs(many[3][0] - T0).max() == 0.0 and many[3][1] == c0
    out[name + "_x8"] = {"loops": 8, "total_ms": many_s * 1e3, "ms_per_loop": many_s / 8 * 1e3,
                         "speedup_vs_one_at_a_time": gpu_s * 8 / many_s}
    out[name + "_device"] = device_ms(src, tgt, n_iter, n_out)
    out[name] = {"n_iter": n_iter, "Ns": ns, "Nt": nt, "gpu_ms": gpu_s * 1e3, "gpu_us_per_iter": gpu_s / n_iter * 1e6,
                 "cpu_oracle_ms": cpu_s * 1e3, "speedup": cpu_s / gpu_s,
                 "gpu_pair_evals_per_s": ns * nt * n_iter / gpu_s, "max_abs_T_diff": float(np.abs(T - To).max())}
print(json.dumps({"icp": out, "legacy_launch_pairs": os.environ.get("FOHO_ICP_LEGACY") == "1",
                  "note": "gpu_ms includes H2D of the point sets, the loop, D2H of T (host clock); *_device: point sets resident, "
                          "CUDA events; one alignment = coarse + fine"}))
