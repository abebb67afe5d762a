#!/usr/bin/env python
"""Benchmark of the guidance hot path (BASELINE.json metric: guided-denoise steps/sec;
guidance kernel HBM GB/s).

Workload at every N: BASELINE.json configs[2] -- per GPU a batch of 8 synthetic 256^3
volumes + random MANO-topology hand poses + 65 536-point clouds, mock latents.  One bench
"step" = one guided-denoise step of the whole batch = 50 guidance evaluations (decode ->
fused energy fwd+bwd -> decoder adjoint -> fused AdamW/step_final) + one scheduler.step,
replayed as ONE CUDA graph.  value = image-steps per second over all GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (ours)
  python bench.py --impl reference ...                            (CPU oracle arm)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # see followmyhold_b200/__init__.py

METRIC = "guided_denoise_steps_per_sec"
UNIT = "image-steps/s"
B_PER_GPU, D, P, EVALS_PER_STEP = 8, 256, 65536, 50
STEP_INDEX = 15          # a phase-2 outer step of the 20-step schedule (pipelines.py:1455)


def algorithmic_bytes_per_eval(D: int, P: int, Vh: int = 778) -> int:
    """read SDF once + write dense dE/dSDF once + verts/cloud in, grads out."""
    return 4 * D ** 3 + 4 * D ** 3 + 2 * 12 * (Vh + P)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[1])); mx = max(mx, float(s[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one stream-kernel launch, from the committed
    `ncu --set full` capture of the same workload (profiles/); None when absent."""
    p = os.path.join(ROOT, "profiles", "r01_stream_kernel_ncu.json")
    try:
        return int(json.load(open(p))["traffic_bytes_per_launch"])
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# --------------------------------------------------------------------------- CPU arm
def workload_config(micro_batches=None, variant=0):
    """The workload both arms run and print (identical dicts: the driver compares them)."""
    return {"workload": "configs[2]: batch-8 synthetic 256^3 volumes + random MANO poses, P=65536, mock latents",
            "images_per_gpu": B_PER_GPU, "D": D, "P": P, "evals_per_step": EVALS_PER_STEP, "step_index": STEP_INDEX,
            "l2": "inputs larger than L2 (0.54 GB of volumes touched per launch, 1.07 GB per evaluation of the batch, vs 126 MB L2)"}


class CpuGuidance:
    """The CPU oracle (the restated reference arithmetic; the reference itself cannot be imported offline,
    SURVEY.md section 8c) on ONE image of the bench workload (D=256, P=65 536), doing what the GPU arm does per
    evaluation: mock decode of x1 = x_t + (1 - sigma) v, energy forward + backward, decoder adjoint, AdamW on the
    16 leaves and on the 196 608-element velocity.  All host threads unless ``threads`` is given."""

    def __init__(self, threads: int = 0, seed: int = 0):
        import torch
        from followmyhold_b200.guidance.loop import LATENT_SHAPE, set_timesteps_sigmas
        from followmyhold_b200.synthetic import make_guidance_sample
        self.torch = torch
        self.cores = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.cores)
        self.s = make_guidance_sample(D, P, seed=seed)
        L = LATENT_SHAPE[0] * LATENT_SHAPE[1]
        g = torch.Generator().manual_seed(1234)
        vol = D ** 3
        starts = torch.randperm(vol // 64, generator=g)[: L // 64].sort().values * 64
        self.tap = (starts.view(-1, 1) + torch.arange(64).view(1, -1)).reshape(-1)
        self.alpha = 0.05
        self.sigma = float(set_timesteps_sigmas(20)[STEP_INDEX])
        self.x_t = torch.randn(L, generator=g)
        self.vel = 0.1 * torch.randn(L, generator=g)
        self.th = torch.cat([self.s.theta_h, self.s.theta_o]).clone()
        self.m, self.v = torch.zeros(16), torch.zeros(16)
        self.vm, self.vv = torch.zeros(L), torch.zeros(L)
        self.n = 0

    def evaluate(self) -> None:
        from oracle import guidance_oracle as O
        torch, s = self.torch, self.s
        vel = self.vel.clone().requires_grad_(True)
        x1 = self.x_t + (1.0 - self.sigma) * vel
        sdf = s.sdf.reshape(-1).index_add(0, self.tap, self.alpha * x1).reshape(s.sdf.shape)
        a = self.th[:8].clone().requires_grad_(True); b = self.th[8:].clone().requires_grad_(True)
        out = O.guidance_energy(sdf, s.hand_rest, s.hand_faces, s.cloud, a, b, s.T_h2m, s.obj_center,
                                j_regressor=s.j_regressor, kps_2d=s.kps_2d, fov_deg=s.fov_deg, image_hw=s.image_hw)
        out["total"].backward()
        self.n += 1
        self.th, self.m, self.v = O.adamw_step(self.th, torch.cat([a.grad, b.grad]), self.m, self.v, self.n, 1e-2)
        self.vel, self.vm, self.vv = O.adamw_step(self.vel, vel.grad, self.vm, self.vv, self.n, 1e-2)


def cpu_guidance_evals_per_sec(seconds_budget: float = 20.0, threads: int = 0, n_min: int = 2):
    cg = CpuGuidance(threads)
    n, t0 = 0, time.perf_counter()
    while True:
        cg.evaluate()
        n += 1
        el = time.perf_counter() - t0
        if (n >= n_min and el >= seconds_budget) or el > 4 * seconds_budget:
            break
    el = time.perf_counter() - t0
    return n / el, cg.cores, n, el


def run_reference(args):
    """Reference arm: the CPU implementation of the path on the host cores.  A "step" here is a bounded SAMPLE of
    a guided-denoise step -- ``--ref-evals`` evaluations of one image out of the 50 x 8 a full step of the batch
    has -- so that K + W steps end within minutes; value extrapolates to whole image-steps per second."""
    import traceback
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pg = None
    if world > 1:
        # the other ranks do no work but stay until rank 0 is done, so the launcher never sees ranks of one job
        # ending minutes apart
        try:
            import datetime
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("gloo", timeout=datetime.timedelta(minutes=30))
            pg = dist
        except Exception as e:          # no rendezvous: rank 0 still measures, the others leave
            print(json.dumps({"impl": "reference", "note": f"rank {rank}: no process group ({type(e).__name__}: {e})"}),
                  file=sys.stderr, flush=True)
    try:
        if rank == 0:
            _reference_rank0(args)
    except Exception as e:
        tb = traceback.format_exc().strip().splitlines()
        print(json.dumps({"impl": "reference", "error": f"{type(e).__name__}: {e}", "traceback": tb[-6:]}), flush=True)
        raise
    finally:
        if pg is not None:
            try:
                pg.barrier()
                pg.destroy_process_group()
            except Exception:
                pass


def _reference_rank0(args):
    K, W = args.steps, args.warmup
    n_evals = max(1, args.ref_evals)
    cg = CpuGuidance()
    times = []
    for i in range(W + K):
        t0 = time.perf_counter()
        for _ in range(n_evals):
            cg.evaluate()
        if i >= W:
            times.append(time.perf_counter() - t0)
    t_total = sum(times)
    evals_per_s = K * n_evals / t_total
    value = evals_per_s / EVALS_PER_STEP            # image-steps per second
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "sampled": True,
        "step_is": f"a bounded sample: {n_evals} evaluation(s) of ONE image of the workload (a full step of the batch is "
                   f"{EVALS_PER_STEP} x {B_PER_GPU}); value = evaluations/s / {EVALS_PER_STEP}; ms_per_step is the measured "
                   "time of one sample",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cg.cores, "kind": "port",
                         "sample": f"{K * n_evals} oracle guidance evaluations (mock decode + fwd + bwd + adjoint + AdamW on leaves "
                                   f"and velocity) of ONE image of the workload in {t_total:.1f}s; value = evals/s / {EVALS_PER_STEP}",
                         "evals_per_sec": evals_per_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# --------------------------------------------------------------------------- decoder leg (row f1)
def decoder_leg(dev, n_images: int = 1, D_lat: int = 65, n_active: int = 8192, reps: int = 3, layers: int = 16):
    """The reference's real inner iteration shape: ``latent2sdf`` of the 65^3 lattice
    (pipelines.py:292-312, 1126-1137) and its adjoint on the tensor cores, random-init weights of the released
    architecture, timed with CUDA events.  Returned as extra keys of the bench line; the roofline here is the
    tensor one: achieved = algorithmic FLOPs / time against the measured cuBLAS bf16 rate."""
    import torch
    from followmyhold_b200.decoder import tc
    from followmyhold_b200.decoder.shapevae import (DecoderWeights, LatentDecoder, adjoint_flops, decode_flops, lattice_points,
                                                    random_state_dict)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_sus = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_burst = float(peaks.get("bf16_tflops", 1590.0))
    kind = "measured" if peaks else "fallback"
    W = DecoderWeights(random_state_dict(layers, seed=0), dev)
    dec = LatentDecoder(W, n_images, device=dev)
    dec.set_queries(lattice_points(D_lat))
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(n_images, 3072, 64, generator=g).to(dev)
    Nq = D_lat ** 3
    idx = torch.randint(0, Nq, (n_images, n_active), generator=g).to(torch.int32).to(dev)
    gs = (1e-2 * torch.randn(n_images, n_active, generator=g)).to(dev)
    sdf = torch.empty(n_images, Nq, device=dev)
    gl = torch.empty(n_images, 3072, 64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n):
        fn(); torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    fwd_ms = timed(lambda: dec.forward(lat, out=sdf), reps)
    bwd_ms = timed(lambda: dec.backward(idx, gs, out=gl), reps)
    assert torch.isfinite(sdf).all() and torch.isfinite(gl).all()
    fl = decode_flops(Nq, layers)
    fwd_tf = n_images * fl["forward"] / fwd_ms / 1e9
    bwd_tf = n_images * adjoint_flops(n_active, layers) / bwd_ms / 1e9
    # the two tensor-core kernels alone, at the shapes the decode launches them with
    n = min(32768, Nq) // 128 * 128
    q = torch.randn(n, 16, 64, device=dev).half()
    kv = torch.randn(n_images * 3072, 16, 128, device=dev).half()
    o = torch.empty(n_images, n, 1024, dtype=torch.float16, device=dev)
    att_ms = timed(lambda: tc.attention(q, kv[:, :, :64], kv[:, :, 64:], n_images, out=o, q_shared=True), 5)
    att_tf = 4.0 * n_images * n * 3072 * 1024 / att_ms / 1e9
    a = torch.randn(n_images * n, 1024, device=dev).half()
    wfc = torch.randn(4096, 1024, device=dev).half()
    u = torch.empty(n_images * n, 4096, dtype=torch.float16, device=dev)
    bias = torch.zeros(4096, device=dev)
    gemm_ms = timed(lambda: tc.gemm(a, wfc, out=u, bias=bias, act=tc.ACT_GELU), 5)
    gemm_tf = 2.0 * n_images * n * 4096 * 1024 / gemm_ms / 1e9
    return {
        "workload": f"latent2sdf + adjoint, {n_images} image(s), {D_lat}^3 = {Nq} lattice queries, 3072 x 64 latents, {layers}-layer "
                    f"ShapeVAE decoder (random-init weights), {n_active} lattice points carrying a gradient",
        "dtype": "f16 operands, f32 accumulation (the reference decodes in fp16, pipelines.py:302-306)",
        "forward_ms": fwd_ms, "adjoint_ms": bwd_ms, "forward_tflops": fwd_tf, "adjoint_tflops": bwd_tf,
        "flops_forward_per_image": fl, "flops_adjoint_per_image": adjoint_flops(n_active, layers),
        "decodes_per_sec": n_images * 1e3 / fwd_ms, "evaluations_per_sec_decoder_only": n_images * 1e3 / (fwd_ms + bwd_ms),
        "roofline": {"bound": "tensor", "kernel": "latent2sdf forward (k_attn_fwd2 + k_gemm_tc + row kernels)", "achieved": fwd_tf,
                     "peak": peak_sus, "peak_kind": f"{kind} cuBLAS bf16, sustained (kernels timed inside a long step)",
                     "unit": "TFLOP/s", "frac": fwd_tf / peak_sus, "traffic": None},
        "kernels_alone": {
            "k_gemm_tc": {"shape": [n_images * n, 4096, 1024], "epilogue": "bias + GELU", "ms": gemm_ms, "tflops": gemm_tf,
                          "frac_of_burst_peak": gemm_tf / peak_burst, "peak": peak_burst},
            "k_attn_fwd2": {"queries": n, "keys": 3072, "heads": 16, "images": n_images, "ms": att_ms, "tflops": att_tf,
                            "frac_of_burst_peak": att_tf / peak_burst, "peak": peak_burst}},
    }


# --------------------------------------------------------------------------- the reference's own lattices (extra keys)
def icp_leg(dev, cpu_seconds: float = 2.0):
    """Row a16 (iterations/s and alignments/s, no roofline claim): the reference's two ICP stages (coarse 50
    iterations 1 000 / 5 000 points, fine 100 iterations 5 000 / 10 000, 20 % trimmed, scale clip [0.7, 3], h2m.py:35-54) on
    synthetic clouds, point sets resident, CUDA events; the CPU oracle (scipy cKDTree + numpy, one thread like the
    reference) on a bounded number of iterations beside it."""
    import ctypes as C
    import numpy as np
    import torch
    from followmyhold_b200 import _lib
    from oracle import icp_oracle as IO
    lib = _lib.load()
    rng = np.random.default_rng(0)
    out = {}
    total_ms = 0.0
    for name, n_iter, ns, nt in (("coarse", 50, 1000, 5000), ("fine", 100, 5000, 10000)):
        tgt = rng.normal(size=(nt, 3)) * np.array([1.0, 0.6, 0.3])
        src = (tgt[rng.choice(nt, ns, replace=False)] - 0.05) / 1.15 + 0.002 * rng.normal(size=(ns, 3))
        d_src, d_tgt = torch.as_tensor(src).to(dev), torch.as_tensor(tgt).to(dev)
        nbytes = lib.foho_icp_workspace_bytes(ns, nt)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        ws_ptr = ws.data_ptr() + ((-ws.data_ptr()) % 256)
        T = torch.zeros(16, dtype=torch.float64, device=dev)
        cost = torch.zeros(1, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev)

        def call():
            _lib.check("foho_icp_run", lib.foho_icp_run(d_src.data_ptr(), ns, d_tgt.data_ptr(), nt, n_iter, int(0.2 * ns), 0, 0.7, 3.0,
                                                        T.data_ptr(), cost.data_ptr(), None, None, C.c_void_p(ws_ptr), nbytes,
                                                        C.c_void_p(st.cuda_stream)))
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        total_ms += ms
        k = max(2, min(n_iter, int(cpu_seconds / (0.004 if name == "fine" else 0.001)) // 8))
        t0 = time.perf_counter()
        IO.icp_points(src, tgt, k, int(0.2 * ns), False, 0.7, 3.0)
        cpu_ms_per_iter = (time.perf_counter() - t0) / k * 1e3
        out[name] = {"n_iter": n_iter, "Ns": ns, "Nt": nt, "ms": ms, "iterations_per_sec": n_iter / ms * 1e3,
                     "pair_evals_per_sec": ns * nt * n_iter / ms * 1e3, "cpu_oracle_ms_per_iter": cpu_ms_per_iter,
                     "cpu_sample": f"{k} iterations, 1 thread"}
    out["alignments_per_sec"] = 1e3 / total_ms
    out["note"] = "one alignment = coarse + fine run (mesh_align.py:178-217), one persistent launch each; sampling and file I/O not included"
    return out


def reference_lattice_leg(dev, steps: int = 10):
    """Not the headline: the same loop at the reference's real shapes -- B = 8 images on the 65^3
    lattice with the largest cloud (512^2 crop = 262 144 points), 50 evaluations per step, mock latents; and one
    evaluation on the 385^3 export lattice.  Both go through the generic stream kernel (`k_stream_any`: D is odd)."""
    import torch
    from followmyhold_b200.guidance.engine import GuidanceEngine
    from followmyhold_b200.guidance.loop import GuidanceLoop
    from followmyhold_b200.synthetic import make_guidance_sample, stack_samples
    out = {}
    B, D65, P65 = 8, 65, 262144
    samples = [make_guidance_sample(D65, P65, seed=900 + i) for i in range(B)]
    sdf0, theta0, st = stack_samples(samples, device=dev, cap=True)
    loop = GuidanceLoop(B, D65, st, P65, device=dev, micro_batches=2)
    g = torch.Generator().manual_seed(5)
    loop.sdf0.copy_(sdf0); loop.sdf.copy_(sdf0); loop.theta.copy_(theta0)
    loop.x_t.copy_(torch.randn(B, loop.L, generator=g)); loop.velocity.copy_(0.1 * torch.randn(B, loop.L, generator=g))
    loop.capture(STEP_INDEX)
    for _ in range(3):
        loop.run_step_device(STEP_INDEX)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loop.run_step_device(STEP_INDEX)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert torch.isfinite(loop.terms).all()
    out["ref65"] = {"workload": "B=8, D=65 (pipelines.py:1126-1137), P=262144, 50 evaluations per step, mock latents",
                    "image_steps_per_sec": B * 1e3 / ms, "ms_per_step": ms, "evals_per_sec": B * EVALS_PER_STEP * 1e3 / ms,
                    "us_per_evaluation_of_the_batch": 1e3 * ms / EVALS_PER_STEP,
                    "launches_per_evaluation": loop.kernels_per_eval() * loop.micro_batches,
                    "note": "1.1 MB of volume per image: launch / latency bound, not HBM bound"}
    del loop
    D385, P385 = 385, 65536
    s = make_guidance_sample(D385, P385, seed=950)
    sdf, theta, st1 = stack_samples([s], device=dev, cap=True)
    eng = GuidanceEngine(1, D385, 778, st1.hand_faces.shape[0], P385, device=dev)
    eng.prepare(st1)
    for _ in range(3):
        eng.energy_fwd_bwd(sdf, theta, st1)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        eng.energy_fwd_bwd(sdf, theta, st1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    peak, _ = measured_peak_gbs()
    out["export385"] = {"workload": "one 385^3 volume (the export lattice, pipelines.py:1624-1639), P=65536, one evaluation",
                        "eval_ms": ms, "GBps_algorithmic": 8 * D385 ** 3 / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 8 * D385 ** 3 / (ms * 1e-3) / 1e9 / peak,
                        "stream_kernel": "k_stream_any"}
    return out


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from followmyhold_b200 import _lib
    from followmyhold_b200.guidance.loop import GuidanceLoop
    from followmyhold_b200.synthetic import cap_boundary_loops, make_guidance_sample, stack_samples

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep rank 0's stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if not torch.cuda.is_available():
        raise _lib.FohoLibraryError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from followmyhold_b200.parallel import bind_to_gpu_numa
    orig_affinity = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(local)      # before the pinned staging buffers exist
    _lib.load()
    K, W = args.steps, max(3, args.warmup)
    B = B_PER_GPU

    # ---- synthetic inputs: images rank*B .. rank*B+B-1 (independent units, no exchange: SURVEY.md §8e)
    samples = [make_guidance_sample(D, P, seed=rank * B + i) for i in range(B)]
    sdf0, theta0, st = stack_samples(samples, device=dev, cap=True)
    loop = GuidanceLoop(B, D, st, P, device=dev, stream_variant=args.variant, micro_batches=args.micro_batches)
    g = torch.Generator().manual_seed(1234 + rank)
    x_t_h = torch.randn(B, loop.L, generator=g).pin_memory()
    vel_h = (0.1 * torch.randn(B, loop.L, generator=g)).pin_memory()
    theta_h = theta0.cpu().pin_memory()
    sdf0_h = torch.empty(sdf0.shape, dtype=sdf0.dtype, pin_memory=True)
    sdf0_h.copy_(sdf0)

    def load_device_state():
        loop.sdf0.copy_(sdf0); loop.sdf.copy_(sdf0)
        loop.x_t.copy_(x_t_h); loop.velocity.copy_(vel_h); loop.theta.copy_(theta0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    load_device_state()
    loop.capture(STEP_INDEX)
    for _ in range(W):
        loop.run_step_device(STEP_INDEX)
    torch.cuda.synchronize()

    # ---- timed region: K graph replays, inputs resident in HBM.  The 1.07 GB of volumes
    #      touched per evaluation exceeds the 126 MB L2, so no L2 flush is needed between steps.
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loop.run_step_device(STEP_INDEX)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    terms = loop.terms.cpu()
    assert torch.isfinite(terms).all(), "non-finite guidance terms in the timed region"

    # ---- end to end through the host-buffer API (H2D of the step's inputs + D2H of its results inside)
    #      K2 batches of B images go through the public host-buffer API in one call: upload of batch k+1
    #      and download of batch k-1 overlap the graph replay of batch k (three streams); every batch's
    #      549 MB still crosses PCIe inside the timed region.
    K2 = max(2, K)           # as many batches as timed steps; the first upload cannot be hidden and is inside
    batch = (sdf0_h, x_t_h, vel_h, theta_h)
    loop.denoise_steps_host(STEP_INDEX, [batch] * K2)      # warm-up: same length, so every pinned buffer exists
    barrier()
    t0 = time.perf_counter()
    outs = loop.denoise_steps_host(STEP_INDEX, [batch] * K2)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert all(torch.isfinite(o["terms"]).all() for o in outs), "non-finite terms in the end-to-end run"
    # same, with the volumes (the mock decoder's state) resident and only latents / model output / leaves
    # uploaded per step -- the traffic of the real loop, where the decoder produces the volume on the device
    lat = (None, x_t_h, vel_h, theta_h)
    loop.denoise_steps_host(STEP_INDEX, [batch] + [lat] * (K2 - 1))
    barrier()
    t2 = time.perf_counter()
    loop.denoise_steps_host(STEP_INDEX, [lat] * K2)
    barrier()
    e2e_lat_s = time.perf_counter() - t2
    # same, with the volumes uploaded in the dtype the reference's decoder emits them (fp16 logits, widened by
    # `.float()`, pipelines.py:303-309) and widened on the device: half the PCIe bytes.  No gain on one GPU (the
    # path is not PCIe bound there); it matters once several GPUs share the host's PCIe fabric.
    sdf0_h16 = torch.empty(sdf0.shape, dtype=torch.float16, pin_memory=True)
    sdf0_h16.copy_(sdf0)
    b16 = (sdf0_h16, x_t_h, vel_h, theta_h)
    loop.denoise_steps_host(STEP_INDEX, [b16] * K2)
    barrier()
    t3 = time.perf_counter()
    loop.denoise_steps_host(STEP_INDEX, [b16] * K2)
    barrier()
    e2e_h16_s = time.perf_counter() - t3
    # one batch alone (no overlap possible): the latency a single call sees
    t1 = time.perf_counter()
    loop.denoise_steps_host(STEP_INDEX, [lat])
    e2e_single_s = time.perf_counter() - t1

    # ---- dominant kernel alone: the dense stream of one micro-batch (stage_mask = prep|stream), CUDA events
    #      on its stream.  One launch streams the nb = B / micro_batches volumes of its lane.
    ln = loop.lanes[0]
    eng, nb = ln.engine, ln.nb
    sdf_l, theta_l = loop.sdf.narrow(0, ln.off, nb), loop.theta.narrow(0, ln.off, nb)
    desc = eng.make_desc(sdf_l, theta_l, ln.statics)
    desc.stage_mask = 1
    eng.launch(desc)
    desc.stage_mask = 2
    NREP = 50
    for _ in range(5):
        eng.launch(desc)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(NREP):
        eng.launch(desc)
    s1.record()
    torch.cuda.synchronize()
    stream_ms = s0.elapsed_time(s1) / NREP
    # the stream kernels of all lanes at once, each on its lane's stream -- how they meet in the step graph
    pair_ms = None
    if len(loop.lanes) > 1:
        descs = []
        for l2 in loop.lanes:
            d2 = l2.engine.make_desc(loop.sdf.narrow(0, l2.off, l2.nb), loop.theta.narrow(0, l2.off, l2.nb), l2.statics)
            d2.stage_mask = 1
            l2.engine.launch(d2)
            d2.stage_mask = 2
            descs.append(d2)
        cur = torch.cuda.current_stream()
        lane_streams = [cur] + [l2.stream for l2 in loop.lanes[1:]]
        torch.cuda.synchronize()
        s0.record()
        for _ in range(NREP):
            for l2, d2, ls in zip(loop.lanes, descs, lane_streams):
                ls.wait_stream(cur) if ls is not cur else None
                l2.engine.launch(d2, ls)
            for ls in lane_streams[1:]:
                cur.wait_stream(ls)
        s1.record()
        torch.cuda.synchronize()
        pair_ms = s0.elapsed_time(s1) / NREP
    # whole evaluation of that micro-batch (all 11 kernels, sparse chains beside the stream), same method
    desc.stage_mask = 0
    for _ in range(3):
        eng.launch(desc)
    torch.cuda.synchronize()
    s0.record()
    for _ in range(NREP):
        eng.launch(desc)
    s1.record()
    torch.cuda.synchronize()
    eval_ms = s0.elapsed_time(s1) / NREP
    # the same evaluation with every kernel in series on one stream (what a profiler's serialised launch
    # list shows): the stream kernel's share of THAT is the number to hold against the ncu launch list
    eng.serial = 1
    desc_s = eng.make_desc(sdf_l, theta_l, ln.statics)
    eng.serial = 0
    for _ in range(3):
        eng.launch(desc_s)
    torch.cuda.synchronize()
    s0.record()
    for _ in range(NREP):
        eng.launch(desc_s)
    s1.record()
    torch.cuda.synchronize()
    eval_serial_ms = s0.elapsed_time(s1) / NREP

    # ---- max over ranks
    t = torch.tensor([ms, e2e_s, stream_ms, eval_ms, e2e_lat_s, e2e_h16_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s, stream_ms, eval_ms, e2e_lat_s, e2e_h16_s = [float(x) for x in t.tolist()]

    if rank == 0:
        value = world * B * K / (ms / 1e3)
        e2e_value = world * B * K2 / e2e_s
        peak, peak_kind = measured_peak_gbs()
        stream_bytes = nb * 8 * D ** 3                     # dense stream: 4 B read + 4 B written per voxel
        achieved = stream_bytes / (stream_ms * 1e-3) / 1e9
        eval_bytes = nb * algorithmic_bytes_per_eval(D, P)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(),
            "detail": {"micro_batches": args.micro_batches, "images_per_launch": nb,
                       "stream_variant": {0: "tma", 1: "ldg", 2: "tma"}.get(args.variant, "tma"),
                       "evals_per_sec": value * EVALS_PER_STEP, "eval_ms_standalone": eval_ms,
                       "eval_ms_serialised": eval_serial_ms, "host_numa": numa,
                       "eval_GBps_algorithmic": eval_bytes / (eval_ms * 1e-3) / 1e9},
            "clocks": clocks,
            # End to end = the call the real loop makes once per denoise step: the step's inputs (latents, model
            # output, leaves) come from pinned host memory, its results (optimised model output, prev_sample, leaves,
            # loss terms) go back -- all inside the timed region.  The volumes are the DECODER's output and are
            # produced on the device (latent2sdf runs there, row f1), so they are not a per-step input; the
            # round-1 variant that also uploads 512 MiB of volumes per step is kept as `volumes_uploaded`.
            "e2e": {"value": world * B * K2 / e2e_lat_s, "unit": UNIT,
                    "h2d_bytes_per_step": 4 * (2 * B * loop.L + B * 16),
                    "d2h_bytes_per_step": loop.d2h_bytes_per_step(), "steps": K2,
                    "api": "GuidanceLoop.denoise_steps_host (pinned host buffers in and out, 3-stream pipeline); "
                           "inputs per step: latents x_t, model output, leaves",
                    "single_batch_ms": e2e_single_s * 1e3,
                    "volumes_uploaded": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": loop.h2d_bytes_per_step(),
                                         "note": "round-1 definition: the decoder base volumes (512 MiB fp32) also cross "
                                                 "PCIe every step"},
                    "volumes_uploaded_fp16": {"value": world * B * K2 / e2e_h16_s, "unit": UNIT,
                                              "h2d_bytes_per_step": 2 * B * D ** 3 + 4 * (2 * B * loop.L + B * 16)}},
            "gpu_launches": K * loop.launches_per_step(),
            "roofline": {"bound": "hbm", "kernel": "k_stream_tma" if args.variant in (0, 2) else "k_stream_ldg",
                         "achieved": achieved, "peak": peak, "peak_kind": f"{peak_kind} burst copy (kernel timed alone)",
                         "unit": "GB/s", "frac": achieved / peak, "bytes_per_launch": stream_bytes,
                         "ms_per_launch": stream_ms, "traffic": ncu_traffic_bytes(),
                         "all_lanes_concurrent": (None if pair_ms is None else
                                                  {"ms": pair_ms, "achieved": B * 8 * D ** 3 / (pair_ms * 1e-3) / 1e9,
                                                   "frac": B * 8 * D ** 3 / (pair_ms * 1e-3) / 1e9 / peak}),
                         "share_of_eval": stream_ms / eval_ms,
                         "share_of_eval_serialised": stream_ms / eval_serial_ms,
                         "note": "the sparse kernels run beside the stream kernel (fork/join), so it covers "
                                 "share_of_eval of the evaluation's wall time; with every kernel in series "
                                 "(as ncu replays them) its share is share_of_eval_serialised"},
        }
        if world == 1 and not args.no_decoder:
            try:
                line["decoder"] = decoder_leg(dev)
                d8 = decoder_leg(dev, n_images=8, reps=2)
                line["decoder"]["batch8"] = {k: d8[k] for k in ("workload", "forward_ms", "adjoint_ms", "forward_tflops", "adjoint_tflops",
                                                                 "decodes_per_sec", "evaluations_per_sec_decoder_only")}
                line["decoder"]["batch8"]["frac_forward"] = d8["roofline"]["frac"]
            except Exception as e:          # the headline line must survive a failure of this extra leg
                line.setdefault("decoder", {})["error"] = f"{type(e).__name__}: {e}"
            torch.cuda.empty_cache()
            try:
                line["reference_lattices"] = reference_lattice_leg(dev)
            except Exception as e:
                line["reference_lattices"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
            try:
                line["icp"] = icp_leg(dev)
            except Exception as e:
                line["icp"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, orig_affinity)      # the CPU leg gets every host core back
            eps, cores, n, el = cpu_guidance_evals_per_sec(args.cpu_seconds)
            eps1, _, n1, el1 = cpu_guidance_evals_per_sec(min(8.0, args.cpu_seconds), threads=1, n_min=1)
            line["cpu_baseline"] = {"value": eps / EVALS_PER_STEP, "unit": UNIT, "cores": cores, "kind": "port",
                                    "one_thread": {"value": eps1 / EVALS_PER_STEP, "unit": UNIT, "cores": 1,
                                                   "sample": f"{n1} evaluations in {el1:.1f}s; the reference pins its CPU "
                                                             "stages to one thread (src/foho/main.py:65-68)"},
                                    "sample": f"{n} oracle guidance evaluations (fwd+bwd+AdamW) of ONE image of the "
                                              f"workload in {el:.1f}s; value = evals/s / {EVALS_PER_STEP}",
                                    "evals_per_sec": eps}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200, help="timed steps (default: ~2 s of device time)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="dense stream kernel: 0/2 = TMA bulk, 1 = LDG")
    ap.add_argument("--micro-batches", type=int, default=2,
                    help="groups of images that advance independently inside the step's graph (1, 2 or 4)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-decoder", action="store_true", help="skip the tensor-core decoder leg (row f1)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-evals", type=int, default=3, help="--impl reference: oracle evaluations per (sampled) step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
