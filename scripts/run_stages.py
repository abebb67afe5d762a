#!/usr/bin/env python
"""BASELINE.json configs[3] / configs[4] in synthetic form: a batch of frames goes through the three
stages this package replaces -- ``alignment.h2m.run`` -> ``alignment.mano.run`` -> ``guidance.run`` -- by
their reference-facing ``run(...)`` entry points, on file sets shaped like the reference's (SURVEY.md
section 8b), sharded ``sorted(images)[rank::world]`` under torchrun.  The real frames (OakInk / DexYCB / ARCTIC)
and the Hunyuan3D networks are not available offline: frames are synthetic, the networks are the
``MockGuidanceModel`` (a linear tap decoder), so the numbers say what the stages cost around the networks.

  python scripts/run_stages.py --images 32 [--steps 20] [--out gpurun_out/stages.json]
  python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 scripts/run_stages.py --images 32
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def write_frames(root: str, n_images: int, D: int, P: int) -> dict:
    """Outputs of the stages upstream of alignment + guidance for ``n_images`` synthetic frames."""
    import cv2
    import torch
    from followmyhold_b200.meshio import write_obj, write_ply
    from followmyhold_b200.synthetic import icosphere, make_guidance_sample, standin_j_regressor
    d = {k: os.path.join(root, k) for k in ("img", "mask", "moge", "hun", "hamer", "h2m", "mano", "out")}
    for v in d.values():
        os.makedirs(v, exist_ok=True)
    ov, of = icosphere(3, 1.0)
    rng = np.random.default_rng(0)
    for k in range(n_images):
        i = f"{k:03d}"
        s = make_guidance_sample(16, 64, 500 + k)          # hand pose, T_h2m, key-points; the volume is the model's
        open(os.path.join(d["img"], f"{i}_cropped_obj_1.png"), "wb").close()
        m = np.zeros((64, 64), np.uint8); m[20:40, 20:40] = 255
        cv2.imwrite(os.path.join(d["mask"], f"{i}_cropped_hand_mask.png"), m)
        cv2.imwrite(os.path.join(d["mask"], f"{i}_cropped_obj_mask.png"), m)
        md = os.path.join(d["moge"], f"{i}_cropped_hoi"); os.makedirs(md, exist_ok=True)
        json.dump({"fov_x": 41.0}, open(os.path.join(md, "fov.json"), "w"))
        T = s.T_h2m.numpy().astype(np.float64)
        Ainv = np.linalg.inv(T[:3, :3])
        hand_hun = (s.hand_rest.numpy().astype(np.float64) - T[:3, 3]) @ Ainv.T          # MoGe -> Hunyuan
        faces = s.hand_faces.numpy()
        obj_hun = ov.astype(np.float64) * np.array([0.5, 0.4, 0.3]) * (0.9 + 0.02 * (k % 5))
        # Hunyuan HOI mesh = hand + object, one vertex list (geometry/hunyuan.py writes {i}_hoi_mesh.ply)
        hoi_v = np.concatenate([hand_hun, obj_hun])
        hoi_f = np.concatenate([faces, of + len(hand_hun)])
        write_ply(os.path.join(d["hun"], f"{i}_hoi_mesh.ply"), hoi_v, hoi_f)
        # MoGe sees the same scene through T_h2m: surface samples of the HOI mesh, mildly noisy
        tri = hoi_v[hoi_f]
        fi = rng.integers(0, len(hoi_f), max(P, 20000))
        bc = rng.random((len(fi), 2)); flip = bc.sum(1) > 1; bc[flip] = 1 - bc[flip]
        pts = tri[fi, 0] + bc[:, :1] * (tri[fi, 1] - tri[fi, 0]) + bc[:, 1:] * (tri[fi, 2] - tri[fi, 0])
        cloud = pts @ T[:3, :3].T + T[:3, 3] + 0.002 * rng.standard_normal(pts.shape)
        write_ply(os.path.join(md, "pointcloud.ply"), cloud)
        # HaMeR's hand: the same hand in its own frame and scale (hand/hamer.py writes {i}_hamer.obj)
        write_obj(os.path.join(d["hamer"], f"{i}_hamer.obj"), hand_hun / 1.25 + np.array([0.05, -0.02, 0.4]), faces)
        np.save(os.path.join(d["hamer"], f"{i}_kps_for_guidance.npy"),
                {"mano_2d_kps": s.kps_2d.numpy() / 8.0, "mano_3d_kps": np.zeros((21, 3)), "cam_t": np.zeros(3)},
                allow_pickle=True)
    torch.save(torch.from_numpy(standin_j_regressor(0)), os.path.join(root, "J_regressor_hamer.pt"))
    return d


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=32)
    ap.add_argument("--batch-size", type=int, default=8)
    ap.add_argument("--D", type=int, default=64, help="lattice of the mock decoder (the reference loop uses 65)")
    ap.add_argument("--P", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=20, help="num_inference_steps (reference 20; BASELINE configs 50)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--root", default=None, help="where to write the synthetic frames (default: a temp dir)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from followmyhold_b200 import _lib
    from followmyhold_b200.alignment import h2m, mano
    from followmyhold_b200.guidance import run as G
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.parallel import aggregate_throughput, gather_timings

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise _lib.FohoLibraryError("run_stages.py needs a GPU; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    root = args.root or os.path.join(tempfile.gettempdir(), "foho_stage_frames")
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
        d = write_frames(root, args.images, args.D, args.P)
    if world > 1:
        dist.barrier()
    d = {k: os.path.join(root, k) for k in ("img", "mask", "moge", "hun", "hamer", "h2m", "mano", "out")}
    mine = len(sorted(os.listdir(d["img"]))[rank::world])

    def timed(fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    t_h2m = timed(lambda: h2m.run(d["hun"], d["moge"], d["h2m"], device=dev, concurrent=args.batch_size))
    t_mano = timed(lambda: mano.run(d["hamer"], d["hun"], d["mano"], device=dev, concurrent=args.batch_size))
    cfg = OptimizationConfig().with_steps(args.steps)
    model = G.MockGuidanceModel(D=args.D, latent_elems=3072 * 64)
    t_guid = timed(lambda: G.run("/nonexistent", d["img"], d["mask"], d["moge"], d["hun"], d["hamer"], d["h2m"], d["mano"],
                                 d["out"], model=model, batch_size=args.batch_size, n_cloud=args.P, device=dev, config=cfg,
                                 j_regressor_path=os.path.join(root, "J_regressor_hamer.pt")))
    written = len([f for f in os.listdir(d["out"]) if f.endswith("_obj.ply")])
    res = {}
    for name, t in (("alignment_h2m", t_h2m), ("alignment_mano", t_mano), ("guidance", t_guid),
                    ("all_stages", t_h2m + t_mano + t_guid)):
        per_rank = gather_timings({"units": float(mine), "seconds": t}, device=dev if world > 1 else None)
        res[name] = {"images_per_sec": aggregate_throughput(per_rank), "seconds_max_rank": max(r["seconds"] for r in per_rank)}
    if world > 1:
        dist.barrier()
    if rank == 0:
        evals = cfg.optimization_steps_hand + cfg.optimization_steps_scale + \
            (cfg.guidance_end_step - cfg.handopt_start_step - 2) * cfg.optimization_steps_joint
        line = {"what": "alignment + guidance stages through their run() entry points, synthetic frames, mock networks",
                "n_gpus": world, "images": args.images, "batch_size": args.batch_size, "D": args.D, "P": args.P,
                "num_inference_steps": args.steps, "guidance_evaluations_per_image": evals,
                "icp_iterations_per_image": 2 * (50 + 100), "obj_meshes_written": written, "stages": res,
                "note": "wall clock incl. file I/O, host-side sampling and CUDA-graph capture (one graph per "
                        "optimised denoise step and batch shape); max over ranks"}
        print(json.dumps(line), flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            json.dump(line, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
