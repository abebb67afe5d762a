"""Guidance stage: same ``run(...)`` signature, CLI flags, per-index file contract, skip rules and
error behaviour as the reference stage ``src/foho/guidance/run.py:188-289`` -- with the guided
denoise loop of ``Hunyuan3DDiTFlowMatchingPipeline_main.__call__``
(third_party_patches/hy3dgen/shapegen/pipelines.py:1262-1612) running through ``GuidanceLoop``
(batched over images, one CUDA graph per denoise step) instead of one image at a time in PyTorch.

What this stage does NOT contain are the two networks of the reference loop, which are outside the
hot path this package replaces: the Hunyuan3D DiT that predicts the flow velocity at
every step (pipelines.py:1262-1291) and the ShapeVAE that decodes latents to the volume
(``latent2sdf``, :292-338).  They enter through a ``GuidanceModel`` object (protocol below); the
reference's own networks would be wrapped in one.  Without a model the stage fails loudly -- there
is no fallback.  ``MockGuidanceModel`` (synthetic volumes, the linear tap decoder of
``foho_mock_decoder_*``) exists for tests and synthetic runs.

Differences a maintainer should know:
  * the MoGe geometry ``{i}_cropped_hoi/mesh.glb`` (run.py:215; else ``pointcloud.ply`` / ``mesh.ply``)
    is consumed as a point cloud -- its vertices -- where the reference renders it;
  * images are processed ``batch_size`` at a time; under ``torchrun`` rank r takes
    ``sorted(images)[r::world]`` (or chunk r of ``task_list_file``, like ``SLURM_ARRAY_TASK_ID``);
  * the mesh post-processors of run.py:158-161 (``FloaterRemover``, ``DegenerateFaceRemover``,
    ``FaceReducer``) are ``followmyhold_b200.meshproc``'s (hy3dgen / MeshLab are not needed).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from ..meshio import TriMesh, load, write_ply
from ..meshproc import DegenerateFaceRemover, FaceReducer, FloaterRemover
from ..parallel import rank_world, shard_images, task_chunk
from ..synthetic import cap_boundary_loops, quat_to_mat_np
from .config import OptimizationConfig
from .engine import GuidanceStatics

J_REGRESSOR_PATH = "./third_party/estimator/hamer/J_regressor_hamer.pt"     # cwd-relative, pipelines.py:1218
MODEL_ENV = "FOHO_B200_GUIDANCE_MODEL"                                      # "package.module:factory"


class GuidanceModel:
    """What the stage needs from the networks (duck-typed; subclassing is optional).

    ``D``: lattice points per axis of the decoded volume (reference 65); ``latent_elems``: 3072*64.
    All tensors live on ``device``."""
    D: int
    latent_elems: int

    def begin_batch(self, indices: Sequence[str], image_paths: Sequence[str], device) -> None:
        """Condition on the batch's images (pipelines.py:1204-1260: image encoder, CFG setup)."""
        raise NotImplementedError

    def initial_latents(self, batch: int, generator: torch.Generator) -> torch.Tensor:
        """x_T [B, latent_elems] (``prepare_latents``, pipelines.py:700)."""
        raise NotImplementedError

    def predict(self, step: int, x_t: torch.Tensor) -> torch.Tensor:
        """The DiT's classifier-free-guided flow velocity for this step, [B, latent_elems] (:1262-1291)."""
        raise NotImplementedError

    def decoder_state(self):
        """(sdf0 [B,D,D,D], tap int64 [latent_elems], alpha) of the linear tap decoder the graph-captured loop
        drives (``GuidanceLoop.run_schedule_device``).  A model that has a ``decode`` method is driven through
        ``GuidanceLoop.run_schedule_decoder`` instead and need not implement this."""
        raise NotImplementedError

    # Optional: ``decode(x1 [B, latent_elems]) -> sdf [B,D,D,D]`` float32, negative inside, differentiable torch
    # ops (``latent2sdf``, pipelines.py:292-312).  The reference's VAE + geo_decoder are wrapped this way.

    def extract_mesh(self, sdf: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Surface of one decoded volume [D,D,D] (negative inside) -> (verts in Hunyuan space, faces)."""
        raise NotImplementedError


def load_model_from_env() -> Optional[GuidanceModel]:
    spec = os.environ.get(MODEL_ENV)
    if not spec:
        return None
    mod, _, fn = spec.partition(":")
    return getattr(importlib.import_module(mod), fn or "make_model")()


# --------------------------------------------------------------------------- per-index inputs
def index_paths(cropped_obj_img: str, cropped_obj_img_dir: str, mask_dir: str, moge_out_dir: str,
                hunyuan_hoi_mesh_dir: str, hamer_out_dir: str, h2m_rt_dir: str, aligned_mano_dir: str,
                guidance_out_dir: str) -> dict:
    """File names of one image, exactly as the reference derives them (run.py:210-222)."""
    index = cropped_obj_img.split("_")[0]
    moge_dir = os.path.join(moge_out_dir, f"{index}_cropped_hoi")
    return dict(
        index=index,
        cropped_obj_img_path=os.path.join(cropped_obj_img_dir, cropped_obj_img),
        cropped_hand_mask_path=os.path.join(mask_dir, f"{index}_cropped_hand_mask.png"),
        cropped_obj_mask_path=os.path.join(mask_dir, f"{index}_cropped_obj_mask.png"),
        moge_dir=moge_dir,
        moge_fov_path=os.path.join(moge_dir, "fov.json"),
        T_h2m_path=os.path.join(h2m_rt_dir, f"{index}_hoi_mesh.npy"),
        aligned_mano_mesh_path=os.path.join(aligned_mano_dir, f"{index}_hamer_aligned_mano.ply"),
        hunyuan_hoi_mesh_path=os.path.join(hunyuan_hoi_mesh_dir, f"{index}_hoi_mesh.ply"),
        hamer_for_guid_path=os.path.join(hamer_out_dir, f"{index}_kps_for_guidance.npy"),
        save_path_obj=os.path.join(guidance_out_dir, f"{index}_obj.ply"),
        save_path_hand=os.path.join(guidance_out_dir, f"{index}_hand.ply"),
    )


def _read_mask(path: str) -> np.ndarray:
    import cv2
    m = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if m is None:
        raise FileNotFoundError(path)
    return m


def load_image_inputs(p: dict, n_cloud: int, rng: np.random.Generator) -> dict:
    """Everything ``__call__`` loads per image before its loop (pipelines.py:1218-1256), as arrays."""
    with open(p["moge_fov_path"], "r", encoding="utf-8") as f:
        fovx = float(json.load(f)["fov_x"])                                        # run.py:228-230
    hand_mask = _read_mask(p["cropped_hand_mask_path"])
    obj_mask = _read_mask(p["cropped_obj_mask_path"])
    if hand_mask.max() == 0 or obj_mask.max() == 0:                                # run.py:234-236
        return {"skip": "empty mask"}
    H, W = hand_mask.shape[:2]                                                     # run.py:80-82
    hamer = np.load(p["hamer_for_guid_path"], allow_pickle=True).item()            # pipelines.py:1219-1220
    kps = np.asarray(hamer["mano_2d_kps"], dtype=np.float32).reshape(21, 2)
    mano = load(p["aligned_mano_mesh_path"])                                       # :1223 (Hunyuan space)
    if not isinstance(mano, TriMesh):
        raise ValueError(f"{p['aligned_mano_mesh_path']} has no faces")
    T = np.load(p["T_h2m_path"]).astype(np.float64).reshape(4, 4)                  # :1240
    hand_moge = mano.vertices.astype(np.float64) @ T[:3, :3].T + T[:3, 3]          # :1241 transform_hunyuan2moge
    cloud_path = None
    for name in ("mesh.glb", "pointcloud.ply", "mesh.ply"):       # mesh.glb is what the reference loads (run.py:215)
        if os.path.isfile(os.path.join(p["moge_dir"], name)):
            cloud_path = os.path.join(p["moge_dir"], name)
            break
    if cloud_path is None:
        raise FileNotFoundError(f"no MoGe geometry (mesh.glb / pointcloud.ply / mesh.ply) in {p['moge_dir']}")
    geo = load(cloud_path)
    pts = np.asarray(geo.vertices, dtype=np.float64)
    # MoGe's MESH (mesh.glb in the reference, pipelines.py:1247-1250) also supplies the rendered targets of the image terms
    moge_mesh = (pts.astype(np.float32), np.asarray(geo.faces, dtype=np.int32)) if isinstance(geo, TriMesh) and len(geo.faces) else None
    if pts.shape[0] >= n_cloud:
        pts = pts[rng.choice(pts.shape[0], n_cloud, replace=False)]
    else:                                                                           # fewer points than the batch size: repeat them cyclically
        pts = pts[np.resize(np.arange(pts.shape[0]), n_cloud)]
    return dict(fovx=fovx, hw=(H, W), kps=kps, hand_moge=hand_moge.astype(np.float32), moge_mesh=moge_mesh,
                hand_mask=(hand_mask if hand_mask.ndim == 2 else hand_mask[..., 0]) > 0,
                obj_mask=(obj_mask if obj_mask.ndim == 2 else obj_mask[..., 0]) > 0,
                faces=np.asarray(mano.faces, dtype=np.int32), T_h2m=T.astype(np.float32), cloud=pts.astype(np.float32))


def similarity_about(points: np.ndarray, theta8: np.ndarray, center: np.ndarray) -> np.ndarray:
    """``transform_mesh_around_center_w_scale`` (pipelines.py:108-118) for a fixed centre."""
    R = quat_to_mat_np(theta8[4:8])
    return (float(theta8[0]) * (points - center)) @ R.T + center + theta8[1:4]


# --------------------------------------------------------------------------- the stage
def _load_task_list(task_list_file: Optional[str], cropped_obj_img_dir: str) -> List[str]:
    """run.py:178-185, with the torchrun rank standing in for SLURM_ARRAY_TASK_ID."""
    rank, world = rank_world()
    if task_list_file and os.path.exists(task_list_file):
        with open(task_list_file, "r", encoding="utf-8") as f:
            chunks = json.load(f)
        task_id = int(os.environ.get("SLURM_ARRAY_TASK_ID", rank))
        return task_chunk(chunks, task_id)
    return shard_images(os.listdir(cropped_obj_img_dir), rank, world)


def run(
    project_root: str,
    cropped_obj_img_dir: str,
    mask_dir: str,
    moge_out_dir: str,
    hunyuan_hoi_mesh_dir: str,
    hamer_out_dir: str,
    h2m_rt_dir: str,
    aligned_mano_dir: str,
    guidance_out_dir: str,
    task_list_file: Optional[str] = None,
    *,
    model: Optional[GuidanceModel] = None,
    batch_size: int = 8,
    n_cloud: int = 65536,
    device: Optional[str] = None,
    config: Optional[OptimizationConfig] = None,
    j_regressor_path: str = J_REGRESSOR_PATH,
    seed: int = 2,
    obj_cap_factor: int = 6,
    latent_dtype: Optional[torch.dtype] = None,
    export_resolution: Optional[int] = None,
) -> None:
    """Positional/keyword arguments up to ``task_list_file`` are the reference's (run.py:188-199).  ``latent_dtype``: the dtype
    the latents and the model output being optimised are carried in -- default: half with the reference's own networks
    on the tensor cores (a model with ``tc_decoder``; the reference's pipeline is half, pipelines.py:1204), float otherwise.
    ``export_resolution``: octree resolution of the FINAL decode with a ``tc_decoder`` model -- default 384 like the
    reference (pipelines.py:1624-1641: a 385^3 lattice, 57 M queries, surface extracted from it); 0 = the loop's lattice."""
    del project_root                      # the reference only uses it to extend sys.path (run.py:57-62)
    if model is None:
        model = load_model_from_env()
    if model is None:
        raise _lib.FohoLibraryError(
            "the guidance stage needs the Hunyuan3D networks (DiT velocity + latent->SDF decoder) wrapped in a "
            f"GuidanceModel: pass model=... or set {MODEL_ENV}=package.module:factory.  There is no fallback.")
    _lib.load()
    config = config or OptimizationConfig()
    os.makedirs(guidance_out_dir, exist_ok=True)
    assigned_imgs = _load_task_list(task_list_file, cropped_obj_img_dir)
    rng = np.random.default_rng(seed)

    # ---- gather the images that are to be processed (skip rules of run.py:224-236)
    todo = []
    for cropped_obj_img in assigned_imgs:
        try:
            p = index_paths(cropped_obj_img, cropped_obj_img_dir, mask_dir, moge_out_dir, hunyuan_hoi_mesh_dir,
                            hamer_out_dir, h2m_rt_dir, aligned_mano_dir, guidance_out_dir)
            if os.path.exists(p["save_path_obj"]) and os.path.exists(p["save_path_hand"]):
                print(f"{p['index']} already exists, skipping")
                continue
            inp = load_image_inputs(p, n_cloud, rng)
            if "skip" in inp:
                print(f"Skipping {p['index']} due to {inp['skip']}")
                continue
            todo.append((p, inp))
        except Exception as e:                                   # run.py:257-259: report and go on
            print(f"Error in processing {cropped_obj_img} : {e}")
    if not todo:
        print("Finished processing all images")
        return
    J = torch.load(j_regressor_path, map_location="cpu")
    J = torch.as_tensor(np.asarray(J), dtype=torch.float32).reshape(16, -1)

    from .loop import GuidanceLoop
    dev = torch.device(device if device is not None else f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}")
    for chunk in batches_of_compatible_images(todo, batch_size):
        idx = [p["index"] for p, _ in chunk]
        try:
            _run_batch(chunk, model, J, config, n_cloud, dev, seed, GuidanceLoop, obj_cap_factor, latent_dtype, export_resolution)
            for i in idx:
                print(f"Reconstructed object {i}")
        except Exception as e:
            if len(chunk) == 1:
                print(f"Error in reconstruction for {idx} : {e}")
                continue
            # one bad image must not take its batch mates with it (the reference works image by image,
            # run.py:208-259): retry them one at a time
            print(f"Error in reconstruction for batch {idx} : {e}; retrying its images one by one")
            for item in chunk:
                try:
                    _run_batch([item], model, J, config, n_cloud, dev, seed, GuidanceLoop, obj_cap_factor, latent_dtype, export_resolution)
                    print(f"Reconstructed object {item[0]['index']}")
                except Exception as e1:
                    print(f"Error in reconstruction for {item[0]['index']} : {e1}")
    print("Finished processing all images")


def batch_key(inp: dict) -> tuple:
    """Images that may share one batched launch: the kernels take ONE camera (fov, crop size) and ONE hand
    topology per batch.  MoGe estimates fov_x per image (geometry/moge.py:126-131), crops differ, and left hands
    come with flipped winding, so real image sets are heterogeneous: group first, never assume."""
    import hashlib
    f = np.ascontiguousarray(inp["faces"], dtype=np.int32)
    return (tuple(int(x) for x in inp["hw"]), round(float(inp["fovx"]), 6), f.shape[0], hashlib.sha1(f.tobytes()).hexdigest())


def batches_of_compatible_images(todo: list, batch_size: int) -> List[list]:
    """Split ``todo`` [(paths, inputs)] into batches of at most ``batch_size`` images with equal ``batch_key``,
    keeping the sorted image order inside every group; singletons become batches of one."""
    groups: Dict[tuple, list] = {}
    for item in todo:
        groups.setdefault(batch_key(item[1]), []).append(item)
    out = []
    for items in groups.values():
        for b0 in range(0, len(items), max(1, int(batch_size))):
            out.append(items[b0:b0 + batch_size])
    return out


def _run_batch(chunk, model: GuidanceModel, J: torch.Tensor, config: OptimizationConfig, n_cloud: int, dev, seed: int,
               GuidanceLoop, obj_cap_factor: int = 6, latent_dtype: Optional[torch.dtype] = None,
               export_resolution: Optional[int] = None) -> None:
    B = len(chunk)
    inputs = [inp for _, inp in chunk]
    faces0 = inputs[0]["faces"]
    for inp in inputs[1:]:
        if inp["faces"].shape != faces0.shape or not np.array_equal(inp["faces"], faces0):
            raise ValueError("images of one batch must share the hand topology (MANO)")
    if len({inp["hw"] for inp in inputs}) != 1 or len({round(inp["fovx"], 6) for inp in inputs}) != 1:
        raise ValueError("images of one batch must share the crop size and field of view; use batch_size=1 otherwise")
    capped = cap_boundary_loops(faces0)                          # closed topology for the sign rule
    st = GuidanceStatics(
        hand_rest=torch.from_numpy(np.stack([i["hand_moge"] for i in inputs])).to(dev).contiguous(),
        hand_faces=torch.from_numpy(capped.astype(np.int32)).to(dev).contiguous(),
        cloud=torch.from_numpy(np.stack([i["cloud"] for i in inputs])).to(dev).contiguous(),
        T_h2m=torch.from_numpy(np.stack([i["T_h2m"] for i in inputs])).to(dev).contiguous(),
        obj_center=torch.from_numpy(np.stack([i["T_h2m"][:3, 3] for i in inputs])).to(dev).contiguous(),
        j_regressor=J.to(dev).contiguous(),
        kps_2d=torch.from_numpy(np.stack([i["kps"] for i in inputs])).to(dev).contiguous(),
        fov_deg=inputs[0]["fovx"], image_hw=inputs[0]["hw"])
    model.begin_batch([p["index"] for p, _ in chunk], [p["cropped_obj_img_path"] for p, _ in chunk], dev)
    debug_root = os.environ.get("FOHO_DEBUG_DIR")               # pipelines.py:1076-1091: debug dumps when set
    gen = torch.Generator().manual_seed(seed)                   # run.py:120 torch.manual_seed(2)
    export_meshes = None                                        # tensor-core decoder path: surfaces of the final decode
    last = config.num_inference_steps - 1
    decode = getattr(model, "decode", None)
    tc_decoder = getattr(model, "tc_decoder", None)
    if callable(tc_decoder):
        # the reference's own decoder on the tensor cores: latent2sdf + adjoint as kernels, no autograd (row f1)
        dec = tc_decoder(B)
        cap_obj = B * int(obj_cap_factor) * model.D * model.D      # extracted-surface capacity: a closed surface has O(D^2) cubes
        loop = GuidanceLoop(B, model.D, st, n_cloud, device=dev, config=config, latent_elems=model.latent_elems,
                            loss_log_every=10 if debug_root else 0, mock_decoder=False, max_obj_verts=cap_obj,
                            latent_dtype=latent_dtype or torch.float16)
        if all(inp.get("moge_mesh") is not None for inp in inputs):
            # the reference's image terms (pipelines.py:1327-1349,1413-1440,1544-1569): targets rendered once per image from
            # MoGe's mesh, then hand / object-only / joined renders every inner iteration, object mesh extracted from the volume
            from .render import stack_targets, targets_from_moge
            per = [targets_from_moge(i["moge_mesh"][0], i["moge_mesh"][1], i["fovx"], i["hand_mask"], i["obj_mask"], device=dev) for i in inputs]
            fovs = [i["fovx"] for i in inputs]
            loop.enable_image_terms(stack_targets(per, "hand", fovs), hand_faces_render=torch.from_numpy(faces0.astype(np.int32)).to(dev))
            loop.enable_object_terms(hoi_targets=stack_targets(per, "hoi", fovs), obj_targets=stack_targets(per, "obj", fovs), cap_verts=cap_obj)
        else:
            loop.enable_object_terms(cap_verts=cap_obj)          # REF mesh terms a7 / a10 on the extracted surface, no renders
        loop.sdf.fill_(1.0)                                     # finite "outside" volume for the hand-only phase
        loop.x_t.copy_(model.initial_latents(B, gen))
        loop.reset_leaves()
        loop.run_schedule_tc_decoder(model.predict, dec)
        loop.check_overflow()
        loop._obj.ex.check_flags()                              # a truncated surface = wrong terms: fail loudly
        x1 = (loop.x_t + (1.0 - float(loop.sigmas[last])) * loop.velocity).contiguous()      # final decode (:1641), sigma_last = 1
        # the final step re-grids to octree resolution 384 (:1624-1641) and extracts the surface from THAT volume
        res = 384 if export_resolution is None else int(export_resolution)
        eD = res + 1 if res > 0 else model.D
        sdf_dev = (dec.forward(x1.view(B, 3072, 64)) if eD == model.D else dec.decode_lattice(x1.view(B, 3072, 64), eD)).view(B, eD, eD, eD)
        from .surface import SurfaceExtractor
        ex = SurfaceExtractor(1, eD, device=dev, cap_verts=int(obj_cap_factor) * eD * eD, with_edges=False)
        export_meshes = []
        for b in range(B):                                      # one image at a time: 385^3 floats and its surface per pass
            ex.extract(sdf_dev[b:b + 1])
            ex.check_flags()
            v, f, _ = ex.meshes()[0]
            export_meshes.append((v.numpy().astype(np.float64), f.numpy()))
        sdf = None
    elif callable(decode):
        # a differentiable network decoder in the loop (eager; autograd carries dE/dSDF to the model output)
        loop = GuidanceLoop(B, model.D, st, n_cloud, device=dev, config=config, latent_elems=model.latent_elems,
                            loss_log_every=10 if debug_root else 0)
        loop.sdf.fill_(1.0)                                     # finite "outside" volume for the hand-only phase
        loop.x_t.copy_(model.initial_latents(B, gen))
        loop.reset_leaves()
        loop.run_schedule_decoder(model.predict, decode)
        with torch.no_grad():                                   # final decode (:1641), sigma_last = 1
            x1 = loop.x_t + (1.0 - float(loop.sigmas[last])) * loop.velocity
            sdf = decode(x1).float().reshape(B, model.D, model.D, model.D).cpu().numpy()
    else:
        sdf0, tap, alpha = model.decoder_state()
        loop = GuidanceLoop(B, model.D, st, n_cloud, device=dev, config=config, latent_elems=model.latent_elems,
                            decoder_alpha=float(alpha), loss_log_every=10 if debug_root else 0,
                            latent_dtype=latent_dtype or torch.float32)
        loop.tap = tap.to(dev).to(torch.int64).contiguous()
        loop.sdf0.copy_(sdf0); loop.sdf.copy_(sdf0)
        loop.x_t.copy_(model.initial_latents(B, gen))
        loop.reset_leaves()
        loop.run_schedule_device(model.predict)
        # ---- outputs (pipelines.py:1641-1679: final decode, meshes in MoGe space)
        # step_final on the already advanced latents, as the reference does it (:1612-1623); sigma_last = 1
        x1 = loop.x_t + (1.0 - float(loop.sigmas[last])) * loop.velocity
        flat0 = loop.sdf0.reshape(B, -1)
        sdf = flat0.clone()
        sdf[:, loop.tap] = flat0[:, loop.tap] + loop.alpha * x1
        sdf = sdf.reshape(B, model.D, model.D, model.D).cpu().numpy()
    theta = loop.theta.cpu().numpy().astype(np.float64)
    torch.cuda.synchronize(dev)
    loop.check_flags()                  # truncated penetration term = wrong result: fail the batch loudly (it is retried per image)
    nan_rep, failed = loop.nan_report(), set(loop.failed_images())
    if debug_root:
        _write_debug_dumps(debug_root, [p["index"] for p, _ in chunk], config, loop)
    skip = report_nan_images([p["cropped_obj_img_path"] for p, _ in chunk], nan_rep, failed)
    for b, (p, inp) in enumerate(chunk):
        if b in skip:
            continue
        hand = inp["hand_moge"].astype(np.float64)
        ch = (hand.min(0) + hand.max(0)) / 2.0
        verts, faces = export_meshes[b] if export_meshes is not None else model.extract_mesh(sdf[b])
        T = inp["T_h2m"].astype(np.float64)
        try:                                                               # run.py:155-167
            vm = np.asarray(verts, dtype=np.float64).reshape(-1, 3) @ T[:3, :3].T + T[:3, 3]
            obj = TriMesh(similarity_about(vm, theta[b, 8:], T[:3, 3]) if len(vm) else vm,
                          np.asarray(faces, dtype=np.int64).reshape(-1, 3))
            obj = FaceReducer()(DegenerateFaceRemover()(FloaterRemover()(obj)))
            if len(obj.vertices) == 0:
                print(f"Empty mesh for {p['cropped_obj_img_path']}")      # run.py:170-172
                continue
            # object first, hand only once the object is on disk (run.py:168-175): a failed object must not
            # leave a hand-only output behind
            write_ply(p["save_path_obj"], obj.vertices, obj.faces)
            write_ply(p["save_path_hand"], similarity_about(hand, theta[b, :8], ch), inp["faces"])
        except Exception:
            print(f"Error in saving mesh for {p['cropped_obj_img_path']}")
            for k in ("save_path_obj", "save_path_hand"):
                if os.path.exists(p[k]):
                    os.remove(p[k])
            continue


def report_nan_images(image_paths: Sequence[str], nan_report: dict, failed) -> set:
    """Messages of the reference for images whose loss became NaN: "Total loss is NaN" (pipelines.py:1443,1591),
    and for those whose ``__call__`` returned ``None`` (:1442-1444) the per-image error of the stage -- unpacking
    ``None`` raises inside its try block (run.py:141,257-259), the image is reported and nothing is written.
    Returns the images to skip."""
    skip = set()
    for b, path in enumerate(image_paths):
        if any(b in imgs for imgs in nan_report.values()):
            print("Total loss is NaN")
        if b in failed:
            print(f"Error in processing {os.path.basename(path)} : cannot unpack non-iterable NoneType object")
            skip.add(b)
    return skip


def _write_debug_dumps(debug_root: str, indices: Sequence[str], config: OptimizationConfig, loop) -> None:
    """``FOHO_DEBUG_DIR`` dumps of the reference (pipelines.py:1076-1091,1145-1183,1446-1450,1594-1598): one
    ``{timestamp}_exp_obj{index}_inpainted`` directory per image with ``params.json`` (the optimisation
    configuration) and ``losses.txt`` (the loss terms of every 10th inner iteration)."""
    import datetime
    stamp = datetime.datetime.now().strftime("%Y%m%d_%H%M%S")
    params = {k: v for k, v in vars(config).items() if isinstance(v, (int, float, bool, str, dict, list, tuple))}
    for b, index in enumerate(indices):
        save_dir = os.path.join(debug_root, f"{stamp}_exp_obj{index}_inpainted")
        os.makedirs(save_dir, exist_ok=True)
        with open(os.path.join(save_dir, "params.json"), "w") as f:
            json.dump(params, f, indent=4)
        with open(os.path.join(save_dir, "losses.txt"), "w") as f:
            for line in loop.loss_log_lines(b):
                f.write(line + "\n")


# --------------------------------------------------------------------------- mock networks
class MockGuidanceModel(GuidanceModel):
    """Synthetic stand-in for the two networks: a per-image ellipsoid base volume, the linear tap
    decoder, a fixed random velocity field decaying over the steps.  Tests and synthetic runs only."""

    def __init__(self, D: int = 64, latent_elems: int = 3072 * 64, alpha: float = 0.05, seed: int = 0):
        self.D, self.latent_elems, self.alpha, self.seed = D, latent_elems, alpha, seed
        vol = D ** 3
        if latent_elems > vol or latent_elems % 64 or vol % 64:
            raise ValueError("need latent_elems <= D^3, both multiples of 64")
        g = torch.Generator().manual_seed(seed)
        starts = torch.randperm(vol // 64, generator=g)[: latent_elems // 64].sort().values * 64
        self.tap = (starts.view(-1, 1) + torch.arange(64).view(1, -1)).reshape(-1)
        self._sdf0 = None
        self._v = None

    def begin_batch(self, indices, image_paths, device) -> None:
        from ..synthetic import ellipsoid_volume
        vols = []
        for i in indices:
            s = sum(ord(c) for c in str(i)) + self.seed
            vols.append(ellipsoid_volume(self.D, s, device="cpu")[0])
        self._sdf0 = torch.stack(vols).to(device)
        g = torch.Generator().manual_seed(self.seed + 17)
        self._v = (0.1 * torch.randn(len(indices), self.latent_elems, generator=g)).to(device)

    def initial_latents(self, batch, generator):
        return torch.randn(batch, self.latent_elems, generator=generator)

    def predict(self, step, x_t):
        return self._v / (1.0 + step)

    def decoder_state(self):
        return self._sdf0, self.tap, self.alpha

    def extract_mesh(self, sdf: np.ndarray):
        """Boundary faces of the occupied voxels (a closed, blocky surface) on the Hunyuan lattice."""
        D = sdf.shape[0]
        occ = np.zeros((D + 2,) * 3, dtype=bool)
        occ[1:-1, 1:-1, 1:-1] = sdf < 0
        step = 2.2 / (D - 1)
        quads = []
        corner = {0: [(0, 0, 0), (0, 1, 0), (0, 1, 1), (0, 0, 1)], 1: [(0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0)],
                  2: [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)]}
        for ax in range(3):
            for sgn in (0, 1):                         # face on the low / high side of the voxel along ax
                nb = np.roll(occ, 1 if sgn == 0 else -1, axis=ax)
                cells = np.argwhere(occ & ~nb) - 1     # voxel indices (un-padded)
                if cells.size == 0:
                    continue
                offs = np.array(corner[ax], dtype=np.int64)
                if sgn:
                    offs = offs[::-1].copy()
                    offs[:, ax] = 1
                quads.append(cells[:, None, :] + offs[None])
        if not quads:
            return np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int32)
        q = np.concatenate(quads, 0).reshape(-1, 3)
        uniq, inv = np.unique(q, axis=0, return_inverse=True)
        inv = inv.reshape(-1, 4)
        faces = np.concatenate([inv[:, [0, 1, 2]], inv[:, [0, 2, 3]]], 0).astype(np.int32)
        verts = (uniq.astype(np.float64) - 0.5) * step - 1.10          # voxel corners around lattice points
        return verts, faces


def main() -> None:
    parser = argparse.ArgumentParser(description="Hunyuan3D-2 guidance")
    parser.add_argument("--project_root", required=True)
    parser.add_argument("--cropped_obj_img_dir", required=True)
    parser.add_argument("--mask_dir", required=True)
    parser.add_argument("--moge_out_dir", required=True)
    parser.add_argument("--hunyuan_hoi_mesh_dir", required=True)
    parser.add_argument("--hamer_out_dir", required=True)
    parser.add_argument("--h2m_rt_dir", required=True)
    parser.add_argument("--aligned_mano_dir", required=True)
    parser.add_argument("--guidance_out_dir", required=True)
    parser.add_argument("--task_list_file", default=None)
    args = parser.parse_args()

    run(
        project_root=args.project_root,
        cropped_obj_img_dir=args.cropped_obj_img_dir,
        mask_dir=args.mask_dir,
        moge_out_dir=args.moge_out_dir,
        hunyuan_hoi_mesh_dir=args.hunyuan_hoi_mesh_dir,
        hamer_out_dir=args.hamer_out_dir,
        h2m_rt_dir=args.h2m_rt_dir,
        aligned_mano_dir=args.aligned_mano_dir,
        guidance_out_dir=args.guidance_out_dir,
        task_list_file=args.task_list_file,
    )


if __name__ == "__main__":
    main()
