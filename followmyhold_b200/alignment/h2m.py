"""Align Hunyuan HOI mesh to MoGe geometry and write transforms.

Same ``run(...)`` signature, CLI flags, file naming and ICP hyper-parameters as the
reference stage ``src/foho/alignment/h2m.py:12-72``; the ICP loop runs on the GPU
(``mesh_align.align_meshes_many``: the loops of ``concurrent`` images at a time, results identical to one
``align_meshes_impl`` call per image).  Targets are looked up in the reference's order: ``mesh.ply``,
``pointcloud.ply``, ``mesh.glb`` (h2m.py:23-31).  Under ``torchrun`` rank r aligns ``sorted(meshes)[r::world]``.
"""
from __future__ import annotations

import argparse
import glob
import os

from ..parallel import default_device, rank_world, shard_images
from .mesh_align import STAGE_ICP_KWARGS, align_meshes_many


def run(hunyuan_mesh_dir: str, moge_out_dir: str, h2m_rt_dir: str, seed: int = 0, device: str = None,
        concurrent: int = 8) -> None:
    meshes = sorted(glob.glob(os.path.join(hunyuan_mesh_dir, "*.ply")))
    if not meshes:
        print(f"No Hunyuan HOI meshes found in {hunyuan_mesh_dir}")
        return
    os.makedirs(h2m_rt_dir, exist_ok=True)
    rank, world = rank_world()
    meshes = shard_images(meshes, rank, world)

    jobs = []
    for mesh_path in meshes:
        base_name = os.path.basename(mesh_path)
        i = base_name.split("_")[0]
        j = os.path.splitext(base_name)[0]
        moge_dir = os.path.join(moge_out_dir, f"{i}_cropped_hoi")
        # first of mesh.ply, pointcloud.ply, mesh.glb that exists -- the reference's order (h2m.py:23-31); MoGe
        # itself writes pointcloud.ply and / or mesh.glb
        target_mesh = next((c for c in (os.path.join(moge_dir, n) for n in ("mesh.ply", "pointcloud.ply", "mesh.glb"))
                            if os.path.isfile(c)), None)
        if target_mesh is None:
            print(f"No MoGe mesh found for {i} in {moge_dir}. Skipping.")
            continue
        jobs.append((mesh_path, target_mesh, os.path.join(h2m_rt_dir, j), None))
    align_meshes_many(jobs, **STAGE_ICP_KWARGS, seed=seed, device=device or default_device(), concurrent=concurrent)


FLAGS = ('hunyuan_mesh_dir', 'moge_out_dir', 'h2m_rt_dir')      # the reference stage's CLI flags = run()'s arguments


def main() -> None:
    parser = argparse.ArgumentParser()
    for flag in FLAGS:
        parser.add_argument(f"--{flag}", required=True)
    run(**vars(parser.parse_args()))


if __name__ == "__main__":
    main()
