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

from ..parallel import rank_world, shard_images
from .mesh_align import align_meshes_many


def run(hunyuan_mesh_dir: str, moge_out_dir: str, h2m_rt_dir: str, seed: int = 0, device: str = "cuda:0",
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
        target_mesh = os.path.join(moge_dir, "mesh.ply")
        if not os.path.isfile(target_mesh):
            # MoGe typically writes pointcloud.ply and/or mesh.glb (h2m.py:25-31).
            pointcloud_mesh = os.path.join(moge_dir, "pointcloud.ply")
            glb_mesh = os.path.join(moge_dir, "mesh.glb")
            if os.path.isfile(pointcloud_mesh):
                target_mesh = pointcloud_mesh
            elif os.path.isfile(glb_mesh):
                target_mesh = glb_mesh
            else:
                print(f"No MoGe mesh found for {i} in {moge_dir}. Skipping.")
                continue
        jobs.append((mesh_path, target_mesh, os.path.join(h2m_rt_dir, j), None))
    align_meshes_many(
        jobs,
        fixed_scale=False,
        outliers=0.2,
        test_rotations=False,
        test_reflections=False,
        on_surface=False,
        iterations_coarse=50,
        count_source_coarse=1000,
        count_target_coarse=5000,
        iterations_fine=100,
        count_source_fine=5000,
        count_target_fine=10000,
        min_scale=0.7,
        max_scale=3.0,
        plot=False,
        seed=seed,
        device=device,
        concurrent=concurrent,
    )


def main() -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("--hunyuan_mesh_dir", required=True)
    parser.add_argument("--moge_out_dir", required=True)
    parser.add_argument("--h2m_rt_dir", required=True)
    args = parser.parse_args()
    run(hunyuan_mesh_dir=args.hunyuan_mesh_dir, moge_out_dir=args.moge_out_dir, h2m_rt_dir=args.h2m_rt_dir)


if __name__ == "__main__":
    main()
