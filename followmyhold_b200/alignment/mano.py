"""Align MANO hand mesh to the Hunyuan HOI mesh.

Same ``run(...)`` signature, CLI flags, file naming and ICP hyper-parameters as the
reference stage ``src/foho/alignment/mano.py:12-61``; the ICP loops run on the GPU, those of ``concurrent``
images at a time (``mesh_align.align_meshes_many``, results identical to one ``align_meshes_impl`` call per
image).  Under ``torchrun`` rank r aligns ``sorted(meshes)[r::world]``.
"""
from __future__ import annotations

import argparse
import glob
import os

from ..parallel import default_device, rank_world, shard_images
from .mesh_align import STAGE_ICP_KWARGS, align_meshes_many


def run(hamer_out_dir: str, hunyuan_mesh_dir: str, aligned_mano_dir: str, seed: int = 0, device: str = None,
        concurrent: int = 8) -> None:
    meshes = sorted(glob.glob(os.path.join(hamer_out_dir, "*.obj")))
    if not meshes:
        print(f"No HaMeR meshes found in {hamer_out_dir}")
        return
    os.makedirs(aligned_mano_dir, exist_ok=True)
    rank, world = rank_world()
    meshes = shard_images(meshes, rank, world)

    def job(mesh_path):
        # {i}_*.obj is aligned to {i}_hoi_mesh.ply and written as {name}_aligned_mano.ply (mano.py:18-23)
        stem = os.path.splitext(os.path.basename(mesh_path))[0]
        return (mesh_path, os.path.join(hunyuan_mesh_dir, f"{stem.split('_')[0]}_hoi_mesh.ply"), None,
                os.path.join(aligned_mano_dir, f"{stem}_aligned_mano.ply"))

    jobs = [job(m) for m in meshes]
    align_meshes_many(jobs, **STAGE_ICP_KWARGS, seed=seed, device=device or default_device(), concurrent=concurrent)


FLAGS = ('hamer_out_dir', 'hunyuan_mesh_dir', 'aligned_mano_dir')      # the reference stage's CLI flags = run()'s arguments


def main() -> None:
    parser = argparse.ArgumentParser()
    for flag in FLAGS:
        parser.add_argument(f"--{flag}", required=True)
    run(**vars(parser.parse_args()))


if __name__ == "__main__":
    main()
