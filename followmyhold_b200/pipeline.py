"""The three replaced stages in ONE resident process per GPU, driven from the reference's own config file.

``foho.main.run_pipeline`` (src/foho/main.py:94-278) runs every stage as ``python3 -m foho.<stage>`` inside a
freshly activated conda environment (``runner.py:10-21``): for the three stages this package replaces that is
three interpreter start-ups, three CUDA contexts and -- inside the guidance stage -- one
``from_pretrained`` per image (``guidance/run.py:140``).  ``run_hot_stages`` reads the same ``.env`` file
(``configs/pipeline.py:51-146``: same keys, same defaults, same quoting rules), derives the same directories,
and calls the stage mirrors in the same order with the same arguments as ``main.py:229-278`` -- in process:
one CUDA context, one library load, the networks (a ``GuidanceModel``) loaded once.  Under ``torchrun`` every
rank takes ``sorted(images)[rank::world]`` in all three stages.

    python -m followmyhold_b200.pipeline --config configs/pipeline.env        (FOHO_B200_GUIDANCE_MODEL set)
"""
from __future__ import annotations

import argparse
import os
from dataclasses import dataclass
from typing import Dict, Optional


@dataclass(frozen=True)
class StagePaths:
    """The ``PipelineConfig`` fields the alignment and guidance stages use (configs/pipeline.py:12-40)."""
    project_root: str
    base_dir: str
    cropped_inpainted_obj: str
    mask_dir_path: str
    moge_out_path: str
    hunyuan_hoi_mesh_path: str
    hamer_out_path: str
    h2m_rt_path: str
    aligned_mano_path: str
    guidance_out_path: str


def parse_env_file(path: str) -> Dict[str, str]:
    """``KEY=value`` lines; blank lines, ``#`` comments and lines without ``=`` are skipped; one layer of double
    then single quotes is stripped from the value (configs/pipeline.py:51-64)."""
    data: Dict[str, str] = {}
    with open(path, "r", encoding="utf-8") as f:
        for raw in f:
            line = raw.strip()
            if not line or line.startswith("#") or "=" not in line:
                continue
            key, val = line.split("=", 1)
            data[key.strip()] = val.strip().strip('"').strip("'")
    return data


def load_stage_paths(path: str) -> StagePaths:
    """Same derivation as ``load_config`` (configs/pipeline.py:67-146) for the directories of the three stages."""
    if not os.path.isfile(path):
        raise FileNotFoundError(f"Missing config: {path}")
    env = parse_env_file(path)
    project_root, base_dir = env.get("PROJECT_ROOT"), env.get("BASE_DIR")
    if not project_root or not base_dir:
        raise ValueError("PROJECT_ROOT and BASE_DIR are required in config")
    p = lambda key, default: env.get(key, default)
    return StagePaths(
        project_root=project_root, base_dir=base_dir,
        cropped_inpainted_obj=p("CROPPED_INPAINTED_OBJ", f"{base_dir}/ours_inpaint"),
        mask_dir_path=p("MASK_DIR_PATH", f"{base_dir}/cropped_hand_masks"),
        moge_out_path=p("MOGE_OUT_PATH", f"{base_dir}/moge_out"),
        hunyuan_hoi_mesh_path=p("HUNYUAN_HOI_MESH_PATH", f"{base_dir}/hunyuan_hoi_out"),
        hamer_out_path=p("HAMER_OUT_PATH", f"{base_dir}/hamer_out"),
        h2m_rt_path=p("H2M_RT_PATH", f"{base_dir}/h2m_transformations"),
        aligned_mano_path=p("ALIGNED_MANO_PATH", f"{base_dir}/aligned_mano"),
        guidance_out_path=p("GUIDANCE_OUT_PATH", f"{base_dir}/guidance_out"))


def stage_calls(sp: StagePaths):
    """(module, kwargs) of the three stages exactly as ``run_pipeline`` passes them (main.py:229-278)."""
    return [
        ("alignment.h2m", {"hunyuan_mesh_dir": sp.hunyuan_hoi_mesh_path, "moge_out_dir": sp.moge_out_path,
                           "h2m_rt_dir": sp.h2m_rt_path}),
        ("alignment.mano", {"hamer_out_dir": sp.hamer_out_path, "hunyuan_mesh_dir": sp.hunyuan_hoi_mesh_path,
                            "aligned_mano_dir": sp.aligned_mano_path}),
        ("guidance.run", {"project_root": sp.project_root, "cropped_obj_img_dir": sp.cropped_inpainted_obj,
                          "mask_dir": sp.mask_dir_path, "moge_out_dir": sp.moge_out_path,
                          "hunyuan_hoi_mesh_dir": sp.hunyuan_hoi_mesh_path, "hamer_out_dir": sp.hamer_out_path,
                          "h2m_rt_dir": sp.h2m_rt_path, "aligned_mano_dir": sp.aligned_mano_path,
                          "guidance_out_dir": sp.guidance_out_path}),
    ]


def run_hot_stages(config_path: str, model=None, device: str = "cuda:0", **guidance_kwargs) -> StagePaths:
    """alignment.h2m -> alignment.mano -> guidance.run in this process.  ``model``: a ``GuidanceModel`` (default:
    the factory named by ``FOHO_B200_GUIDANCE_MODEL``); ``guidance_kwargs`` go to ``guidance.run.run``
    (``batch_size``, ``n_cloud``, ``config`` ...).  The guidance stage resolves its key-point regressor relative
    to the working directory (pipelines.py:1218), so -- like ``foho.main`` (``cwd=cfg.project_root``) -- the
    stages run with the project root as the working directory when it exists."""
    import importlib
    sp = load_stage_paths(config_path)
    for d in (sp.h2m_rt_path, sp.aligned_mano_path, sp.guidance_out_path):           # main.py:109-124
        os.makedirs(d, exist_ok=True)
    cwd = os.getcwd()
    try:
        if os.path.isdir(sp.project_root):
            os.chdir(sp.project_root)
        for name, kwargs in stage_calls(sp):
            mod = importlib.import_module(f"followmyhold_b200.{name}")
            if name == "guidance.run":
                mod.run(**kwargs, model=model, device=device, **guidance_kwargs)
            else:
                mod.run(**kwargs, device=device)
    finally:
        os.chdir(cwd)
    return sp


def main() -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", required=True)                                   # main.py:281-284
    parser.add_argument("--device", default=None, help="default: cuda:$LOCAL_RANK")
    args = parser.parse_args()
    device: Optional[str] = args.device or f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
    run_hot_stages(args.config, device=device)


if __name__ == "__main__":
    main()
