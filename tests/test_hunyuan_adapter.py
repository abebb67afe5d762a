"""The concrete GuidanceModel around the reference's pipeline object (followmyhold_b200/guidance/hunyuan_adapter.py),
exercised against stand-ins with the surface the adapter uses (``vae.state_dict()``, ``model(...)``,
``scheduler.timesteps``, ``encode_cond``, ``prepare_latents``): hy3dgen and its weights are not available offline.
The GPU test drives synthetic frames through the whole guidance stage with the tensor-core decoder in the loop."""
import os
import types

import numpy as np
import pytest
import torch


class _StandInDiT(torch.nn.Module):
    """(latents [2,3072,64] fp16, timesteps, cond, guidance) -> velocity: cond half and uncond half differ."""

    def forward(self, x, t, cond, guidance=None):
        bias = torch.cat([cond["main"][:1], cond["main"][1:]]).view(2, 1, 1).to(x.dtype)
        return 0.1 * torch.tanh(x) * (1.0 - t.view(-1, 1, 1)) + bias


def _standin_pipe(layers=2, seed=0, device="cpu"):
    from oracle import decoder_oracle as DO
    torch.manual_seed(seed)
    vae = DO.ShapeVAE(num_decoder_layers=layers).float()
    with torch.no_grad():
        vae.geo_decoder.output_proj.weight.mul_(6.0)
    pipe = types.SimpleNamespace()
    pipe.vae = vae
    pipe.model = _StandInDiT()
    pipe.scheduler = types.SimpleNamespace(timesteps=torch.linspace(0, 1000, 6), config=types.SimpleNamespace(num_train_timesteps=1000))
    pipe.encode_cond = lambda image, mask, do_classifier_free_guidance, dual_guidance: {"main": torch.tensor([0.02, -0.01], device=device)}
    pipe.prepare_latents = lambda b, dtype, dev, gen: torch.randn(b, 3072, 64, generator=gen).to(dtype).to(dev)
    return pipe


def test_cfg_velocity_follows_the_reference_schedule():
    """:1281-1291: v = uncond + s (cond - uncond), s = 5 until guidance_start_step, then 5 (1 - i / N)."""
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.hunyuan_adapter import HunyuanGuidanceModel
    pipe = _standin_pipe()
    cfg = OptimizationConfig().with_steps(6)
    m = HunyuanGuidanceModel.__new__(HunyuanGuidanceModel)          # no CUDA needed for the DiT side
    m.pipe, m.cfg, m.guidance, m.device = pipe, cfg, None, torch.device("cpu")
    m.cond = [pipe.encode_cond(None, None, True, False)]
    x = torch.randn(1, 3072 * 64)
    for i in (0, cfg.guidance_start_step, cfg.guidance_start_step + 1, 5):
        v = m.predict(i, x)
        t = pipe.scheduler.timesteps[i] / 1000
        base = (0.1 * torch.tanh(x.half().view(1, 3072, 64)) * (1 - t.half())).float().reshape(1, -1)
        s = 5.0 if i < cfg.guidance_start_step + 1 else 5.0 * (1 - i / 6)
        ref = base + (-0.01 + s * (0.02 - (-0.01)))
        assert torch.allclose(v, ref, atol=2e-3), i


@pytest.mark.gpu
def test_stage_runs_with_the_tensor_core_decoder(tmp_path, capsys):
    from followmyhold_b200.guidance import run as R
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.hunyuan_adapter import HunyuanGuidanceModel
    from tests.test_guidance_stage import _kwargs, write_dataset
    D = 17
    d, jpath = write_dataset(str(tmp_path), 2, D=D, P=600)
    cfg = OptimizationConfig()
    cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = 3, 2, 2
    cfg.with_steps(6)
    model = HunyuanGuidanceModel(_standin_pipe(device="cuda:0"), cfg, D=D, device="cuda:0")
    # the stand-in decoder has random weights: its volume is noise with a large surface -> generous extraction capacity
    R.run(**_kwargs(d), model=model, batch_size=2, n_cloud=512, config=cfg, j_regressor_path=jpath, obj_cap_factor=24,
          export_resolution=20)
    out = capsys.readouterr().out
    assert "Finished processing all images" in out and "Error" not in out, out
    produced = sorted(os.listdir(d["out"]))
    assert all(f.endswith(("_hand.ply", "_obj.ply")) for f in produced)
    assert any(f.endswith("_hand.ply") for f in produced) == any(f.endswith("_obj.ply") for f in produced)   # never a hand without its object


@pytest.mark.gpu
def test_stage_with_moge_mesh_runs_every_image_term(tmp_path, capsys):
    """MoGe geometry WITH faces (mesh.glb in the reference, pipelines.py:1247-1250; a PLY mesh here): the stage renders the
    targets from it and runs the hand, object-only and joined hand + object image terms on the extracted surface."""
    from followmyhold_b200.guidance import run as R
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.hunyuan_adapter import HunyuanGuidanceModel
    from followmyhold_b200.meshio import load, write_ply
    from followmyhold_b200.synthetic import standin_hand_mesh
    from tests.test_guidance_stage import _kwargs, write_dataset
    D = 17
    d, jpath = write_dataset(str(tmp_path), 2, D=D, P=600)
    hv, hf = standin_hand_mesh(0.5)
    for k in range(2):
        md = os.path.join(d["moge"], f"{k:03d}_cropped_hoi")
        cloud = np.asarray(load(os.path.join(md, "pointcloud.ply")).vertices)
        os.remove(os.path.join(md, "pointcloud.ply"))
        write_ply(os.path.join(md, "mesh.ply"), hv.astype(np.float64) * 0.6 + cloud.mean(0), hf)      # a surface where the scene is
    cfg = OptimizationConfig()
    cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = 2, 2, 2
    cfg.with_steps(6)
    model = HunyuanGuidanceModel(_standin_pipe(device="cuda:0"), cfg, D=D, device="cuda:0")
    R.run(**_kwargs(d), model=model, batch_size=2, n_cloud=512, config=cfg, j_regressor_path=jpath, obj_cap_factor=24,
          export_resolution=20)
    out = capsys.readouterr().out
    assert "Finished processing all images" in out and "Error" not in out, out
