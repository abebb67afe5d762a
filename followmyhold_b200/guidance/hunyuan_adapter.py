"""A concrete ``GuidanceModel`` around the reference's networks: the object returned by
``Hunyuan3DDiTFlowMatchingPipeline_main.from_pretrained(...)`` (src/foho/guidance/run.py:140), loaded ONCE per process
(the reference reloads it for every image).

Every network call below is the one the reference's ``__call__`` makes at the cited line of
third_party_patches/hy3dgen/shapegen/pipelines.py; nothing of ``hy3dgen`` is imported here -- the adapter only
relies on the attributes the pipeline object exposes (``vae``, ``model``, ``scheduler``, ``encode_cond``,
``prepare_latents``), so it is exercised offline against stand-ins with the same surface
(tests/test_hunyuan_adapter.py).

The DiT stays a torch module (network inference: out of this path's scope); the VAE decoder does NOT: its
``state_dict`` is handed to ``decoder.shapevae.DecoderWeights`` and ``latent2sdf`` + its adjoint run on the
tensor-core kernels (``GuidanceLoop.run_schedule_tc_decoder``).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch

from ..decoder.shapevae import DecoderWeights, LatentDecoder, lattice_points
from .config import OptimizationConfig

LATENT_TOKENS, LATENT_CHANNELS = 3072, 64


class HunyuanGuidanceModel:
    """Duck-typed ``guidance.run.GuidanceModel`` with a tensor-core decoder (``tc_decoder``)."""

    def __init__(self, pipe, config: Optional[OptimizationConfig] = None, D: int = 65, device="cuda:0",
                 prepare_image: Optional[Callable] = None, guidance: Optional[torch.Tensor] = None,
                 extract: Optional[Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray]]] = None):
        """``pipe``: the reference's pipeline object.  ``prepare_image(path) -> (image, mask)``: the image preparation
        of :1095-1118 (rembg / white -> alpha, resize) -- outside this path; ``extract(sdf [D,D,D]) -> (verts, faces)``:
        the surface extraction of the final export (:1624-1660; FlexiCubes in the reference)."""
        self.pipe, self.cfg = pipe, config or OptimizationConfig()
        self.D, self.latent_elems = int(D), LATENT_TOKENS * LATENT_CHANNELS
        self.device = torch.device(device)
        self.prepare_image, self.guidance, self.extract = prepare_image, guidance, extract
        vae = pipe.vae
        self.weights = DecoderWeights({k: v for k, v in vae.state_dict().items()}, self.device,
                                      scale_factor=float(getattr(vae, "scale_factor", 1.0)))
        self.xyz = lattice_points(self.D)                      # generate_dense_grid_points(indexing="ij"), :1126-1139
        self._decoders = {}
        self.cond = None

    # ---- decoder on the tensor cores
    def tc_decoder(self, batch: int) -> LatentDecoder:
        d = self._decoders.get(batch)
        if d is None:
            d = LatentDecoder(self.weights, batch, device=self.device)
            d.set_queries(self.xyz)
            self._decoders[batch] = d
        return d

    # ---- GuidanceModel protocol
    def begin_batch(self, indices: Sequence[str], image_paths: Sequence[str], device) -> None:
        """:1095-1123 per image: image preparation, then ``encode_cond`` with classifier-free guidance."""
        conds = []
        for path in image_paths:
            image, mask = self.prepare_image(path) if self.prepare_image is not None else (path, None)
            conds.append(self.pipe.encode_cond(image=image, mask=mask, do_classifier_free_guidance=True, dual_guidance=False))
        self.cond = conds

    def initial_latents(self, batch: int, generator: torch.Generator) -> torch.Tensor:
        """``prepare_latents`` (:1204-1205; fp16 in the reference), one draw per image."""
        lat = [self.pipe.prepare_latents(1, torch.float16, self.device, generator).float().reshape(1, -1) for _ in range(batch)]
        return torch.cat(lat).to(self.device)

    def predict(self, step: int, x_t: torch.Tensor) -> torch.Tensor:
        """:1269-1291: DiT on ``[latents] * 2``, classifier-free guidance with the scale decaying after
        ``guidance_start_step`` (``scale * (1 - i / N)`` for ``i >= guidance_start_step + 1``)."""
        cfg, pipe = self.cfg, self.pipe
        t = pipe.scheduler.timesteps[step].to(x_t.device)
        scale = cfg.obj_guidance_scale
        if step >= cfg.guidance_start_step + 1:
            scale = cfg.obj_guidance_scale * (1 - step / cfg.num_inference_steps)
        out = []
        with torch.no_grad():
            for b in range(x_t.shape[0]):
                lat = x_t[b].view(1, LATENT_TOKENS, LATENT_CHANNELS).to(torch.float16)
                inp = torch.cat([lat] * 2)
                ts = t.expand(inp.shape[0]).to(lat.dtype) / pipe.scheduler.config.num_train_timesteps
                v = pipe.model(inp, ts, self.cond[b], guidance=self.guidance)
                v_cond, v_uncond = v.chunk(2)
                out.append((v_uncond + scale * (v_cond - v_uncond)).float().reshape(1, -1))
        return torch.cat(out)

    def extract_mesh(self, sdf: np.ndarray):
        if self.extract is not None:
            return self.extract(sdf)
        from .run import MockGuidanceModel                      # closed blocky surface of the occupied voxels
        return MockGuidanceModel.extract_mesh(self, sdf)
