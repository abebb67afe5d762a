"""Guidance constants -- same names and values as the reference's
``OptimizationConfig`` (src/foho/configs/guid_config.py:6-32)."""
from __future__ import annotations


class OptimizationConfig:
    def __init__(self):
        self.obj_guidance_scale = 5.0
        self.batch_size = 1  # the reference processes one image at a time; the engine batches B images

        # Optimization steps
        self.optimization_steps_hand = 200
        self.optimization_steps_joint = 50
        self.optimization_steps_scale = 100
        self.num_inference_steps = 20
        self.guidance_start_step = self.num_inference_steps // 2
        self.handopt_start_step = self.guidance_start_step - 1
        self.guidance_end_step = self.num_inference_steps

        # Learning rates
        self.phase1_hand_lrs = {"scale": 1e-2, "trans": 1e-2, "rot": 0.5}
        self.phase2_hand_lrs = {"scale": 1e-4, "trans": 1e-4, "rot": 1e-2}
        self.obj_2half_lrs = {"scale": 1e-2, "trans": 1e-2, "rot": 1e-2}
        self.obj_lrs = {"scale": 5e-2, "trans": 1e-2, "rot": 1e-2}
        self.noise_obj_lr1 = 1e-4
        self.noise_obj_lr2 = 1e-2

        # Losses
        self.use_intersection_loss = True

    def with_steps(self, num_inference_steps: int) -> "OptimizationConfig":
        """BASELINE.json configs 2/4/5 use a 50-step loop; phases scale as in the reference."""
        self.num_inference_steps = num_inference_steps
        self.guidance_start_step = num_inference_steps // 2
        self.handopt_start_step = self.guidance_start_step - 1
        self.guidance_end_step = num_inference_steps
        return self

    def __call__(self):
        return self
