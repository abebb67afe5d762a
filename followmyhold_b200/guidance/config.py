"""Guidance constants: the attribute names and values of the reference's ``OptimizationConfig``
(src/foho/configs/guid_config.py:6-32), which the pipeline reads through ``config()`` (pipelines.py:1145-1162).
The golden vectors hold the values of the reference class itself (tests/test_golden_reference.py)."""
from __future__ import annotations

import copy

_LR = lambda scale, trans, rot: {"scale": scale, "trans": trans, "rot": rot}

# name -> value, in the reference's order (guid_config.py:8-29).  Step bookkeeping (guidance_start_step,
# handopt_start_step, guidance_end_step) derives from num_inference_steps and is set by ``_schedule``.
DEFAULTS = (
    ("obj_guidance_scale", 5.0),                 # classifier-free guidance weight of the object branch
    ("batch_size", 1),                           # the reference runs one image at a time; the engine batches B images
    ("optimization_steps_hand", 200),            # inner iterations: hand-only step
    ("optimization_steps_joint", 50),            #                   every joint step
    ("optimization_steps_scale", 100),           #                   object-only step
    ("num_inference_steps", 20),
    ("phase1_hand_lrs", _LR(1e-2, 1e-2, 0.5)),   # per-leaf learning rates, by phase
    ("phase2_hand_lrs", _LR(1e-4, 1e-4, 1e-2)),
    ("obj_2half_lrs", _LR(1e-2, 1e-2, 1e-2)),
    ("obj_lrs", _LR(5e-2, 1e-2, 1e-2)),
    ("noise_obj_lr1", 1e-4),                     # velocity leaf: object-only phase
    ("noise_obj_lr2", 1e-2),                     #                joint phase
    ("use_intersection_loss", True),
)


class OptimizationConfig:
    def __init__(self):
        for name, value in DEFAULTS:
            setattr(self, name, copy.deepcopy(value))
        self._schedule(self.num_inference_steps)

    def _schedule(self, n: int) -> None:
        """Guidance starts half way, the hand-only step comes one step earlier, guidance runs to the end
        (guid_config.py:16-18)."""
        self.num_inference_steps = n
        self.guidance_start_step = n // 2
        self.handopt_start_step = self.guidance_start_step - 1
        self.guidance_end_step = n

    def with_steps(self, num_inference_steps: int) -> "OptimizationConfig":
        """BASELINE.json configs 2/4/5 use a 50-step loop; phases scale as in the reference."""
        self._schedule(num_inference_steps)
        return self

    def __call__(self):
        return self
