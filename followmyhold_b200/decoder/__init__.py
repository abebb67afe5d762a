"""Row f1: the latent -> SDF decode of the reference's ``latent2sdf``
(third_party_patches/hy3dgen/shapegen/pipelines.py:292-312) on the B200's tcgen05 tensor cores."""
