"""CPU oracle for the Adv-GRPO rollout -> score -> advantage -> update hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is a plain torch-fp32 / numpy-fp64
restatement of the reference's algorithm (showlab/Adv-GRPO @ 8287f90) and of the
third-party modules the reference calls (diffusers 0.33.1, transformers 4.54.0,
timm DINOv2 -- none of which are vendored in /root/reference nor installed in
this image; their published algorithms are restated from the architecture
configs named in SURVEY.md section 8c).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or as the reported CPU baseline -- never on the product path.  The
product (``adv_grpo_b200``) does not import anything from here and fails loudly
when its CUDA library is missing.

Parity pinning: the reference holds NO tests or golden
vectors.  The pieces of the reference that are importable in the build
container (``adv_grpo/stat_tracking.py``, ``adv_grpo/diffusers_patch/
sd3_sde_with_logprob.py`` behind a diffusers stub, ``CLIPCriterion.calc_loss``,
``adv_grpo/ema.py``) were executed by ``tests/golden/make_golden.py`` and their
outputs are committed under ``tests/golden/``; ``tests/test_oracle_golden.py``
pins the corresponding oracle functions to them.  The model bodies (MMDiT-X,
VAE decoder, CLIP-ViT-H, DINOv2-B) are cross-checked against the same-
architecture implementations shipped in ``transformers`` where one exists
(CLIP, DINOv2); MMDiT-X and the VAE decoder have no runnable reference on this
box: **parity unpinned** for those two bodies (restated from the diffusers
0.33.1 architecture).
"""
