"""CPU: when the unmodified reference is installed under baseline/_ref (baseline/install_ref.sh), its own loop
(ImageSlicer.split -> tta.*_image_augment -> model -> tta.*_image_deaugment -> TileMerger) must agree bit for bit with the
oracle's restatement -- this pins oracle/pipeline.py (the end-to-end checker of the -m gpu suite) to the real reference
on every box where the install travelled. Skipped only when the install is absent."""
import numpy as np
import pytest
import torch

from baseline import reference_arm as RA
from oracle import pipeline as OP

needs_ref = pytest.mark.skipif(not RA.available(), reason="baseline/_ref not installed (run baseline/install_ref.sh)")


@needs_ref
@pytest.mark.parametrize("shape,tile,step,group,K,weight", [
    ((300, 260, 3), 64, 32, "d4", 4, "pyramid"),
    ((257, 301, 3), 64, 32, "flips", 8, "pyramid"),
    ((400, 390, 1), 128, 64, None, 1, "gaussian"),
    ((200, 180, 3), 64, 32, "d2", 2, "mean"),
])
def test_oracle_pipeline_is_bit_identical_to_the_installed_reference(shape, tile, step, group, K, weight):
    rng = np.random.default_rng(11)
    image = rng.integers(0, 256, shape).astype(np.uint8)
    ref = RA.tiled_inference(image, tile, step, K, weight=weight, group=group, batch_size=5)
    ours = OP.tiled_inference_cpu(image, tile, step, K, weight=weight, group=group or "none", batch_size=5)
    assert torch.equal(ref, ours)


@needs_ref
def test_reference_arm_imports_the_install_not_the_repo_package():
    mod = RA.load()
    assert "baseline/_ref" in mod.__file__.replace("\\", "/")
    import pytorch_toolbelt_b200

    assert "baseline" not in pytorch_toolbelt_b200.__file__
