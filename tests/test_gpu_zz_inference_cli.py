"""(named to run last under `pytest -x`)
-m gpu: the inference.py command line end to end on synthetic weights (full SD-2.1 config):
folder of LQ images -> BSRInferenceLoop -> SwinIRPipeline -> PNGs of upscale x the input size."""
import sys
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def test_cli_synthetic_end_to_end(tmp_path):
    import inference as cli
    from diffbir_b200.utils.synth import synthetic_lq
    (tmp_path / "in").mkdir()
    lq = synthetic_lq(128, 160, seed=3)[0]
    Image.fromarray(lq).save(tmp_path / "in" / "img.png")
    cli.main(["--task", "sr", "--version", "v2.1", "--upscale", "4", "--sampler", "spaced", "--steps", "3",
              "--cfg_scale", "4.0", "--pos_prompt", "", "--neg_prompt", "low quality, blurry", "--captioner", "none",
              "--input", str(tmp_path / "in"), "--output", str(tmp_path / "out"), "--synthetic"])
    out = np.array(Image.open(tmp_path / "out" / "img.png"))
    assert out.shape == (512, 640, 3) and out.dtype == np.uint8
    assert out.std() > 1.0                                    # an image, not a constant
    assert (tmp_path / "out" / "prompt.csv").exists()
