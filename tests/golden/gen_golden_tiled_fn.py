"""Golden fixture for the image-tiling wrapper and the tiled stage-1 branch, produced by the REFERENCE
(`diffbir.utils.common.make_tiled_fn`, `SwinIRPipeline.apply_cleaner(tiled=True)`, imported read-only
from /root/reference) with an analytic stand-in for the cleaner network.

    python tests/golden/gen_golden_tiled_fn.py        ->  tests/golden/tiled_fn.npz
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_shims"))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from _ref_import import use_reference  # noqa: E402

use_reference()      # `import diffbir` = the reference checkout, not this repo's alias package

OUT = Path(__file__).resolve().parent


def stand_in(t: torch.Tensor) -> torch.Tensor:
    """Depends on the tile's content as a whole (mean), so the tiling is visible in the result."""
    return torch.tanh(t) * 0.5 + t.mean(dim=(2, 3), keepdim=True) * 0.25


def stand_in_up4(t: torch.Tensor) -> torch.Tensor:
    """A x4 up-scaler whose output depends on the whole tile."""
    return torch.nn.functional.interpolate(stand_in(t), scale_factor=4, mode="bilinear", align_corners=False)


@torch.no_grad()
def main():
    import diffbir.pipeline as P
    from diffbir.utils.common import make_tiled_fn

    class Null:
        def __init__(self, *a, **k): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False
    P.VRAMPeakMonitor = Null
    g = torch.Generator().manual_seed(60)
    x = torch.rand(2, 3, 72, 56, generator=g)
    out = {"x": x.numpy(), "tiled_32_16": make_tiled_fn(stand_in, size=32, stride=16, progress=False)(x).numpy(),
           "tiled_40_24": make_tiled_fn(stand_in, size=40, stride=24, progress=False)(x).numpy()}
    lq = torch.rand(1, 3, 160, 136, generator=g)
    pipe = P.SwinIRPipeline(stand_in, None, None, None, "cpu")
    out["lq"] = lq.numpy()
    # tiled, then resized to short edge 512; and an input smaller than a tile (un-tiled branch). The
    # fixtures keep every 8th pixel (bit-exact comparison on the subsample keeps the file small).
    full = pipe.apply_cleaner(lq, True, 128, 64)
    tiny = pipe.apply_cleaner(lq[..., :100, :90], True, 128, 64)
    out["cleaner_tiled_128_64_shape"] = np.array(full.shape)
    out["cleaner_tiled_128_64_sub8"] = full[..., ::8, ::8].numpy()
    out["cleaner_tiny_untiled_shape"] = np.array(tiny.shape)
    out["cleaner_tiny_untiled_sub8"] = tiny[..., ::8, ::8].numpy()
    # x4 up-scaling tiles (make_tiled_fn(scale_type="up", scale=4), utils/common.py:196-205) and the tiled branches of the
    # two other stage-1 pipelines: BSRNetPipeline.apply_cleaner (pipeline.py:342-366) in both output-size regimes,
    # SCUNetPipeline.apply_cleaner (pipeline.py:402-421)
    out["tiled_up4_24_16_sub4"] = make_tiled_fn(stand_in_up4, size=24, stride=16, scale_type="up", scale=4, progress=False)(x)[..., ::4, ::4].numpy()
    for scale, key in ((2.0, "bsr_tiled_small"), (4.0, "bsr_tiled_big")):
        bp = P.BSRNetPipeline(stand_in_up4, None, None, None, "cpu", scale)
        bp.set_output_size(lq.size())
        y = bp.apply_cleaner(lq, True, 64, 48)
        out[key + "_shape"] = np.array(y.shape)
        out[key + "_sub8"] = y[..., ::8, ::8].numpy()
    sp = P.SCUNetPipeline(stand_in, None, None, None, "cpu")
    y = sp.apply_cleaner(lq, True, 64, 48)
    out["scunet_tiled_shape"] = np.array(y.shape)
    out["scunet_tiled_sub8"] = y[..., ::8, ::8].numpy()
    np.savez_compressed(OUT / "tiled_fn.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
