"""Golden fixture for Diffusion.q_sample (start_point_type="cond", noise_aug), produced by the
REFERENCE (`diffbir.model.gaussian_diffusion.Diffusion`, read-only from /root/reference).

    python tests/golden/gen_golden_qsample.py        ->  tests/golden/qsample.npz
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle" / "_shims"))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from _ref_import import use_reference  # noqa: E402

use_reference()      # `import diffbir` = the reference checkout, not this repo's alias package
OUT = Path(__file__).resolve().parent


def main():
    from diffbir.model.gaussian_diffusion import Diffusion
    g = torch.Generator().manual_seed(80)
    x0, noise = torch.randn(4, 4, 8, 8, generator=g), torch.randn(4, 4, 8, 8, generator=g)
    t = torch.tensor([0, 37, 500, 999])
    out = {"x0": x0.numpy(), "noise": noise.numpy(), "t": t.numpy()}
    for name, kw in (("eps", {}), ("v", dict(parameterization="v", zero_snr=True))):
        d = Diffusion(linear_start=0.00085, linear_end=0.0120, timesteps=1000, **kw)
        out[f"q_{name}"] = d.q_sample(x0, t, noise).numpy()
    np.savez_compressed(OUT / "qsample.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
