"""Fixture for the SCUNet cleaner, produced by the REFERENCE module (diffbir.model.SCUNet) with seeded synthetic
weights loaded strictly (which also pins diffbir_b200.arch.scunet_shapes):

    python tests/golden/gen_golden_scunet.py          # needs /root/reference; writes scunet_small.npz
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from _ref_import import use_reference  # noqa: E402

use_reference()
from diffbir_b200 import arch  # noqa: E402
from diffbir_b200.utils.synth import make_state_dict  # noqa: E402
from tests.small_cfg import SCUNET_SMALL  # noqa: E402

OUT = Path(__file__).resolve().parent


@torch.no_grad()
def main():
    from diffbir.model.scunet import SCUNet
    net = SCUNet(in_nc=3, config=list(SCUNET_SMALL["config"]), dim=SCUNET_SMALL["dim"]).eval()
    sd = make_state_dict(arch.scunet_shapes(SCUNET_SMALL), 9)
    net.load_state_dict(sd, strict=True)
    x = torch.rand(2, 3, 120, 72, generator=torch.Generator().manual_seed(90))     # padded to 128 x 128 inside
    y = net(x)
    np.savez_compressed(OUT / "scunet_small.npz", x=x.numpy(), y=y.numpy())
    print("wrote scunet_small.npz", tuple(y.shape), float(y.abs().mean()), float(y.std()))


if __name__ == "__main__":
    main()
