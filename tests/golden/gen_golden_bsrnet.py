"""Fixture for the RRDBNet (BSRNet) cleaner, produced by the REFERENCE module (diffbir.model.RRDBNet) with
seeded synthetic weights loaded strictly (which also pins diffbir_b200.arch.rrdbnet_shapes):

    python tests/golden/gen_golden_bsrnet.py          # needs /root/reference; writes bsrnet_small.npz
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
from tests.small_cfg import RRDB_SMALL  # noqa: E402

OUT = Path(__file__).resolve().parent


@torch.no_grad()
def main():
    from diffbir.model.bsrnet import RRDBNet
    from diffbir.pipeline import BSRNetPipeline
    net = RRDBNet(**RRDB_SMALL).eval()
    sd = make_state_dict(arch.rrdbnet_shapes(RRDB_SMALL), 7)
    net.load_state_dict(sd, strict=True)
    x = torch.rand(2, 3, 24, 40, generator=torch.Generator().manual_seed(70))
    y = net(x)
    pipe = BSRNetPipeline(net, None, None, None, "cpu", upscale=4.0)
    lq = torch.rand(1, 3, 20, 12, generator=torch.Generator().manual_seed(71))
    pipe.set_output_size(lq.size())
    cond_small = pipe.apply_cleaner(lq, False, 512, 256)             # 80 x 48 output < 512 -> short edge resized to 512
    pipe.upscale = 45.0
    pipe.set_output_size(lq.size())
    cond_big = pipe.apply_cleaner(lq, False, 512, 256)               # 900 x 540: resized to the output size
    # the resized images are stored sub-sampled (every 6th row / column): they only pin the resize branch
    np.savez_compressed(OUT / "bsrnet_small.npz", x=x.numpy(), y=y.numpy(), lq=lq.numpy(),
                        cond_small=cond_small[..., ::6, ::6].numpy(), cond_small_shape=np.array(cond_small.shape),
                        cond_big=cond_big[..., ::6, ::6].numpy(), cond_big_shape=np.array(cond_big.shape))
    print("wrote bsrnet_small.npz", y.shape, cond_small.shape, cond_big.shape, float(y.abs().mean()))


if __name__ == "__main__":
    main()
