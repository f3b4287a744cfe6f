"""CPU/GPU fp32 restatement of the reference's RRDBNet (diffbir/model/bsrnet.py:36-104) and of
BSRNetPipeline.apply_cleaner (diffbir/pipeline.py:343-366), driven by the reference's state_dict keys.
TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs only; pinned by
tests/golden/bsrnet_small.npz, which the reference module itself produced (gen_golden_bsrnet.py)."""
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _conv(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=1)


def _rdb(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """ResidualDenseBlock_5C.forward — bsrnet.py:51-58."""
    feats = [x]
    for k in range(1, 5):
        feats.append(F.leaky_relu(_conv(sd, p + f"conv{k}", torch.cat(feats, 1)), 0.2))
    return _conv(sd, p + "conv5", torch.cat(feats, 1)) * 0.2 + x


def rrdbnet_forward(sd: SD, x: torch.Tensor, sf: int = 4) -> torch.Tensor:
    """RRDBNet.forward — bsrnet.py:89-104."""
    nb = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("RRDB_trunk."))
    fea = _conv(sd, "conv_first", x.float())
    t = fea
    for b in range(nb):
        o = t
        for r in (1, 2, 3):
            o = _rdb(sd, f"RRDB_trunk.{b}.RDB{r}.", o)
        t = o * 0.2 + t                                   # RRDB.forward, bsrnet.py:66-70
    fea = fea + _conv(sd, "trunk_conv", t)
    fea = F.leaky_relu(_conv(sd, "upconv1", F.interpolate(fea, scale_factor=2, mode="nearest")), 0.2)
    if sf == 4:
        fea = F.leaky_relu(_conv(sd, "upconv2", F.interpolate(fea, scale_factor=2, mode="nearest")), 0.2)
    return _conv(sd, "conv_last", F.leaky_relu(_conv(sd, "HRconv", fea), 0.2))


def bsrnet_apply_cleaner(cleaner, lq: torch.Tensor, upscale: float) -> torch.Tensor:
    """BSRNetPipeline.apply_cleaner, un-tiled branch (pipeline.py:343-366) with set_output_size (:339-341)."""
    h, w = lq.shape[2:]
    out_size = (int(h * upscale), int(w * upscale))
    up4 = cleaner(lq)
    if min(out_size) < 512:
        hh, ww = up4.shape[2:]
        if hh == ww:
            size = (512, 512)
        elif hh < ww:
            size = (512, int(ww * (512 / hh)))
        else:
            size = (int(hh * (512 / ww)), 512)
        return F.interpolate(up4, size=size, mode="bicubic", antialias=True)
    return F.interpolate(up4, size=out_size, mode="bicubic", antialias=True)
