"""RRDBNet — drop-in counterpart of the reference's diffbir.model.RRDBNet (model/bsrnet.py:72-104; the BSRNet
stage-1 cleaner of the v2 blind-SR recipe, configs/inference/bsrnet.yaml), backed by engine.bsrnet.RRDBNetEngine."""
from __future__ import annotations

from typing import Dict

import torch

from .. import arch
from ..engine.bsrnet import RRDBNetEngine


class RRDBNet:
    def __init__(self, in_nc=3, out_nc=3, nf=64, nb=23, gc=32, sf=4, device="cuda"):
        if sf not in (2, 4):
            raise NotImplementedError("RRDBNet: sf must be 2 or 4 (bsrnet.py:83-86)")
        self.cfg = dict(in_nc=in_nc, out_nc=out_nc, nf=nf, nb=nb, gc=gc, sf=sf)
        self.sf = sf
        self.device = torch.device(device)
        self.engine = None

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        if "state_dict" in sd:
            sd = sd["state_dict"]
        if next(iter(sd)).startswith("module."):
            sd = {k[len("module."):]: v for k, v in sd.items()}
        exp = arch.rrdbnet_shapes(self.cfg)
        if strict:
            extra, lack = set(sd) - set(exp), set(exp) - set(sd)
            if extra or lack:
                raise RuntimeError(f"RRDBNet state_dict mismatch: missing {sorted(lack)[:3]}, unexpected {sorted(extra)[:3]}")
        self.engine = RRDBNetEngine({k: sd[k] for k in exp}, self.cfg, self.device)

    def eval(self):
        return self

    def to(self, device):
        return self

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[B, 3, H, W] in [0, 1] -> [B, 3, sf*H, sf*W] (bsrnet.py:89-104)."""
        if self.engine is None:
            raise RuntimeError("load_state_dict() first")
        return self.engine.forward(x.to(self.device, torch.float32).contiguous())

    __call__ = forward
