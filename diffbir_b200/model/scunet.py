"""SCUNet — drop-in counterpart of the reference's diffbir.model.SCUNet (model/scunet.py:163-243; the stage-1
cleaner of the v2 blind-denoising recipe, configs/inference/scunet.yaml), backed by engine.scunet.SCUNetEngine."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .. import arch
from ..engine.scunet import SCUNetEngine


class SCUNet:
    def __init__(self, in_nc=3, config=(2, 2, 2, 2, 2, 2, 2), dim=64, drop_path_rate=0.0, input_resolution=256,
                 device="cuda"):
        if input_resolution // 8 <= 8:
            raise NotImplementedError("SCUNet: input_resolution <= 64 turns shifted blocks into plain ones (scunet.py:106-107)")
        self.cfg = dict(in_nc=in_nc, config=tuple(config), dim=dim)
        self.device = torch.device(device)
        self.engine = None

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        if "state_dict" in sd:
            sd = sd["state_dict"]
        if next(iter(sd)).startswith("module."):
            sd = {k[len("module."):]: v for k, v in sd.items()}
        exp = arch.scunet_shapes(self.cfg)
        if strict:
            extra, lack = set(sd) - set(exp), set(exp) - set(sd)
            if extra or lack:
                raise RuntimeError(f"SCUNet state_dict mismatch: missing {sorted(lack)[:3]}, unexpected {sorted(extra)[:3]}")
        self.engine = SCUNetEngine({k: sd[k] for k in exp}, self.cfg, self.device)

    def eval(self):
        return self

    def to(self, device):
        return self

    @torch.no_grad()
    def forward(self, x0: torch.Tensor) -> torch.Tensor:
        """scunet.py:221-243: replicate-pad to a multiple of 64, run, crop."""
        if self.engine is None:
            raise RuntimeError("load_state_dict() first")
        h, w = x0.shape[-2:]
        pb, pr = int(np.ceil(h / 64) * 64 - h), int(np.ceil(w / 64) * 64 - w)
        x = F.pad(x0.to(self.device, torch.float32), (0, pr, 0, pb), mode="replicate").contiguous()
        return self.engine.forward(x)[..., :h, :w]

    __call__ = forward
