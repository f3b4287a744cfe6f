"""SwinIR — drop-in counterpart of the reference's diffbir.model.SwinIR (model/swinir.py:571-905)
for the configuration of configs/inference/swinir.yaml, backed by engine.swinir.SwinIREngine."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .. import arch
from ..engine.swinir import SwinIREngine


class SwinIR:
    def __init__(self, img_size=64, patch_size=1, in_chans=3, embed_dim=96, depths=(6, 6, 6, 6),
                 num_heads=(6, 6, 6, 6), window_size=7, mlp_ratio=4.0, sf=4, img_range=1.0,
                 upsampler="", resi_connection="1conv", unshuffle=False, unshuffle_scale=None,
                 device="cuda", **unused):
        if not (upsampler == "nearest+conv" and resi_connection == "1conv" and unshuffle and patch_size == 1):
            raise NotImplementedError("only the DiffBIR stage-1 SwinIR variant (configs/inference/swinir.yaml)")
        self.cfg = dict(img_size=img_size, in_chans=in_chans, embed_dim=embed_dim, depths=tuple(depths),
                        num_heads=tuple(num_heads), window_size=window_size, mlp_ratio=mlp_ratio, sf=sf,
                        img_range=img_range, unshuffle_scale=unshuffle_scale, num_feat=64)
        self.window_size = window_size
        self.upscale = sf
        self.device = torch.device(device)
        self.engine = None

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """Accepts the reference checkpoint layout, including the relative_position_index /
        attn_mask buffers (recomputed analytically by the kernels, so ignored here)."""
        if "state_dict" in sd:
            sd = sd["state_dict"]
        if next(iter(sd)).startswith("module."):
            sd = {k[len("module."):]: v for k, v in sd.items()}
        params = {k: v for k, v in sd.items()
                  if not (k.endswith("relative_position_index") or k.endswith("attn_mask"))}
        exp = arch.swinir_shapes(self.cfg)
        if strict:
            extra, lack = set(params) - set(exp), set(exp) - set(params)
            if extra or lack:
                raise RuntimeError(f"SwinIR state_dict mismatch: missing {sorted(lack)[:3]}, unexpected {sorted(extra)[:3]}")
        self.engine = SwinIREngine(params, self.cfg, self.device)

    def eval(self):
        return self

    def to(self, device):
        return self

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """SwinIR.forward (swinir.py:856-894): reflect-pad to the window size, run, crop."""
        if self.engine is None:
            raise RuntimeError("load_state_dict() first")
        hh, ww = x.shape[2:]
        m = self.cfg["unshuffle_scale"] * self.window_size     # token grid must tile into windows
        ph, pw = (m - hh % m) % m, (m - ww % m) % m
        if ph or pw:
            if ph > (self.window_size - hh % self.window_size) % self.window_size or \
               pw > (self.window_size - ww % self.window_size) % self.window_size:
                raise ValueError("SwinIR input must be a multiple of 64 (the pipeline pads it, pipeline.py:385)")
            x = F.pad(x, (0, pw, 0, ph), mode="reflect")
        y = self.engine.forward(x.to(self.device, torch.float32).contiguous())
        return y[:, :, : hh * self.upscale, : ww * self.upscale]

    __call__ = forward
