"""ORACLE (test infrastructure): fp32 functional restatement of the stage-1 SwinIR forward
(reference model/swinir.py), driven by the reference state_dict.  Configuration as in
configs/inference/swinir.yaml: PixelUnshuffle(8) stem, 8 RSTB x 6 Swin blocks, dim 180,
6 heads, window 8, mlp ratio 2, 'nearest+conv' x8 reconstruction, '1conv' residual.

Pinned by tests/test_oracle_golden.py against fixtures generated from the reference module.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .cldm import _indices

SD = Dict[str, torch.Tensor]
RGB_MEAN = (0.4488, 0.4371, 0.4040)   # swinir.py:687-689 (plain attribute, not a buffer)


def shift_mask(h: int, w: int, ws: int, shift: int, device) -> torch.Tensor:
    """SW-MSA mask [nW, ws*ws, ws*ws] with -100 / 0 entries — swinir.py:222-243."""
    img = torch.zeros(h, w, device=device)
    bounds = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
    cnt = 0
    for hs in bounds:
        for wsl in bounds:
            img[hs, wsl] = cnt
            cnt += 1
    win = img.view(h // ws, ws, w // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def rel_pos_index(ws: int) -> torch.Tensor:
    """relative_position_index buffer — swinir.py:91-103."""
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    coords = torch.stack([ys.flatten(), xs.flatten()])          # [2, ws*ws]
    rel = coords[:, :, None] - coords[:, None, :]
    return (rel[0] + ws - 1) * (2 * ws - 1) + (rel[1] + ws - 1)


def swin_block(sd: SD, p: str, x, h, w, ws, heads, shift):
    """SwinTransformerBlock.forward + WindowAttention.forward — swinir.py:245-285, 120-151."""
    b, n, c = x.shape
    y = F.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5).view(b, h, w, c)
    if shift:
        y = torch.roll(y, (-shift, -shift), (1, 2))
    win = y.view(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, c)
    nwin = win.shape[0]
    qkv = F.linear(win, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    qkv = qkv.view(nwin, ws * ws, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * ((c // heads) ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-1, -2)
    table = sd[p + "attn.relative_position_bias_table"]
    idx = rel_pos_index(ws).to(table.device)
    bias = table[idx.view(-1)].view(ws * ws, ws * ws, heads).permute(2, 0, 1)
    attn = attn + bias[None]
    if shift:
        m = shift_mask(h, w, ws, shift, x.device)                  # [nW, N, N]
        attn = attn.view(b, m.shape[0], heads, ws * ws, ws * ws) + m[None, :, None]
        attn = attn.view(nwin, heads, ws * ws, ws * ws)
    o = (torch.softmax(attn, dim=-1) @ v).transpose(1, 2).reshape(nwin, ws * ws, c)
    o = F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    o = o.view(b, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    x = x + o.reshape(b, n, c)
    y = F.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    y = F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    return x + F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def swinir_forward(sd: SD, x: torch.Tensor, window: int = 8, heads: int = 6, sf: int = 8,
                   img_range: float = 1.0) -> torch.Tensor:
    """SwinIR.forward ('nearest+conv', unshuffle) — swinir.py:856-894."""
    hh, ww = x.shape[2:]
    ph, pw = (window - hh % window) % window, (window - ww % window) % window
    x = F.pad(x.float(), (0, pw, 0, ph), mode="reflect")
    mean = torch.tensor(RGB_MEAN, device=x.device).view(1, 3, 1, 1)
    x = (x - mean) * img_range
    # conv_first = PixelUnshuffle(sf) + conv3x3 (swinir.py:700-705)
    x = F.conv2d(F.pixel_unshuffle(x, sf), sd["conv_first.1.weight"], sd["conv_first.1.bias"], padding=1)
    b, c, h, w = x.shape
    t = x.flatten(2).transpose(1, 2)
    t = F.layer_norm(t, (c,), sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"], 1e-5)
    for li in _indices(sd, "layers."):
        res = t
        for bi in _indices(sd, f"layers.{li}.residual_group.blocks."):
            shift = 0 if bi % 2 == 0 else window // 2
            t = swin_block(sd, f"layers.{li}.residual_group.blocks.{bi}.", t, h, w, window, heads, shift)
        y = t.transpose(1, 2).reshape(b, c, h, w)
        y = F.conv2d(y, sd[f"layers.{li}.conv.weight"], sd[f"layers.{li}.conv.bias"], padding=1)
        t = y.flatten(2).transpose(1, 2) + res
    t = F.layer_norm(t, (c,), sd["norm.weight"], sd["norm.bias"], 1e-5)
    y = t.transpose(1, 2).reshape(b, c, h, w)
    x = F.conv2d(y, sd["conv_after_body.weight"], sd["conv_after_body.bias"], padding=1) + x
    # reconstruction (swinir.py:876-885); LeakyReLU default slope 0.01 after conv_before_upsample
    x = F.leaky_relu(F.conv2d(x, sd["conv_before_upsample.0.weight"], sd["conv_before_upsample.0.bias"], padding=1), 0.01)
    for name in ("conv_up1", "conv_up2", "conv_up3"):
        if name + ".weight" in sd:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.leaky_relu(F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=1), 0.2)
    x = F.leaky_relu(F.conv2d(x, sd["conv_hr.weight"], sd["conv_hr.bias"], padding=1), 0.2)
    x = F.conv2d(x, sd["conv_last.weight"], sd["conv_last.bias"], padding=1)
    x = x / img_range + mean
    return x[:, :, : hh * sf, : ww * sf]
