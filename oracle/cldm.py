"""ORACLE (test infrastructure, never shipped or timed as the product).

CPU/any-device fp32 restatement, in plain torch functional ops, of the reference's stage-2
networks: ControlNet + ControlledUnetModel + AutoencoderKL + the CLIP text tower.  Every
function takes the reference's own state_dict (same key names) and recovers the block
structure from the keys, so it is independent of diffbir_b200's packed layouts.

Parity pin: tests/test_oracle_golden.py checks these functions against fixtures produced by
running the reference modules themselves (tests/golden/gen_golden.py, which imports
/root/reference).  The reference ships no tests or golden vectors of its own.

Reference citations are relative to the DiffBIR checkout.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _sub(sd: SD, prefix: str) -> SD:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _indices(sd: SD, prefix: str) -> List[int]:
    """Sorted distinct integer components that directly follow `prefix` in the keys."""
    out = set()
    for k in sd:
        if k.startswith(prefix):
            head = k[len(prefix):].split(".", 1)[0]
            if head.isdigit():
                out.add(int(head))
    return sorted(out)


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """cat(cos, sin) sinusoidal embedding — model/util.py:128-148."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def group_norm(x, w, b, eps):
    return F.group_norm(x.float(), 32, w, b, eps)


# --------------------------------------------------------------------------------------
# UNet / ControlNet building blocks
# --------------------------------------------------------------------------------------
def res_block(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """ResBlock._forward — model/unet.py:203-223 (GroupNorm32 eps 1e-5, util.py:191-193)."""
    h = F.silu(group_norm(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"], 1e-5))
    h = F.conv2d(h, sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.silu(group_norm(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"], 1e-5))
    h = F.conv2d(h, sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if p + "skip_connection.weight" in sd:
        x = F.conv2d(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
    return x + h


def attention(q, k, v, heads: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(dh)) v over `heads` heads — model/attention.py:189-216."""
    b, n, c = q.shape
    dh = c // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, dh).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    s = torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5)
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    return o.permute(0, 2, 1, 3).reshape(b, n, c)


def cross_attention(sd: SD, p: str, x, ctx, heads):
    q = F.linear(x, sd[p + "to_q.weight"])
    k = F.linear(ctx, sd[p + "to_k.weight"])
    v = F.linear(ctx, sd[p + "to_v.weight"])
    o = attention(q, k, v, heads)
    return F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def transformer_block(sd: SD, p: str, x, ctx, heads):
    """BasicTransformerBlock._forward — model/attention.py:265-274; GEGLU FF :19-45."""
    c = x.shape[-1]
    h = F.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    x = cross_attention(sd, p + "attn1.", h, h, heads) + x
    h = F.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    x = cross_attention(sd, p + "attn2.", h, ctx, heads) + x
    h = F.layer_norm(x, (c,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], 1e-5)
    u = F.linear(h, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"])
    a, g = u.chunk(2, dim=-1)
    h = a * F.gelu(g)
    x = F.linear(h, sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"]) + x
    return x


def spatial_transformer(sd: SD, p: str, x, ctx, head_dim: int = 64):
    """SpatialTransformer.forward (use_linear=True) — model/attention.py:334-353; GN eps 1e-6."""
    b, c, h, w = x.shape
    heads = c // head_dim
    y = group_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
    y = F.linear(y, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    for d in _indices(sd, p + "transformer_blocks."):
        y = transformer_block(sd, f"{p}transformer_blocks.{d}.", y, ctx, heads)
    y = F.linear(y, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    y = y.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return y + x


def _run_block(sd: SD, p: str, h, emb, ctx, head_dim):
    """TimestepEmbedSequential.forward — model/unet.py:40-48; layer kinds recovered from keys."""
    for j in _indices(sd, p):
        q = f"{p}{j}."
        if q + "in_layers.0.weight" in sd:
            h = res_block(sd, q, h, emb)
        elif q + "transformer_blocks.0.norm1.weight" in sd:
            h = spatial_transformer(sd, q, h, ctx, head_dim)
        elif q + "op.weight" in sd:                       # Downsample, unet.py:82-108
            h = F.conv2d(h, sd[q + "op.weight"], sd[q + "op.bias"], stride=2, padding=1)
        elif q + "conv.weight" in sd:                     # Upsample, unet.py:51-79
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[q + "conv.weight"], sd[q + "conv.bias"], padding=1)
        elif q + "weight" in sd:                          # bare conv (stem / zero-conv)
            w = sd[q + "weight"]
            h = F.conv2d(h, w, sd[q + "bias"], padding=w.shape[-1] // 2)
        else:
            raise KeyError(f"unrecognised layer at {q}")
    return h


def _time_embed(sd: SD, t, model_channels):
    e = timestep_embedding(t, model_channels)
    e = F.linear(e, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    return F.linear(F.silu(e), sd["time_embed.2.weight"], sd["time_embed.2.bias"])


def controlnet_forward(sd: SD, x, hint, t, ctx, head_dim: int = 64) -> List[torch.Tensor]:
    """ControlNet.forward — model/controlnet.py:314-328: 12 input blocks + middle, each
    through its 1x1 zero-conv."""
    mc = sd["time_embed.0.weight"].shape[1]
    emb = _time_embed(sd, t, mc)
    h = torch.cat([x, hint], dim=1).float()
    outs = []
    for i in _indices(sd, "input_blocks."):
        h = _run_block(sd, f"input_blocks.{i}.", h, emb, ctx, head_dim)
        outs.append(F.conv2d(h, sd[f"zero_convs.{i}.0.weight"], sd[f"zero_convs.{i}.0.bias"]))
    h = _run_block(sd, "middle_block.", h, emb, ctx, head_dim)
    outs.append(F.conv2d(h, sd["middle_block_out.0.weight"], sd["middle_block_out.0.bias"]))
    return outs


def unet_forward(sd: SD, x, t, ctx, control: Optional[List[torch.Tensor]] = None,
                 head_dim: int = 64) -> torch.Tensor:
    """ControlledUnetModel.forward — model/controlnet.py:18-47."""
    mc = sd["time_embed.0.weight"].shape[1]
    emb = _time_embed(sd, t, mc)
    control = None if control is None else list(control)
    hs = []
    h = x.float()
    for i in _indices(sd, "input_blocks."):
        h = _run_block(sd, f"input_blocks.{i}.", h, emb, ctx, head_dim)
        hs.append(h)
    h = _run_block(sd, "middle_block.", h, emb, ctx, head_dim)
    if control is not None:
        h = h + control.pop()
    for i in _indices(sd, "output_blocks."):
        skip = hs.pop()
        if control is not None:
            skip = skip + control.pop()
        h = torch.cat([h, skip], dim=1)
        h = _run_block(sd, f"output_blocks.{i}.", h, emb, ctx, head_dim)
    h = F.silu(group_norm(h, sd["out.0.weight"], sd["out.0.bias"], 1e-5))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def cldm_forward(unet_sd: SD, cn_sd: SD, x, t, c_txt, c_img, control_scales, head_dim: int = 64):
    """ControlLDM.forward — model/cldm.py:160-172."""
    control = controlnet_forward(cn_sd, x, c_img, t, c_txt, head_dim)
    control = [c * s for c, s in zip(control, control_scales)]
    return unet_forward(unet_sd, x, t, c_txt, control, head_dim)


# --------------------------------------------------------------------------------------
# AutoencoderKL (model/vae.py)
# --------------------------------------------------------------------------------------
def vae_resnet(sd: SD, p: str, x):
    """ResnetBlock.forward (temb=None) — model/vae.py:97-117; GN eps 1e-6."""
    h = F.silu(group_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6))
    h = F.conv2d(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = F.silu(group_norm(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6))
    h = F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if p + "nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return x + h


def vae_attn(sd: SD, p: str, x):
    """Single-head self-attention over pixels — model/vae.py:232-282 (== :139-164)."""
    b, c, hh, ww = x.shape
    y = group_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    q = F.conv2d(y, sd[p + "q.weight"], sd[p + "q.bias"]).reshape(b, c, hh * ww).transpose(1, 2)
    k = F.conv2d(y, sd[p + "k.weight"], sd[p + "k.bias"]).reshape(b, c, hh * ww).transpose(1, 2)
    v = F.conv2d(y, sd[p + "v.weight"], sd[p + "v.bias"]).reshape(b, c, hh * ww).transpose(1, 2)
    o = attention(q, k, v, 1).transpose(1, 2).reshape(b, c, hh, ww)
    return x + F.conv2d(o, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def vae_decode(sd: SD, z):
    """AutoencoderKL.decode -> Decoder.forward — model/vae.py:579-582, 526-559."""
    h = F.conv2d(z.float(), sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    p = "decoder."
    h = F.conv2d(h, sd[p + "conv_in.weight"], sd[p + "conv_in.bias"], padding=1)
    h = vae_resnet(sd, p + "mid.block_1.", h)
    h = vae_attn(sd, p + "mid.attn_1.", h)
    h = vae_resnet(sd, p + "mid.block_2.", h)
    for lvl in reversed(_indices(sd, p + "up.")):
        for j in _indices(sd, f"{p}up.{lvl}.block."):
            h = vae_resnet(sd, f"{p}up.{lvl}.block.{j}.", h)
        if f"{p}up.{lvl}.upsample.conv.weight" in sd:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"{p}up.{lvl}.upsample.conv.weight"],
                         sd[f"{p}up.{lvl}.upsample.conv.bias"], padding=1)
    h = F.silu(group_norm(h, sd[p + "norm_out.weight"], sd[p + "norm_out.bias"], 1e-6))
    return F.conv2d(h, sd[p + "conv_out.weight"], sd[p + "conv_out.bias"], padding=1)


def vae_encode_moments(sd: SD, x):
    """Encoder.forward + quant_conv — model/vae.py:347-371, 573-577. Returns [B, 2*z, h, w]."""
    p = "encoder."
    h = F.conv2d(x.float(), sd[p + "conv_in.weight"], sd[p + "conv_in.bias"], padding=1)
    for lvl in _indices(sd, p + "down."):
        for j in _indices(sd, f"{p}down.{lvl}.block."):
            h = vae_resnet(sd, f"{p}down.{lvl}.block.{j}.", h)
        if f"{p}down.{lvl}.downsample.conv.weight" in sd:      # asymmetric pad, vae.py:51-55
            h = F.pad(h, (0, 1, 0, 1))
            h = F.conv2d(h, sd[f"{p}down.{lvl}.downsample.conv.weight"],
                         sd[f"{p}down.{lvl}.downsample.conv.bias"], stride=2)
    h = vae_resnet(sd, p + "mid.block_1.", h)
    h = vae_attn(sd, p + "mid.attn_1.", h)
    h = vae_resnet(sd, p + "mid.block_2.", h)
    h = F.silu(group_norm(h, sd[p + "norm_out.weight"], sd[p + "norm_out.bias"], 1e-6))
    h = F.conv2d(h, sd[p + "conv_out.weight"], sd[p + "conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def vae_encode_mode(sd: SD, image, scale_factor: float = 0.18215):
    """ControlLDM.vae_encode(sample=False) — model/cldm.py:92-119: posterior mean * scale."""
    m = vae_encode_moments(sd, image)
    return m[:, : m.shape[1] // 2] * scale_factor


# --------------------------------------------------------------------------------------
# CLIP text tower (supporting; model/clip.py:37-59, open_clip/transformer.py:199-254)
# --------------------------------------------------------------------------------------
def clip_text_encode(sd: SD, tokens: torch.LongTensor, heads: int = 16, skip_last: int = 1):
    """Penultimate-layer text features [B, 77, width]; keys relative to 'model.'."""
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"]
    b, n, c = x.shape
    mask = torch.full((n, n), float("-inf"), device=x.device).triu_(1)
    layers = _indices(sd, "transformer.resblocks.")
    for i in layers[: len(layers) - skip_last]:
        p = f"transformer.resblocks.{i}."
        h = F.layer_norm(x, (c,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = F.linear(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = qkv.chunk(3, dim=-1)
        dh = c // heads

        def split(t):
            return t.reshape(b, n, heads, dh).permute(0, 2, 1, 3)

        s = torch.matmul(split(q), split(k).transpose(-1, -2)) * (dh ** -0.5) + mask
        o = torch.matmul(torch.softmax(s, dim=-1), split(v)).permute(0, 2, 1, 3).reshape(b, n, c)
        x = x + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        h = F.layer_norm(x, (c,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = F.gelu(F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    return F.layer_norm(x, (c,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
