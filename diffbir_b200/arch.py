"""Architecture descriptions: block plans and parameter shapes, in the reference's state_dict
key layout, derived from the reference's YAML configs (configs/inference/*.yaml).

The block plan is what the engine walks to enqueue kernels; the shape tables drive the
synthetic-checkpoint generator (bench / tests use random-init weights of the real
architecture) and the strict checkpoint loaders.  Mirrors the constructor logic of
UNetModel (model/unet.py:412-682), ControlNet (model/controlnet.py:52-312),
Encoder/Decoder (model/vae.py:306-345, 456-524), SwinIR (model/swinir.py:648-812) and the
OpenCLIP text tower (model/open_clip/transformer.py).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

Shapes = "OrderedDict[str, Tuple[int, ...]]"

UNET_CFG = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1),
                num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_head_channels=64,
                transformer_depth=1, context_dim=1024)
CONTROLNET_CFG = dict(UNET_CFG, hint_channels=4)
VAE_CFG = dict(embed_dim=4, z_channels=4, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4, 4),
               num_res_blocks=2)
SWINIR_CFG = dict(img_size=64, in_chans=3, embed_dim=180, depths=(6,) * 8, num_heads=(6,) * 8,
                  window_size=8, mlp_ratio=2, sf=8, img_range=1.0, unshuffle_scale=8, num_feat=64)
CLIP_TEXT_CFG = dict(context_length=77, vocab_size=49408, width=1024, heads=16, layers=24,
                     embed_dim=1024)


# ------------------------------------------------------------------------------- UNet plan
@dataclass
class Layer:
    kind: str                 # conv_in | res | attn | down | up
    prefix: str               # state_dict prefix, ends with '.'
    cin: int
    cout: int


@dataclass
class Block:
    prefix: str
    layers: List[Layer] = field(default_factory=list)
    cout: int = 0


@dataclass
class UNetPlan:
    cfg: dict
    input_blocks: List[Block]
    middle: Block
    output_blocks: List[Block]            # empty for ControlNet
    skip_channels: List[int]              # channels of hs[i] (== zero-conv widths, + middle)
    is_controlnet: bool


def unet_plan(cfg: dict, controlnet: bool = False) -> UNetPlan:
    mc = cfg["model_channels"]
    mult = tuple(cfg["channel_mult"])
    nres = cfg["num_res_blocks"]
    attn_res = tuple(cfg["attention_resolutions"])
    in_ch = cfg["in_channels"] + (cfg.get("hint_channels", 0) if controlnet else 0)
    blocks = [Block("input_blocks.0.", [Layer("conv_in", "input_blocks.0.0.", in_ch, mc)], mc)]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nres):
            i = len(blocks)
            p = f"input_blocks.{i}."
            layers = [Layer("res", p + "0.", ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                layers.append(Layer("attn", p + "1.", ch, ch))
            blocks.append(Block(p, layers, ch))
            chans.append(ch)
        if level != len(mult) - 1:
            i = len(blocks)
            p = f"input_blocks.{i}."
            blocks.append(Block(p, [Layer("down", p + "0.", ch, ch)], ch))
            chans.append(ch)
            ds *= 2
    middle = Block("middle_block.", [Layer("res", "middle_block.0.", ch, ch),
                                     Layer("attn", "middle_block.1.", ch, ch),
                                     Layer("res", "middle_block.2.", ch, ch)], ch)
    skip_channels = list(chans) + [ch]
    outs: List[Block] = []
    if not controlnet:
        stack = list(chans)
        for level, m in list(enumerate(mult))[::-1]:
            for i in range(nres + 1):
                ich = stack.pop()
                k = len(outs)
                p = f"output_blocks.{k}."
                layers = [Layer("res", p + "0.", ch + ich, mc * m)]
                ch = mc * m
                if ds in attn_res:
                    layers.append(Layer("attn", p + f"{len(layers)}.", ch, ch))
                if level and i == nres:
                    layers.append(Layer("up", p + f"{len(layers)}.", ch, ch))
                    ds //= 2
                outs.append(Block(p, layers, ch))
    return UNetPlan(cfg, blocks, middle, outs, skip_channels, controlnet)


def _res_shapes(s: Shapes, p: str, cin: int, cout: int, emb: int):
    s[p + "in_layers.0.weight"] = (cin,)
    s[p + "in_layers.0.bias"] = (cin,)
    s[p + "in_layers.2.weight"] = (cout, cin, 3, 3)
    s[p + "in_layers.2.bias"] = (cout,)
    s[p + "emb_layers.1.weight"] = (cout, emb)
    s[p + "emb_layers.1.bias"] = (cout,)
    s[p + "out_layers.0.weight"] = (cout,)
    s[p + "out_layers.0.bias"] = (cout,)
    s[p + "out_layers.3.weight"] = (cout, cout, 3, 3)
    s[p + "out_layers.3.bias"] = (cout,)
    if cin != cout:
        s[p + "skip_connection.weight"] = (cout, cin, 1, 1)
        s[p + "skip_connection.bias"] = (cout,)


def _attn_shapes(s: Shapes, p: str, c: int, ctx: int, depth: int):
    s[p + "norm.weight"] = (c,)
    s[p + "norm.bias"] = (c,)
    s[p + "proj_in.weight"] = (c, c)
    s[p + "proj_in.bias"] = (c,)
    for d in range(depth):
        q = f"{p}transformer_blocks.{d}."
        s[q + "attn1.to_q.weight"] = (c, c)
        s[q + "attn1.to_k.weight"] = (c, c)
        s[q + "attn1.to_v.weight"] = (c, c)
        s[q + "attn1.to_out.0.weight"] = (c, c)
        s[q + "attn1.to_out.0.bias"] = (c,)
        s[q + "ff.net.0.proj.weight"] = (8 * c, c)
        s[q + "ff.net.0.proj.bias"] = (8 * c,)
        s[q + "ff.net.2.weight"] = (c, 4 * c)
        s[q + "ff.net.2.bias"] = (c,)
        s[q + "attn2.to_q.weight"] = (c, c)
        s[q + "attn2.to_k.weight"] = (c, ctx)
        s[q + "attn2.to_v.weight"] = (c, ctx)
        s[q + "attn2.to_out.0.weight"] = (c, c)
        s[q + "attn2.to_out.0.bias"] = (c,)
        for nm in ("norm1", "norm2", "norm3"):
            s[q + nm + ".weight"] = (c,)
            s[q + nm + ".bias"] = (c,)
    s[p + "proj_out.weight"] = (c, c)
    s[p + "proj_out.bias"] = (c,)


def unet_shapes(cfg: dict, controlnet: bool = False) -> Shapes:
    plan = unet_plan(cfg, controlnet)
    mc, emb = cfg["model_channels"], cfg["model_channels"] * 4
    ctx, depth = cfg["context_dim"], cfg.get("transformer_depth", 1)
    s: Shapes = OrderedDict()
    s["time_embed.0.weight"] = (emb, mc)
    s["time_embed.0.bias"] = (emb,)
    s["time_embed.2.weight"] = (emb, emb)
    s["time_embed.2.bias"] = (emb,)

    def add(block: Block):
        for l in block.layers:
            if l.kind == "conv_in":
                s[l.prefix + "weight"] = (l.cout, l.cin, 3, 3)
                s[l.prefix + "bias"] = (l.cout,)
            elif l.kind == "res":
                _res_shapes(s, l.prefix, l.cin, l.cout, emb)
            elif l.kind == "attn":
                _attn_shapes(s, l.prefix, l.cin, ctx, depth)
            elif l.kind == "down":
                s[l.prefix + "op.weight"] = (l.cout, l.cin, 3, 3)
                s[l.prefix + "op.bias"] = (l.cout,)
            elif l.kind == "up":
                s[l.prefix + "conv.weight"] = (l.cout, l.cin, 3, 3)
                s[l.prefix + "conv.bias"] = (l.cout,)

    for i, b in enumerate(plan.input_blocks):
        add(b)
        if controlnet:
            s[f"zero_convs.{i}.0.weight"] = (b.cout, b.cout, 1, 1)
            s[f"zero_convs.{i}.0.bias"] = (b.cout,)
    add(plan.middle)
    if controlnet:
        s["middle_block_out.0.weight"] = (plan.middle.cout, plan.middle.cout, 1, 1)
        s["middle_block_out.0.bias"] = (plan.middle.cout,)
    else:
        for b in plan.output_blocks:
            add(b)
        s["out.0.weight"] = (mc,)
        s["out.0.bias"] = (mc,)
        s["out.2.weight"] = (cfg["out_channels"], mc, 3, 3)
        s["out.2.bias"] = (cfg["out_channels"],)
    return s


# Tensors the reference constructors zero-initialise (zero_module): re-randomised in synthetic
# checkpoints so parity checks are not vacuous (SURVEY.md headline fact 5).
def is_zero_init(key: str) -> bool:
    return (".out_layers.3." in key or ".proj_out." in key or key.startswith("out.2.")
            or key.startswith("zero_convs.") or key.startswith("middle_block_out."))


# ------------------------------------------------------------------------------- VAE
def _vae_res(s: Shapes, p: str, cin: int, cout: int):
    s[p + "norm1.weight"] = (cin,)
    s[p + "norm1.bias"] = (cin,)
    s[p + "conv1.weight"] = (cout, cin, 3, 3)
    s[p + "conv1.bias"] = (cout,)
    s[p + "norm2.weight"] = (cout,)
    s[p + "norm2.bias"] = (cout,)
    s[p + "conv2.weight"] = (cout, cout, 3, 3)
    s[p + "conv2.bias"] = (cout,)
    if cin != cout:
        s[p + "nin_shortcut.weight"] = (cout, cin, 1, 1)
        s[p + "nin_shortcut.bias"] = (cout,)


def _vae_attn(s: Shapes, p: str, c: int):
    s[p + "norm.weight"] = (c,)
    s[p + "norm.bias"] = (c,)
    for nm in ("q", "k", "v", "proj_out"):
        s[p + nm + ".weight"] = (c, c, 1, 1)
        s[p + nm + ".bias"] = (c,)


def vae_shapes(cfg: dict) -> Shapes:
    ch, mult, nres = cfg["ch"], tuple(cfg["ch_mult"]), cfg["num_res_blocks"]
    z = cfg["z_channels"]
    s: Shapes = OrderedDict()
    # encoder
    p = "encoder."
    s[p + "conv_in.weight"] = (ch, cfg["in_channels"], 3, 3)
    s[p + "conv_in.bias"] = (ch,)
    in_mult = (1,) + mult
    block_in = ch
    for lvl in range(len(mult)):
        block_in = ch * in_mult[lvl]
        block_out = ch * mult[lvl]
        for j in range(nres):
            _vae_res(s, f"{p}down.{lvl}.block.{j}.", block_in, block_out)
            block_in = block_out
        if lvl != len(mult) - 1:
            s[f"{p}down.{lvl}.downsample.conv.weight"] = (block_in, block_in, 3, 3)
            s[f"{p}down.{lvl}.downsample.conv.bias"] = (block_in,)
    _vae_res(s, p + "mid.block_1.", block_in, block_in)
    _vae_attn(s, p + "mid.attn_1.", block_in)
    _vae_res(s, p + "mid.block_2.", block_in, block_in)
    s[p + "norm_out.weight"] = (block_in,)
    s[p + "norm_out.bias"] = (block_in,)
    s[p + "conv_out.weight"] = (2 * z, block_in, 3, 3)
    s[p + "conv_out.bias"] = (2 * z,)
    # decoder
    p = "decoder."
    block_in = ch * mult[-1]
    s[p + "conv_in.weight"] = (block_in, z, 3, 3)
    s[p + "conv_in.bias"] = (block_in,)
    _vae_res(s, p + "mid.block_1.", block_in, block_in)
    _vae_attn(s, p + "mid.attn_1.", block_in)
    _vae_res(s, p + "mid.block_2.", block_in, block_in)
    for lvl in reversed(range(len(mult))):
        block_out = ch * mult[lvl]
        for j in range(nres + 1):
            _vae_res(s, f"{p}up.{lvl}.block.{j}.", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            s[f"{p}up.{lvl}.upsample.conv.weight"] = (block_in, block_in, 3, 3)
            s[f"{p}up.{lvl}.upsample.conv.bias"] = (block_in,)
    s[p + "norm_out.weight"] = (block_in,)
    s[p + "norm_out.bias"] = (block_in,)
    s[p + "conv_out.weight"] = (cfg["out_ch"], block_in, 3, 3)
    s[p + "conv_out.bias"] = (cfg["out_ch"],)
    s["quant_conv.weight"] = (2 * cfg["embed_dim"], 2 * z, 1, 1)
    s["quant_conv.bias"] = (2 * cfg["embed_dim"],)
    s["post_quant_conv.weight"] = (z, cfg["embed_dim"], 1, 1)
    s["post_quant_conv.bias"] = (z,)
    return s


# ------------------------------------------------------------------------------- SwinIR
def swinir_shapes(cfg: dict) -> Shapes:
    """Parameters only; the reference state_dict additionally carries the int64
    relative_position_index and fp32 attn_mask buffers (see swinir_buffers)."""
    c, ws = cfg["embed_dim"], cfg["window_size"]
    nf = cfg.get("num_feat", 64)
    hidden = int(c * cfg["mlp_ratio"])
    in_ch = cfg["in_chans"] * cfg["unshuffle_scale"] ** 2
    s: Shapes = OrderedDict()
    s["conv_first.1.weight"] = (c, in_ch, 3, 3)
    s["conv_first.1.bias"] = (c,)
    s["patch_embed.norm.weight"] = (c,)
    s["patch_embed.norm.bias"] = (c,)
    for li, depth in enumerate(cfg["depths"]):
        heads = cfg["num_heads"][li]
        for bi in range(depth):
            p = f"layers.{li}.residual_group.blocks.{bi}."
            s[p + "norm1.weight"] = (c,)
            s[p + "norm1.bias"] = (c,)
            s[p + "attn.relative_position_bias_table"] = ((2 * ws - 1) ** 2, heads)
            s[p + "attn.qkv.weight"] = (3 * c, c)
            s[p + "attn.qkv.bias"] = (3 * c,)
            s[p + "attn.proj.weight"] = (c, c)
            s[p + "attn.proj.bias"] = (c,)
            s[p + "norm2.weight"] = (c,)
            s[p + "norm2.bias"] = (c,)
            s[p + "mlp.fc1.weight"] = (hidden, c)
            s[p + "mlp.fc1.bias"] = (hidden,)
            s[p + "mlp.fc2.weight"] = (c, hidden)
            s[p + "mlp.fc2.bias"] = (c,)
        s[f"layers.{li}.conv.weight"] = (c, c, 3, 3)
        s[f"layers.{li}.conv.bias"] = (c,)
    s["norm.weight"] = (c,)
    s["norm.bias"] = (c,)
    s["conv_after_body.weight"] = (c, c, 3, 3)
    s["conv_after_body.bias"] = (c,)
    s["conv_before_upsample.0.weight"] = (nf, c, 3, 3)
    s["conv_before_upsample.0.bias"] = (nf,)
    ups = {2: 1, 4: 2, 8: 3}[cfg["sf"]]
    for i in range(1, ups + 1):
        s[f"conv_up{i}.weight"] = (nf, nf, 3, 3)
        s[f"conv_up{i}.bias"] = (nf,)
    s["conv_hr.weight"] = (nf, nf, 3, 3)
    s["conv_hr.bias"] = (nf,)
    s["conv_last.weight"] = (cfg["in_chans"], nf, 3, 3)
    s["conv_last.bias"] = (cfg["in_chans"],)
    return s


# ------------------------------------------------------------------------------- CLIP text
def clip_text_shapes(cfg: dict) -> Shapes:
    w = cfg["width"]
    s: Shapes = OrderedDict()
    s["positional_embedding"] = (cfg["context_length"], w)
    s["text_projection"] = (w, cfg["embed_dim"])
    s["logit_scale"] = ()
    for i in range(cfg["layers"]):
        p = f"transformer.resblocks.{i}."
        s[p + "ln_1.weight"] = (w,)
        s[p + "ln_1.bias"] = (w,)
        s[p + "attn.in_proj_weight"] = (3 * w, w)
        s[p + "attn.in_proj_bias"] = (3 * w,)
        s[p + "attn.out_proj.weight"] = (w, w)
        s[p + "attn.out_proj.bias"] = (w,)
        s[p + "ln_2.weight"] = (w,)
        s[p + "ln_2.bias"] = (w,)
        s[p + "mlp.c_fc.weight"] = (4 * w, w)
        s[p + "mlp.c_fc.bias"] = (4 * w,)
        s[p + "mlp.c_proj.weight"] = (w, 4 * w)
        s[p + "mlp.c_proj.bias"] = (w,)
    s["token_embedding.weight"] = (cfg["vocab_size"], w)
    s["ln_final.weight"] = (w,)
    s["ln_final.bias"] = (w,)
    return s


# ------------------------------------------------------------------------------- RRDBNet (BSRNet)
RRDBNET_CFG = dict(in_nc=3, out_nc=3, nf=64, nb=23, gc=32, sf=4)      # configs/inference/bsrnet.yaml


def rrdbnet_shapes(cfg: dict) -> Shapes:
    """state_dict of the reference's RRDBNet (model/bsrnet.py:36-104): conv_first, nb x RRDB of three
    5-conv dense blocks, trunk_conv, upconv1 (+ upconv2 when sf == 4), HRconv, conv_last."""
    nf, gc = cfg["nf"], cfg["gc"]
    s: Shapes = OrderedDict()
    s["conv_first.weight"] = (nf, cfg["in_nc"], 3, 3)
    s["conv_first.bias"] = (nf,)
    for b in range(cfg["nb"]):
        for r in (1, 2, 3):
            p = f"RRDB_trunk.{b}.RDB{r}."
            for k in range(1, 5):
                s[p + f"conv{k}.weight"] = (gc, nf + (k - 1) * gc, 3, 3)
                s[p + f"conv{k}.bias"] = (gc,)
            s[p + "conv5.weight"] = (nf, nf + 4 * gc, 3, 3)
            s[p + "conv5.bias"] = (nf,)
    names = ["trunk_conv", "upconv1"] + (["upconv2"] if cfg["sf"] == 4 else []) + ["HRconv"]
    for nm in names:
        s[nm + ".weight"] = (nf, nf, 3, 3)
        s[nm + ".bias"] = (nf,)
    s["conv_last.weight"] = (cfg["out_nc"], nf, 3, 3)
    s["conv_last.bias"] = (cfg["out_nc"],)
    return s


# ------------------------------------------------------------------------------- SCUNet
SCUNET_CFG = dict(in_nc=3, config=(4, 4, 4, 4, 4, 4, 4), dim=64)       # configs/inference/scunet.yaml


def scunet_stages(cfg: dict):
    """(module name, channels = conv_dim + trans_dim, number of ConvTransBlocks, first index inside the module)
    for the seven stages of SCUNet (model/scunet.py:176-212); the up stages start with their ConvTranspose2d."""
    d, n = cfg["dim"], cfg["config"]
    return [("m_down1", d, n[0], 0), ("m_down2", 2 * d, n[1], 0), ("m_down3", 4 * d, n[2], 0), ("m_body", 8 * d, n[3], 0),
            ("m_up3", 4 * d, n[4], 1), ("m_up2", 2 * d, n[5], 1), ("m_up1", d, n[6], 1)]


def scunet_shapes(cfg: dict) -> Shapes:
    """state_dict of the reference's SCUNet (model/scunet.py:163-219)."""
    d = cfg["dim"]
    s: Shapes = OrderedDict()
    s["m_head.0.weight"] = (d, cfg["in_nc"], 3, 3)
    for name, c, nblk, first in scunet_stages(cfg):
        if first:                                                  # ConvTranspose2d(2c -> c, 2, 2): [Cin, Cout, 2, 2]
            s[f"{name}.0.weight"] = (2 * c, c, 2, 2)
        t = c // 2                                                 # conv_dim == trans_dim
        for i in range(nblk):
            p = f"{name}.{first + i}."
            s[p + "trans_block.ln1.weight"] = (t,)
            s[p + "trans_block.ln1.bias"] = (t,)
            s[p + "trans_block.msa.relative_position_params"] = (t // 32, 15, 15)
            s[p + "trans_block.msa.embedding_layer.weight"] = (3 * t, t)
            s[p + "trans_block.msa.embedding_layer.bias"] = (3 * t,)
            s[p + "trans_block.msa.linear.weight"] = (t, t)
            s[p + "trans_block.msa.linear.bias"] = (t,)
            s[p + "trans_block.ln2.weight"] = (t,)
            s[p + "trans_block.ln2.bias"] = (t,)
            s[p + "trans_block.mlp.0.weight"] = (4 * t, t)
            s[p + "trans_block.mlp.0.bias"] = (4 * t,)
            s[p + "trans_block.mlp.2.weight"] = (t, 4 * t)
            s[p + "trans_block.mlp.2.bias"] = (t,)
            s[p + "conv1_1.weight"] = (c, c, 1, 1)
            s[p + "conv1_1.bias"] = (c,)
            s[p + "conv1_2.weight"] = (c, c, 1, 1)
            s[p + "conv1_2.bias"] = (c,)
            s[p + "conv_block.0.weight"] = (t, t, 3, 3)
            s[p + "conv_block.2.weight"] = (t, t, 3, 3)
        if name.startswith("m_down"):                              # Conv2d(c -> 2c, 2, 2)
            s[f"{name}.{nblk}.weight"] = (2 * c, c, 2, 2)
    s["m_tail.0.weight"] = (cfg["in_nc"], d, 3, 3)
    return s
