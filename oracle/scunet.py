"""CPU/GPU fp32 restatement of the reference's SCUNet (diffbir/model/scunet.py:9-243) and of
SCUNetPipeline.apply_cleaner (diffbir/pipeline.py:400-420), driven by the reference's state_dict keys.
TEST INFRASTRUCTURE: imported by tests/ only; pinned by tests/golden/scunet_small.npz, which the reference
module itself produced (gen_golden_scunet.py)."""
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
WS, HD = 8, 32


def _rel_index() -> torch.Tensor:
    cord = torch.tensor([[i, j] for i in range(WS) for j in range(WS)])
    return cord[:, None, :] - cord[None, :, :] + WS - 1                       # [64, 64, 2]


def _wmsa(sd: SD, p: str, x: torch.Tensor, shifted: bool) -> torch.Tensor:
    """WMSA.forward — scunet.py:54-91. x [b, h, w, c]."""
    b, h, w, c = x.shape
    heads = c // HD
    if shifted:
        x = torch.roll(x, shifts=(-(WS // 2), -(WS // 2)), dims=(1, 2))
    hw_, ww_ = h // WS, w // WS
    x = x.view(b, hw_, WS, ww_, WS, c).permute(0, 1, 3, 2, 4, 5).reshape(b, hw_ * ww_, WS * WS, c)
    qkv = F.linear(x, sd[p + "embedding_layer.weight"], sd[p + "embedding_layer.bias"])
    qkv = qkv.view(b, hw_ * ww_, WS * WS, 3 * heads, HD).permute(3, 0, 1, 2, 4)   # [3h, b, nw, np, c]
    q, k, v = qkv[:heads], qkv[heads:2 * heads], qkv[2 * heads:]
    sim = torch.einsum("hbwpc,hbwqc->hbwpq", q, k) * HD ** -0.5
    rel = _rel_index()
    sim = sim + sd[p + "relative_position_params"][:, rel[:, :, 0], rel[:, :, 1]][:, None, None]
    if shifted:                                                               # generate_mask, scunet.py:33-52
        m = torch.zeros(hw_, ww_, WS, WS, WS, WS, dtype=torch.bool, device=x.device)
        s = WS - WS // 2
        m[-1, :, :s, :, s:, :] = True
        m[-1, :, s:, :, :s, :] = True
        m[:, -1, :, :s, :, s:] = True
        m[:, -1, :, s:, :, :s] = True
        sim = sim.masked_fill(m.reshape(1, 1, hw_ * ww_, WS * WS, WS * WS), float("-inf"))
    out = torch.einsum("hbwij,hbwjc->hbwic", F.softmax(sim, dim=-1), v)
    out = out.permute(1, 2, 3, 0, 4).reshape(b, hw_ * ww_, WS * WS, c)
    out = F.linear(out, sd[p + "linear.weight"], sd[p + "linear.bias"])
    out = out.view(b, hw_, ww_, WS, WS, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)
    if shifted:
        out = torch.roll(out, shifts=(WS // 2, WS // 2), dims=(1, 2))
    return out


def _block(sd: SD, p: str, x: torch.Tensor, shifted: bool) -> torch.Tensor:
    """ConvTransBlock.forward — scunet.py:147-156 (Block: :120-123)."""
    c = x.shape[1]
    t = c // 2
    y = F.conv2d(x, sd[p + "conv1_1.weight"], sd[p + "conv1_1.bias"])
    cx, tx = y[:, :t], y[:, t:]
    cx = F.conv2d(F.relu(F.conv2d(cx, sd[p + "conv_block.0.weight"], padding=1)), sd[p + "conv_block.2.weight"], padding=1) + cx
    tx = tx.permute(0, 2, 3, 1)
    q = p + "trans_block."
    tx = tx + _wmsa(sd, q + "msa.", F.layer_norm(tx, (t,), sd[q + "ln1.weight"], sd[q + "ln1.bias"]), shifted)
    hdn = F.gelu(F.linear(F.layer_norm(tx, (t,), sd[q + "ln2.weight"], sd[q + "ln2.bias"]), sd[q + "mlp.0.weight"], sd[q + "mlp.0.bias"]))
    tx = tx + F.linear(hdn, sd[q + "mlp.2.weight"], sd[q + "mlp.2.bias"])
    res = F.conv2d(torch.cat((cx, tx.permute(0, 3, 1, 2)), 1), sd[p + "conv1_2.weight"], sd[p + "conv1_2.bias"])
    return x + res


def scunet_forward(sd: SD, x0: torch.Tensor) -> torch.Tensor:
    """SCUNet.forward — scunet.py:221-243 (block types 'W' / 'SW' alternate, :176-209; input_resolution 256)."""
    def nblocks(name, first):
        return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith(name + ".") and "conv1_1" in k) - first

    def stage(name, x, first):
        for i in range(nblocks(name, first)):
            x = _block(sd, f"{name}.{first + i}.", x, shifted=bool(i % 2))
        return x

    h, w = x0.shape[-2:]
    pb, pr = int(np.ceil(h / 64) * 64 - h), int(np.ceil(w / 64) * 64 - w)
    x0 = F.pad(x0.float(), (0, pr, 0, pb), mode="replicate")
    x1 = F.conv2d(x0, sd["m_head.0.weight"], padding=1)
    downs, x = [x1], x1
    for name in ("m_down1", "m_down2", "m_down3"):
        x = stage(name, x, 0)
        x = F.conv2d(x, sd[f"{name}.{nblocks(name, 0)}.weight"], stride=2)
        downs.append(x)
    x1, x2, x3, x4 = downs
    x = stage("m_body", x4, 0)
    for name, skip in (("m_up3", x4), ("m_up2", x3), ("m_up1", x2)):
        x = F.conv_transpose2d(x + skip, sd[name + ".0.weight"], stride=2)
        x = stage(name, x, 1)
    x = F.conv2d(x + x1, sd["m_tail.0.weight"], padding=1)
    return x[..., :h, :w]


def scunet_apply_cleaner(cleaner, lq: torch.Tensor) -> torch.Tensor:
    """SCUNetPipeline.apply_cleaner, un-tiled branch (pipeline.py:400-420)."""
    out = cleaner(lq)
    hh, ww = out.shape[2:]
    if min(hh, ww) < 512:
        size = (512, 512) if hh == ww else ((512, int(ww * (512 / hh))) if hh < ww else (int(hh * (512 / ww)), 512))
        out = F.interpolate(out, size=size, mode="bicubic", antialias=True)
    return out
