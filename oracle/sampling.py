"""ORACLE (test infrastructure): restatement of the reference's diffusion schedule, the spaced
and DDIM samplers, the tiled ("mixture of diffusers") model wrapper, the wavelet colour fix and
the SwinIRPipeline orchestration.  numpy fp64 for the host-side tables (as the reference),
torch fp32 for tensors.  `model` is any callable model(x, t, cond_dict) -> eps/v.

Pinned by tests/test_oracle_golden.py against tables/trajectories dumped from the reference.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ diffusion tables
def make_betas(timesteps=1000, linear_start=0.00085, linear_end=0.0120, zero_snr=False) -> np.ndarray:
    """'linear' schedule = linspace in sqrt space, squared — gaussian_diffusion.py:12-19;
    optional zero-terminal-SNR rescale — :49-72 (done in torch fp64 by the reference)."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2
    if zero_snr:
        ab_sqrt = np.sqrt(np.cumprod(1.0 - betas))
        first, last = ab_sqrt[0].copy(), ab_sqrt[-1].copy()
        ab_sqrt = (ab_sqrt - last) * (first / (first - last))
        ab = ab_sqrt ** 2
        alphas = np.concatenate([ab[:1], ab[1:] / ab[:-1]])
        betas = 1.0 - alphas
    return betas


def q_sample(betas: np.ndarray, x0, t: torch.Tensor, noise):
    """Diffusion.q_sample — gaussian_diffusion.py:124-129 (fp32 tables)."""
    ac = np.cumprod(1.0 - betas)
    a = torch.tensor(np.sqrt(ac), dtype=torch.float32, device=x0.device)[t].view(-1, 1, 1, 1)
    s = torch.tensor(np.sqrt(1.0 - ac), dtype=torch.float32, device=x0.device)[t].view(-1, 1, 1, 1)
    return a * x0 + s * noise


def cfg_scale_at(default: float, model_t: int, rescale: bool) -> float:
    """Sampler.get_cfg_scale — sampler/sampler.py:31-38."""
    if rescale and default > 1:
        return 1 + default * ((1 - math.cos(math.pi * ((1000 - model_t) / 1000) ** 5.0)) / 2)
    return default


# ------------------------------------------------------------------ spaced sampler
def spaced_timesteps(num_timesteps: int, steps: int) -> np.ndarray:
    """space_timesteps(num_timesteps, str(steps)) with one section — spaced_sampler.py:14-64."""
    stride = 1.0 if steps <= 1 else (num_timesteps - 1) / (steps - 1)
    idx, cur = [], 0.0
    for _ in range(steps):
        idx.append(round(cur))
        cur += stride
    return np.array(sorted(set(idx)), dtype=np.int32)


def spaced_tables(betas: np.ndarray, steps: int) -> Dict[str, np.ndarray]:
    """SpacedSampler.make_schedule — spaced_sampler.py:77-116 (fp64, cast to fp32 on use)."""
    ac_train = np.cumprod(1.0 - betas)
    ts = spaced_timesteps(len(betas), steps)
    new_betas, last = [], 1.0
    for i in ts:
        new_betas.append(1 - ac_train[i] / last)
        last = ac_train[i]
    b = np.array(new_betas, dtype=np.float64)
    a = 1.0 - b
    ac = np.cumprod(a)
    ac_prev = np.append(1.0, ac[:-1])
    return dict(
        timesteps=ts,
        sqrt_alphas_cumprod=np.sqrt(ac),
        sqrt_one_minus_alphas_cumprod=np.sqrt(1 - ac),
        sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac),
        sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
        posterior_variance=b * (1.0 - ac_prev) / (1.0 - ac),
        posterior_mean_coef1=b * np.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(a) / (1.0 - ac),
    )


def _f32(tab, i, device):
    return torch.tensor(tab, dtype=torch.float32, device=device)[i]


def spaced_sample(model: Callable, betas: np.ndarray, parameterization: str, steps: int, x_T,
                  cond, uncond, cfg_scale: float, rescale_cfg: bool = False,
                  noises: Optional[List[torch.Tensor]] = None, tiled=False, tile_size=-1,
                  tile_stride=-1, record: Optional[list] = None) -> torch.Tensor:
    """SpacedSampler.sample / p_sample / apply_model — spaced_sampler.py:144-245.
    `noises[i]` replaces the i-th torch.randn_like draw (the reference draws one per step,
    the last one multiplied by 0)."""
    tb = spaced_tables(betas, steps)
    fwd = make_tiled_fn(model, tile_size, tile_stride) if tiled else model
    x = x_T
    dev = x.device
    total = len(tb["timesteps"])
    for i, step in enumerate(tb["timesteps"][::-1]):
        t_idx = total - i - 1
        model_t = torch.full((x.shape[0],), int(step), device=dev, dtype=torch.long)
        s = cfg_scale_at(cfg_scale, int(step), rescale_cfg)
        if uncond is None or s == 1.0:
            out = fwd(x, model_t, cond)
        else:
            oc = fwd(x, model_t, cond)
            ou = fwd(x, model_t, uncond)
            out = ou + s * (oc - ou)
        if parameterization == "eps":
            x0 = _f32(tb["sqrt_recip_alphas_cumprod"], t_idx, dev) * x - _f32(tb["sqrt_recipm1_alphas_cumprod"], t_idx, dev) * out
        else:
            x0 = _f32(tb["sqrt_alphas_cumprod"], t_idx, dev) * x - _f32(tb["sqrt_one_minus_alphas_cumprod"], t_idx, dev) * out
        mean = _f32(tb["posterior_mean_coef1"], t_idx, dev) * x0 + _f32(tb["posterior_mean_coef2"], t_idx, dev) * x
        var = _f32(tb["posterior_variance"], t_idx, dev)
        noise = noises[i] if noises is not None else torch.randn_like(x)
        x = mean + (0.0 if t_idx == 0 else 1.0) * torch.sqrt(var) * noise
        if record is not None:
            record.append(x.clone())
    return x


# ------------------------------------------------------------------ DDIM sampler
def ddim_tables(betas: np.ndarray, steps: int, eta: float = 0.0) -> Dict[str, np.ndarray]:
    """make_ddim_timesteps('uniform') + make_ddim_sampling_parameters — ddim_sampler.py:13-58."""
    ac = np.cumprod(1.0 - betas)
    c = len(betas) // steps
    ts = np.asarray(list(range(0, len(betas), c))) + 1
    alphas = ac[ts]
    alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return dict(timesteps=ts, alphas=alphas, alphas_prev=alphas_prev, sigmas=sigmas,
                sqrt_alphas=np.sqrt(alphas), sqrt_one_minus_alphas=np.sqrt(1.0 - alphas))


def ddim_sample(model: Callable, betas: np.ndarray, parameterization: str, steps: int, x_T,
                cond, uncond, cfg_scale: float, rescale_cfg: bool = False, eta: float = 0.0,
                noises: Optional[List[torch.Tensor]] = None, tiled=False, tile_size=-1,
                tile_stride=-1, record: Optional[list] = None) -> torch.Tensor:
    """DDIMSampler.sample / p_sample (batched CFG) — ddim_sampler.py:105-203."""
    tb = ddim_tables(betas, steps, eta)
    fwd = make_tiled_fn(model, tile_size, tile_stride) if tiled else model
    x = x_T
    dev = x.device
    total = len(tb["timesteps"])
    for i, step in enumerate(tb["timesteps"][::-1]):
        t_idx = total - i - 1
        model_t = torch.full((x.shape[0],), int(step), device=dev, dtype=torch.long)
        s = cfg_scale_at(cfg_scale, int(step), rescale_cfg)
        if uncond is None or s == 1.0:
            out = fwd(x, model_t, cond)
        else:
            both = {k: torch.cat([cond[k], uncond[k]]) for k in cond}
            oc, ou = fwd(torch.cat([x] * 2), torch.cat([model_t] * 2), both).chunk(2)
            out = ou + s * (oc - ou)
        if parameterization == "eps":
            e_t = out
        else:
            e_t = _f32(tb["sqrt_alphas"], t_idx, dev) * out + _f32(tb["sqrt_one_minus_alphas"], t_idx, dev) * x
        a_t = _f32(tb["alphas"], t_idx, dev)
        a_prev = _f32(tb["alphas_prev"], t_idx, dev)
        sigma = _f32(tb["sigmas"], t_idx, dev)
        x0 = (x - _f32(tb["sqrt_one_minus_alphas"], t_idx, dev) * e_t) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - sigma ** 2).sqrt() * e_t
        noise = noises[i] if noises is not None else torch.randn_like(x)
        x = a_prev.sqrt() * x0 + dir_xt + sigma * noise
        if record is not None:
            record.append(x.clone())
    return x


# ------------------------------------------------------------------ tiling
def sliding_windows(h: int, w: int, size: int, stride: int) -> List[Tuple[int, int, int, int]]:
    """Row-major tile list with the last tile snapped to the border — utils/common.py:123-138."""
    his = list(range(0, h - size + 1, stride))
    if (h - size) % stride != 0:
        his.append(h - size)
    wis = list(range(0, w - size + 1, stride))
    if (w - size) % stride != 0:
        wis.append(w - size)
    return [(hi, hi + size, wi, wi + size) for hi in his for wi in wis]


def gaussian_weights(tw: int, th: int) -> np.ndarray:
    """Separable Gaussian, var 0.01; x-midpoint (W-1)/2 but y-midpoint H/2 — common.py:142-169."""
    var = 0.01
    xs = np.arange(tw)
    xp = np.exp(-(xs - (tw - 1) / 2) ** 2 / (tw * tw) / (2 * var)) / np.sqrt(2 * np.pi * var)
    ys = np.arange(th)
    yp = np.exp(-(ys - th / 2) ** 2 / (th * th) / (2 * var)) / np.sqrt(2 * np.pi * var)
    return np.outer(yp, xp)


def make_tiled_fn(model: Callable, size: int, stride: int) -> Callable:
    """Tiled model wrapper as installed by the samplers — spaced_sampler.py:204-219 around
    utils/common.py:172-232: per call accumulate model(tile) * w and w, then divide."""
    def tiled(x, t, cond):
        b, c, h, w = x.shape
        out = torch.zeros_like(x)
        count = torch.zeros_like(x, dtype=torch.float32)
        wts = torch.tensor(gaussian_weights(size, size)[None, None], dtype=x.dtype, device=x.device)
        for hi, he, wi, we in sliding_windows(h, w, size, stride):
            tile_cond = {"c_txt": cond["c_txt"], "c_img": cond["c_img"][..., hi:he, wi:we]}
            out[..., hi:he, wi:we] += model(x[..., hi:he, wi:we], t, tile_cond) * wts
            count[..., hi:he, wi:we] += wts
        return out / count
    return tiled


def make_tiled_image_fn(fn: Callable, size: int, stride: int, scale: int = 1) -> Callable:
    """Image version (Gaussian weights; scale 1, or an up-scaling `fn` with scale_type "up") of utils/common.py:172-232,
    as used by the tiled stage-1 branches (pipeline.py:389-394, 349-357): row-major windows, accumulate fn(tile) * w and
    w, divide."""
    def tiled(x):
        b, c, h, w = x.shape
        out = torch.zeros((b, c, h * scale, w * scale), dtype=x.dtype, device=x.device)
        count = torch.zeros_like(out, dtype=torch.float32)
        wts = torch.tensor(gaussian_weights(size * scale, size * scale)[None, None], dtype=x.dtype, device=x.device)
        for hi, he, wi, we in sliding_windows(h, w, size, stride):
            out[..., hi * scale:he * scale, wi * scale:we * scale] += fn(x[..., hi:he, wi:we]) * wts
            count[..., hi * scale:he * scale, wi * scale:we * scale] += wts
        return out / count
    return tiled


def apply_cleaner(cleaner: Callable, lq, tiled: bool = False, tile_size: int = 512, tile_stride: int = 256):
    """SwinIRPipeline.apply_cleaner — pipeline.py:371-397 (both branches, the reference's order of
    resize and network in each)."""
    if tiled and (lq.shape[2] < tile_size or lq.shape[3] < tile_size):
        tiled = False
    if not tiled:
        if min(lq.shape[2:]) < 512:
            lq = resize_short_edge(lq, 512)
        h0, w0 = lq.shape[2:]
        return cleaner(pad_to_multiple(lq, 64))[:, :, :h0, :w0]
    out = make_tiled_image_fn(cleaner, tile_size, tile_stride)(lq)
    if min(out.shape[2:]) < 512:
        out = resize_short_edge(out, 512)
    return out


# ------------------------------------------------------------------ colour fix / pipeline glue
def wavelet_blur(img, radius: int):
    """3x3 binomial, dilation = radius, replicate pad — utils/common.py:29-47."""
    k = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]],
                     dtype=img.dtype, device=img.device)[None, None].repeat(3, 1, 1, 1)
    img = F.pad(img, (radius,) * 4, mode="replicate")
    return F.conv2d(img, k, groups=3, dilation=radius)


def wavelet_reconstruction(content, style, levels: int = 5):
    """High frequencies of `content` + low frequencies of `style` — utils/common.py:50-77."""
    def decompose(img):
        high = torch.zeros_like(img)
        for i in range(levels):
            low = wavelet_blur(img, 2 ** i)
            high = high + (img - low)
            img = low
        return high, low
    ch, _ = decompose(content)
    _, sl = decompose(style)
    return ch + sl


def pad_to_multiple(x, m: int):
    """pad_to_multiples_of — pipeline.py:35-40."""
    h, w = x.shape[2:]
    ph, pw = (-h) % m, (-w) % m
    return x.clone() if ph == 0 and pw == 0 else F.pad(x, (0, pw, 0, ph))


def resize_short_edge(x, size: int):
    """resize_short_edge_to — pipeline.py:24-32 (bicubic, antialias)."""
    h, w = x.shape[2:]
    if h == w:
        oh, ow = size, size
    elif h < w:
        oh, ow = size, int(w * (size / h))
    else:
        oh, ow = int(h * (size / w)), size
    return F.interpolate(x, size=(oh, ow), mode="bicubic", antialias=True)


def swinir_pipeline_run(lq_u8: np.ndarray, cleaner: Callable, encode_img: Callable,
                        encode_txt: Callable, decode: Callable, model: Callable,
                        betas: np.ndarray, parameterization: str, *, steps: int, strength: float,
                        pos_prompt: str, neg_prompt: str, cfg_scale: float, sampler: str = "spaced",
                        cldm_tiled: bool = False, cldm_tile_size: int = 512,
                        cldm_tile_stride: int = 256, rescale_cfg: bool = False,
                        cleaner_tiled: bool = False, cleaner_tile_size: int = 512, cleaner_tile_stride: int = 256,
                        x_T: Optional[torch.Tensor] = None,
                        noises: Optional[List[torch.Tensor]] = None, device="cpu",
                        set_strength: Optional[Callable] = None, taps: Optional[dict] = None,
                        stage1: Optional[Callable] = None, out_size: Optional[Tuple[int, int]] = None,
                        sample_fn: Optional[Callable] = None):
    """SwinIRPipeline.run (start_point 'noise', noise_aug 0, un-tiled VAE) —
    pipeline.py:235-321, 71-233, 371-397.  Callables stand for the networks:
    cleaner(img01)->img01, encode_img(img_pm1)->latent, encode_txt(list)->c_txt,
    decode(latent)->img_pm1, model(x,t,cond)->eps."""
    lq = torch.tensor(lq_u8, dtype=torch.float32, device=device).div(255).clamp(0, 1).permute(0, 3, 1, 2).contiguous()
    # stage1 / out_size: another pipeline's apply_cleaner / set_output_size (BSRNetPipeline, pipeline.py:339-366)
    out_size = tuple(lq.shape[2:]) if out_size is None else out_size
    clean = (apply_cleaner(cleaner, lq, cleaner_tiled, cleaner_tile_size, cleaner_tile_stride) if stage1 is None
             else stage1(lq))
    bs = clean.shape[0]
    cond_img = pad_to_multiple(clean, 8 if cldm_tiled else 64)
    cond = dict(c_txt=encode_txt([pos_prompt] * bs), c_img=encode_img(cond_img * 2 - 1))
    uncond = dict(c_txt=encode_txt([neg_prompt] * bs), c_img=encode_img(cond_img * 2 - 1))
    h1, w1 = cond["c_img"].shape[2:]
    if cldm_tiled and (h1 < cldm_tile_size // 8 or w1 < cldm_tile_size // 8):
        cldm_tiled = False
    if not cldm_tiled:
        cond["c_img"] = pad_to_multiple(cond["c_img"], 8)
        uncond["c_img"] = pad_to_multiple(uncond["c_img"], 8)
    h2, w2 = cond["c_img"].shape[2:]
    if x_T is None:
        x_T = torch.randn((bs, 4, h2, w2), dtype=torch.float32, device=device)
    if set_strength is not None:
        set_strength(strength)
    # sample_fn(model, x_T, cond, uncond) -> z: another sampler family's loop (pipeline.py:186-218 picks it by name)
    if sample_fn is not None:
        z = sample_fn(model, x_T, cond, uncond)
    else:
        fn = spaced_sample if sampler == "spaced" else ddim_sample
        z = fn(model, betas, parameterization, steps, x_T, cond, uncond, cfg_scale,
               rescale_cfg=rescale_cfg, noises=noises, tiled=cldm_tiled,
               tile_size=cldm_tile_size // 8, tile_stride=cldm_tile_stride // 8)
    z = z[..., :h1, :w1]
    x = decode(z)[:, :, : clean.shape[2], : clean.shape[3]]
    if taps is not None:
        taps.update(clean=clean, z=z, decoded=x, cond=cond, uncond=uncond)
    sample = F.interpolate(wavelet_reconstruction((x + 1) / 2, clean), size=out_size,
                           mode="bicubic", antialias=True)
    return (sample * 255.0).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().cpu().numpy()
