"""Generates the golden fixtures in this directory by running the REFERENCE implementation
(/root/reference, imported read-only through oracle/_shims) on seeded synthetic weights/inputs.

    python tests/golden/gen_golden.py

The reference ships no tests or golden vectors, so these fixtures are what pins
the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Only this script
needs /root/reference; the fixtures travel with the repo.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_shims"))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from _ref_import import use_reference  # noqa: E402

use_reference()      # `import diffbir` = the reference checkout, not this repo's alias package

from diffbir_b200 import arch  # noqa: E402
from diffbir_b200.utils.synth import make_state_dict  # noqa: E402
from tests.small_cfg import CLIP_SMALL, SWIN_SMALL, UNET_SMALL, CN_SMALL, VAE_SMALL  # noqa: E402

OUT = Path(__file__).resolve().parent


def rnd(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@torch.no_grad()
def main():
    import diffbir.model as M
    from diffbir.model.controlnet import ControlledUnetModel, ControlNet
    from diffbir.model.vae import AutoencoderKL
    from diffbir.model.swinir import SwinIR
    from diffbir.model.gaussian_diffusion import Diffusion
    from diffbir.sampler import SpacedSampler, DDIMSampler
    from diffbir.utils.common import (gaussian_weights, make_tiled_fn, sliding_windows,
                                      wavelet_reconstruction)
    from diffbir.pipeline import pad_to_multiples_of, resize_short_edge_to

    ref_kw = dict(use_checkpoint=False, image_size=32, use_spatial_transformer=True,
                  use_linear_in_transformer=True, legacy=False)

    # ---------------- ControlNet + UNet (ControlLDM.forward semantics) ----------------
    ucfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in UNET_SMALL.items()}
    ccfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in CN_SMALL.items() if k != "out_channels"}
    unet = ControlledUnetModel(**ucfg, **ref_kw).eval()
    cnet = ControlNet(**ccfg, **ref_kw).eval()
    usd = make_state_dict(arch.unet_shapes(UNET_SMALL), 1, arch.is_zero_init)
    csd = make_state_dict(arch.unet_shapes(CN_SMALL, True), 2, arch.is_zero_init)
    unet.load_state_dict(usd, strict=True)
    cnet.load_state_dict(csd, strict=True)
    L = 16
    x = rnd(10, 2, 4, L, L)
    hint = rnd(11, 2, 4, L, L)
    ctx = rnd(12, 2, 77, UNET_SMALL["context_dim"])
    t = torch.tensor([981, 981])
    scales = [0.9 + 0.01 * i for i in range(13)]
    control = cnet(x=x, hint=hint, timesteps=t, context=ctx)
    control_s = [c * s for c, s in zip(control, scales)]
    eps = unet(x=x, timesteps=t, context=ctx, control=[c.clone() for c in control_s], only_mid_control=False)
    np.savez_compressed(OUT / "cldm_small.npz", x=x.numpy(), hint=hint.numpy(), ctx=ctx.numpy(),
                        t=t.numpy(), scales=np.array(scales, dtype=np.float64), eps=eps.numpy(),
                        control0=control[0].numpy(), control12=control[12].numpy())
    print("cldm eps", eps.abs().mean().item(), eps.std().item())

    # ---------------- VAE ----------------
    vcfg = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3,
                              ch=VAE_SMALL["ch"], ch_mult=list(VAE_SMALL["ch_mult"]),
                              num_res_blocks=VAE_SMALL["num_res_blocks"], attn_resolutions=[], dropout=0.0),
                embed_dim=4)
    vae = AutoencoderKL(**vcfg).eval()
    vsd = make_state_dict(arch.vae_shapes(VAE_SMALL), 3)
    vae.load_state_dict(vsd, strict=True)
    z = rnd(20, 1, 4, 8, 8)
    dec = vae.decode(z)
    img = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(21)) * 2 - 1
    mom = vae.encode(img).parameters
    np.savez_compressed(OUT / "vae_small.npz", z=z.numpy(), dec=dec.numpy(), img=img.numpy(), moments=mom.numpy())
    print("vae dec", dec.abs().mean().item(), "moments", mom.abs().mean().item())

    # ---------------- SwinIR ----------------
    scfg = dict(SWIN_SMALL)
    swin = SwinIR(img_size=scfg["img_size"], patch_size=1, in_chans=3, embed_dim=scfg["embed_dim"],
                  depths=list(scfg["depths"]), num_heads=list(scfg["num_heads"]), window_size=8,
                  mlp_ratio=scfg["mlp_ratio"], sf=8, img_range=1.0, upsampler="nearest+conv",
                  resi_connection="1conv", unshuffle=True, unshuffle_scale=8).eval()
    ssd = make_state_dict(arch.swinir_shapes(SWIN_SMALL), 4)
    missing, unexpected = swin.load_state_dict(ssd, strict=False)
    assert not unexpected and all(k.endswith("relative_position_index") or k.endswith("attn_mask") for k in missing), missing
    xin = torch.rand(1, 3, 128, 192, generator=torch.Generator().manual_seed(30))
    sout = swin(xin)
    np.savez_compressed(OUT / "swinir_small.npz", x=xin.numpy(), y=sout.numpy())
    print("swinir out", sout.mean().item(), sout.std().item())

    # ---------------- schedules / samplers with an analytic stand-in model ----------------
    out = {}
    for name, zero_snr, param in (("eps", False, "eps"), ("v", True, "v")):
        diff = Diffusion(linear_start=0.00085, linear_end=0.0120, timesteps=1000, zero_snr=zero_snr,
                         parameterization=param)
        out[f"betas_{name}"] = diff.betas
        sp = SpacedSampler(diff.betas, param, rescale_cfg=False)
        sp.make_schedule(50)
        out[f"spaced_ts_{name}"] = sp.timesteps
        for k in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                  "posterior_mean_coef1", "posterior_mean_coef2", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod"):
            out[f"spaced_{k}_{name}"] = getattr(sp, k).numpy()
        dd = DDIMSampler(diff.betas, param, rescale_cfg=False, eta=0)
        dd.make_schedule(50)
        out[f"ddim_ts_{name}"] = dd.ddim_timesteps
        for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas", "ddim_sqrt_one_minus_alphas", "ddim_sqrt_alphas"):
            out[f"{k}_{name}"] = getattr(dd, k).numpy()

        class Stub(torch.nn.Module):            # deterministic stand-in for ControlLDM
            def forward(self, x, t, cond):
                return (0.3 * torch.tanh(x) + 0.05 * cond["c_img"]
                        + 0.01 * cond["c_txt"].mean(dim=(1, 2)).view(-1, 1, 1, 1)
                        + 1e-4 * t.float().view(-1, 1, 1, 1))

        stub = Stub()
        xT = rnd(40, 1, 4, 24, 40)
        cond = dict(c_txt=rnd(41, 1, 77, 8), c_img=rnd(42, 1, 4, 24, 40))
        uncond = dict(c_txt=rnd(43, 1, 77, 8), c_img=cond["c_img"].clone())
        for sname, sampler in (("spaced", sp), ("ddim", dd)):
            for tiled in (False, True):
                torch.manual_seed(7)
                zz = sampler.sample(stub, "cpu", 10, (1, 4, 24, 40), cond, uncond, 4.0, tiled=tiled,
                                    tile_size=16, tile_stride=8, x_T=xT.clone(), progress=False)
                out[f"traj_{sname}_{name}_{'tiled' if tiled else 'full'}"] = zz.numpy()
    out["xT"] = xT.numpy()
    for k in ("c_txt", "c_img"):
        out[f"cond_{k}"] = cond[k].numpy()
        out[f"uncond_{k}"] = uncond[k].numpy()
    # per-step noises drawn by the samplers above (seed 7, 10 draws of xT's shape)
    torch.manual_seed(7)
    out["noises"] = torch.stack([torch.randn_like(xT) for _ in range(10)]).numpy()
    out["gauss_16"] = gaussian_weights(16, 16)
    out["windows_24_40_16_8"] = np.array(sliding_windows(24, 40, 16, 8))
    out["windows_30_30_16_12"] = np.array(sliding_windows(30, 30, 16, 12))
    a = torch.rand(1, 3, 40, 56, generator=torch.Generator().manual_seed(50))
    b = torch.rand(1, 3, 40, 56, generator=torch.Generator().manual_seed(51))
    out["wavelet_a"], out["wavelet_b"] = a.numpy(), b.numpy()
    out["wavelet_out"] = wavelet_reconstruction(a, b).numpy()
    out["resize_in"] = a.numpy()
    out["resize_out"] = resize_short_edge_to(a, 64).numpy()
    np.savez_compressed(OUT / "sampling.npz", **out)
    print("done")


if __name__ == "__main__":
    main()
