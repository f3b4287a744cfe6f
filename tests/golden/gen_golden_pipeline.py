"""Fixture of a WHOLE reference run: diffbir.pipeline.SwinIRPipeline.run (uint8 in -> uint8 out) on the CPU with the
reduced networks and seeded synthetic weights loaded through the reference's own loaders:

    python tests/golden/gen_golden_pipeline.py        # needs /root/reference; writes pipeline_small.npz

SwinIR stage 1 (resize to short edge 512), VAE encode, OpenCLIP text tower, spaced / DDIM sampler with CFG, VAE
decode, wavelet colour fix, antialiased resize back, uint8 truncation. tests/test_oracle_golden.py replays it through
the oracle's restatement of the pipeline. The reduced text tower has a 512-entry vocabulary, so the BPE tokenizer is
replaced by the product's SyntheticTokenizer on BOTH sides (tokenisation is not on the numeric path being pinned).
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from _ref_import import use_reference  # noqa: E402

use_reference()
from diffbir_b200 import arch  # noqa: E402
from diffbir_b200.model.clip import SyntheticTokenizer  # noqa: E402
from diffbir_b200.utils.synth import make_state_dict, synthetic_lq, synthetic_sd_checkpoint  # noqa: E402
from tests.small_cfg import CLIP_SMALL, CN_SMALL, SWIN_SMALL, UNET_SMALL, VAE_SMALL  # noqa: E402

OUT = Path(__file__).resolve().parent
NEG = "low quality, blurry, low-resolution, noisy, unsharp, weird textures"
RUN = dict(strength=1.0, cleaner_tiled=False, cleaner_tile_size=512, cleaner_tile_stride=256, vae_encoder_tiled=False,
           vae_encoder_tile_size=256, vae_decoder_tiled=False, vae_decoder_tile_size=256, cldm_tiled=False, cldm_tile_size=512,
           cldm_tile_stride=256, pos_prompt="a photo", neg_prompt=NEG, cfg_scale=4.0, start_point_type="noise", noise_aug=0,
           rescale_cfg=False, s_churn=0, s_tmin=0, s_tmax=300, s_noise=1, eta=1, order=1)


@torch.no_grad()
def main():
    import contextlib
    import diffbir.model.clip as rclip
    import diffbir.pipeline as rpipe
    from diffbir.model.cldm import ControlLDM
    from diffbir.model.gaussian_diffusion import Diffusion
    from diffbir.model.swinir import SwinIR
    rclip.tokenize = SyntheticTokenizer(CLIP_SMALL["vocab_size"])
    rpipe.VRAMPeakMonitor = lambda *a, **k: contextlib.nullcontext()     # calls torch.cuda.synchronize() (SURVEY 8c patch 2)

    ref_kw = dict(use_checkpoint=False, image_size=32, use_spatial_transformer=True, use_linear_in_transformer=True, legacy=False)
    ucfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in UNET_SMALL.items()}
    ccfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in CN_SMALL.items() if k != "out_channels"}
    vae_yaml = dict(embed_dim=VAE_SMALL["embed_dim"],
                    ddconfig=dict(double_z=True, z_channels=VAE_SMALL["z_channels"], resolution=256, in_channels=VAE_SMALL["in_channels"],
                                  out_ch=VAE_SMALL["out_ch"], ch=VAE_SMALL["ch"], ch_mult=list(VAE_SMALL["ch_mult"]),
                                  num_res_blocks=VAE_SMALL["num_res_blocks"], attn_resolutions=[], dropout=0.0))
    clip_yaml = dict(embed_dim=CLIP_SMALL["embed_dim"], vision_cfg=dict(image_size=32, layers=1, width=64, head_width=32, patch_size=16),
                     text_cfg={k: CLIP_SMALL[k] for k in ("context_length", "vocab_size", "width", "heads", "layers")}, layer="penultimate")
    cldm = ControlLDM(dict(ucfg, **ref_kw), vae_yaml, clip_yaml, dict(ccfg, **ref_kw), 0.18215).eval()
    unused, missing = cldm.load_pretrained_sd(synthetic_sd_checkpoint(UNET_SMALL, VAE_SMALL, CLIP_SMALL, 1234))
    assert all("attn_mask" in k or "logit_scale" in k or "text_projection" in k for k in missing), sorted(missing)[:5]
    cldm.load_controlnet_from_ckpt(make_state_dict(arch.unet_shapes(CN_SMALL, True), 1237, arch.is_zero_init))
    scfg = SWIN_SMALL
    swin = SwinIR(img_size=scfg["img_size"], patch_size=1, in_chans=3, embed_dim=scfg["embed_dim"], depths=list(scfg["depths"]),
                  num_heads=list(scfg["num_heads"]), window_size=8, mlp_ratio=scfg["mlp_ratio"], sf=8, img_range=1.0,
                  upsampler="nearest+conv", resi_connection="1conv", unshuffle=True, unshuffle_scale=8).eval()
    ssd = make_state_dict(arch.swinir_shapes(scfg), 1238)
    missing_s, unexpected_s = swin.load_state_dict(ssd, strict=False)
    assert not unexpected_s and all(k.endswith(("relative_position_index", "attn_mask")) for k in missing_s)
    out = dict()
    lq = synthetic_lq(96, 128, seed=5)
    out["lq"] = lq
    for sampler, steps, pname, zs in (("spaced", 3, "eps", False), ("ddim", 4, "v", True)):   # DDIM: 1000 % steps == 0 (the reference indexes alphas_cumprod[1000] otherwise)
        diffusion = Diffusion(linear_start=0.00085, linear_end=0.0120, timesteps=1000, parameterization=pname, zero_snr=zs)
        pipe = rpipe.SwinIRPipeline(swin, cldm, diffusion, None, "cpu")
        torch.manual_seed(231)
        y = pipe.run(lq, steps, RUN["strength"], RUN["cleaner_tiled"], RUN["cleaner_tile_size"], RUN["cleaner_tile_stride"],
                     RUN["vae_encoder_tiled"], RUN["vae_encoder_tile_size"], RUN["vae_decoder_tiled"], RUN["vae_decoder_tile_size"],
                     RUN["cldm_tiled"], RUN["cldm_tile_size"], RUN["cldm_tile_stride"], RUN["pos_prompt"], RUN["neg_prompt"],
                     RUN["cfg_scale"], RUN["start_point_type"], sampler, RUN["noise_aug"], RUN["rescale_cfg"], RUN["s_churn"],
                     RUN["s_tmin"], RUN["s_tmax"], RUN["s_noise"], RUN["eta"], RUN["order"])
        out[f"out_{sampler}_{pname}"] = y
        print(sampler, pname, y.shape, y.dtype, float(y.mean()), float(y.std()))
    # The other two stage-1 pipelines with the two other sampler families (both step rules are noise-free, so a GPU run
    # with the same x_T is comparable): BSRNetPipeline (x4 RRDBNet, pipeline.py:324-366) + EDM dpm++_2m on the v / zero-SNR
    # model, SCUNetPipeline (pipeline.py:400-421) + DPM-Solver++ multistep order 2 on the eps model.
    from diffbir.model.bsrnet import RRDBNet
    from diffbir.model.scunet import SCUNet
    from tests.small_cfg import RRDB_SMALL, SCUNET_SMALL
    rr = RRDBNet(**RRDB_SMALL).eval()
    rr.load_state_dict(make_state_dict(arch.rrdbnet_shapes(RRDB_SMALL), 91), strict=True)
    sc = SCUNet(in_nc=SCUNET_SMALL["in_nc"], config=list(SCUNET_SMALL["config"]), dim=SCUNET_SMALL["dim"]).eval()
    sc.load_state_dict(make_state_dict(arch.scunet_shapes(SCUNET_SMALL), 9), strict=True)
    lq2 = synthetic_lq(128, 160, seed=3)
    out["lq_bsr"] = lq2
    for tag, sampler, steps, pname, zs in (("bsrnet", "edm_dpm++_2m", 6, "v", True), ("scunet", "dpm++_m2", 6, "eps", False)):
        diffusion = Diffusion(linear_start=0.00085, linear_end=0.0120, timesteps=1000, parameterization=pname, zero_snr=zs)
        if tag == "bsrnet":
            pipe, img = rpipe.BSRNetPipeline(rr, cldm, diffusion, None, "cpu", 4.0), lq2
        else:
            pipe, img = rpipe.SCUNetPipeline(sc, cldm, diffusion, None, "cpu"), lq
        torch.manual_seed(231)
        y = pipe.run(img, steps, RUN["strength"], RUN["cleaner_tiled"], RUN["cleaner_tile_size"], RUN["cleaner_tile_stride"],
                     RUN["vae_encoder_tiled"], RUN["vae_encoder_tile_size"], RUN["vae_decoder_tiled"], RUN["vae_decoder_tile_size"],
                     RUN["cldm_tiled"], RUN["cldm_tile_size"], RUN["cldm_tile_stride"], RUN["pos_prompt"], RUN["neg_prompt"],
                     RUN["cfg_scale"], RUN["start_point_type"], sampler, RUN["noise_aug"], RUN["rescale_cfg"], RUN["s_churn"],
                     RUN["s_tmin"], RUN["s_tmax"], RUN["s_noise"], RUN["eta"], RUN["order"])
        out[f"out_{tag}"] = y
        print(tag, sampler, pname, y.shape, y.dtype, float(y.mean()), float(y.std()))
    # The optional branches of Pipeline.apply_cldm / apply_cleaner: start point "cond" + noise augmentation of the
    # condition (pipeline.py:146-167), control strength 0.7 (:173-174), the cosine CFG ramp; and both tiled modes
    # (Gaussian-blended SwinIR tiles, mixture-of-diffusers latent tiles).
    for tag, sampler, steps, pname, zs, over in (
            ("opts", "spaced", 3, "eps", False, dict(start_point_type="cond", noise_aug=40, rescale_cfg=True, strength=0.7)),
            # Tiled-VAE flags with tiles larger than the image: the reference runs the VAE un-tiled but pads to 8, not 64 (:100-110)
            ("vaetiny", "spaced", 3, "eps", False, dict(vae_encoder_tiled=True, vae_encoder_tile_size=1024, vae_decoder_tiled=True,
                                                         vae_decoder_tile_size=1024)),
            ("tiled", "ddim", 4, "v", True, dict(cleaner_tiled=True, cleaner_tile_size=64, cleaner_tile_stride=32, cldm_tiled=True,
                                                  cldm_tile_size=512, cldm_tile_stride=256))):
        r = dict(RUN, **over)
        diffusion = Diffusion(linear_start=0.00085, linear_end=0.0120, timesteps=1000, parameterization=pname, zero_snr=zs)
        pipe = rpipe.SwinIRPipeline(swin, cldm, diffusion, None, "cpu")
        torch.manual_seed(231)
        y = pipe.run(lq, steps, r["strength"], r["cleaner_tiled"], r["cleaner_tile_size"], r["cleaner_tile_stride"],
                     r["vae_encoder_tiled"], r["vae_encoder_tile_size"], r["vae_decoder_tiled"], r["vae_decoder_tile_size"],
                     r["cldm_tiled"], r["cldm_tile_size"], r["cldm_tile_stride"], r["pos_prompt"], r["neg_prompt"],
                     r["cfg_scale"], r["start_point_type"], sampler, r["noise_aug"], r["rescale_cfg"], r["s_churn"],
                     r["s_tmin"], r["s_tmax"], r["s_noise"], r["eta"], r["order"])
        out[f"out_{tag}"] = y
        print(tag, sampler, pname, y.shape, y.dtype, float(y.mean()), float(y.std()))
    np.savez_compressed(OUT / "pipeline_small.npz", **out)
    print("wrote pipeline_small.npz")


if __name__ == "__main__":
    main()
