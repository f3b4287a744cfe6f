"""(named to run last under `pytest -x`)
-m gpu: the inference.py command line end to end on synthetic weights (full SD-2.1 config):
folder of LQ images -> BSR / BID InferenceLoop -> SwinIR / BSRNet / SCUNet pipeline -> PNGs of upscale x the input size."""
import sys
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


# (random-init weights do not denoise: the EDM family starts from x_T * sigma_max = 1e4 under the eps parameterization
# and only a trained network brings that back to unit scale, so the EDM case uses the v-parameterized v2.1 recipe)
@pytest.mark.parametrize("task,version,sampler", [("sr", "v2.1", "edm_dpm++_3m_sde"),  # SwinIR cleaner; the reference CLI's default sampler
                                                  ("sr", "v2", "spaced"),               # BSRNet cleaner, BSRNetPipeline
                                                  ("denoise", "v2", "dpm++_m2")])       # SCUNet cleaner, SCUNetPipeline, DPM-Solver++
def test_cli_synthetic_end_to_end(tmp_path, task, version, sampler):
    import inference as cli
    from diffbir_b200.utils.synth import synthetic_lq
    (tmp_path / "in").mkdir()
    lq = synthetic_lq(128, 160, seed=3)[0]
    Image.fromarray(lq).save(tmp_path / "in" / "img.png")
    cli.main(["--task", task, "--version", version, "--upscale", "4", "--sampler", sampler, "--steps", "3",
              "--cfg_scale", "4.0", "--pos_prompt", "", "--neg_prompt", "low quality, blurry", "--captioner", "none",
              "--input", str(tmp_path / "in"), "--output", str(tmp_path / "out"), "--synthetic"])
    out = np.array(Image.open(tmp_path / "out" / "img.png"))
    assert out.shape == (512, 640, 3) and out.dtype == np.uint8
    assert out.std() > 1.0                                    # an image, not a constant
    assert (tmp_path / "out" / "prompt.csv").exists()


def test_cli_custom_version_from_a_training_yaml(tmp_path):
    """`--version custom --train_cfg <yaml> --ckpt <controlnet>` (CustomInferenceLoop, custom_loop.py:19-93): networks
    reflected from a training YAML in the reference's layout (reduced widths, reference-only keys included), weights
    read from checkpoint files in the layouts the reference's loaders unwrap (`state_dict` wrapper, `module.` prefix);
    the saved image must be what the same networks give when built directly."""
    import torch
    import yaml
    import inference as cli
    from diffbir_b200 import arch
    from diffbir_b200.utils.synth import RUN_DEFAULTS, build_synthetic_pipeline, make_state_dict, synthetic_lq, synthetic_sd_checkpoint
    from tests.small_cfg import CLIP_SMALL, CN_SMALL, SWIN_SMALL, UNET_SMALL, VAE_SMALL
    ref_only = dict(use_checkpoint=True, image_size=32, use_spatial_transformer=True, use_linear_in_transformer=True, legacy=False)
    plain = lambda d: {k: (list(v) if isinstance(v, tuple) else v) for k, v in d.items()}      # noqa: E731
    vae_yaml = dict(embed_dim=VAE_SMALL["embed_dim"],
                    ddconfig=dict(double_z=True, z_channels=VAE_SMALL["z_channels"], resolution=256, in_channels=VAE_SMALL["in_channels"],
                                  out_ch=VAE_SMALL["out_ch"], ch=VAE_SMALL["ch"], ch_mult=list(VAE_SMALL["ch_mult"]),
                                  num_res_blocks=VAE_SMALL["num_res_blocks"], attn_resolutions=[], dropout=0.0))
    clip_yaml = dict(embed_dim=CLIP_SMALL["embed_dim"], vision_cfg=dict(image_size=32, layers=1, width=64, head_width=32, patch_size=16),
                     text_cfg={k: CLIP_SMALL[k] for k in ("context_length", "vocab_size", "width", "heads", "layers")}, layer="penultimate")
    cfg = dict(model=dict(
        cldm=dict(target="diffbir.model.cldm.ControlLDM",
                  params=dict(latent_scale_factor=0.18215, unet_cfg=dict(plain(UNET_SMALL), **ref_only), vae_cfg=vae_yaml, clip_cfg=clip_yaml,
                              controlnet_cfg=dict(plain({k: v for k, v in CN_SMALL.items() if k != "out_channels"}), **ref_only))),
        swinir=dict(target="diffbir.model.swinir.SwinIR",
                    params=dict(img_size=SWIN_SMALL["img_size"], patch_size=1, in_chans=3, embed_dim=SWIN_SMALL["embed_dim"],
                                depths=list(SWIN_SMALL["depths"]), num_heads=list(SWIN_SMALL["num_heads"]), window_size=8,
                                mlp_ratio=SWIN_SMALL["mlp_ratio"], sf=8, img_range=1.0, upsampler="nearest+conv", resi_connection="1conv",
                                unshuffle=True, unshuffle_scale=8)),
        diffusion=dict(target="diffbir.model.gaussian_diffusion.Diffusion",
                       params=dict(linear_start=0.00085, linear_end=0.0120, timesteps=1000, zero_snr=False, parameterization="eps"))),
        train=dict(sd_path=str(tmp_path / "sd.ckpt"), swinir_path=str(tmp_path / "swinir.ckpt"), learning_rate=1e-4))
    (tmp_path / "train.yaml").write_text(yaml.safe_dump(cfg))
    torch.save({"state_dict": synthetic_sd_checkpoint(UNET_SMALL, VAE_SMALL, CLIP_SMALL, 1234), "global_step": 7}, tmp_path / "sd.ckpt")
    torch.save(make_state_dict(arch.unet_shapes(CN_SMALL, True), 1237, arch.is_zero_init), tmp_path / "controlnet.pt")
    torch.save({"state_dict": {"module." + k: v for k, v in make_state_dict(arch.swinir_shapes(SWIN_SMALL), 1238).items()}},
               tmp_path / "swinir.ckpt")
    (tmp_path / "in").mkdir()
    lq = synthetic_lq(96, 128, seed=5)[0]
    Image.fromarray(lq).save(tmp_path / "in" / "img.png")
    cli.main(["--version", "custom", "--train_cfg", str(tmp_path / "train.yaml"), "--ckpt", str(tmp_path / "controlnet.pt"),
              "--upscale", "4", "--sampler", "spaced", "--steps", "3", "--cfg_scale", "4.0", "--pos_prompt", "a photo",
              "--neg_prompt", RUN_DEFAULTS["neg_prompt"], "--captioner", "none", "--input", str(tmp_path / "in"),
              "--output", str(tmp_path / "out"), "--synthetic"])
    out = np.array(Image.open(tmp_path / "out" / "img.png"))
    assert out.shape == (384, 512, 3) and out.dtype == np.uint8 and out.std() > 1.0
    pipe = build_synthetic_pipeline("cuda", seed=1234, small=True)          # seeds 1234 / +3 / +4: the same three checkpoints
    up = np.array(Image.fromarray(lq).resize((512, 384), Image.BICUBIC))
    cli.set_seed(231)
    direct = pipe.run(up[None], **dict(RUN_DEFAULTS, steps=3, pos_prompt="a photo"))[0]
    mse = ((out.astype(np.float64) - direct.astype(np.float64)) ** 2).mean()
    psnr = float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    print(f"custom loop vs directly built pipeline: PSNR {psnr:.1f} dB")
    assert psnr > 50.0
