"""CPU: the inference CLI / loops (drop-in for the reference's inference.py + diffbir.inference):
flag surface, refusal of everything outside the accelerated path, file iteration, batching of
n_samples over batch_size, the 26 positional arguments handed to pipeline.run, saved files."""
import csv
import sys
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import inference as cli  # noqa: E402
from diffbir_b200.inference import loop as loop_mod  # noqa: E402
from diffbir_b200.inference.bsr_loop import BSRInferenceLoop  # noqa: E402
from diffbir_b200.inference.pretrained_models import MODELS, resolve  # noqa: E402

# every flag of the reference's inference.py:55-287
REFERENCE_FLAGS = """task upscale version train_cfg ckpt sampler steps start_point_type cleaner_tiled cleaner_tile_size
cleaner_tile_stride vae_encoder_tiled vae_encoder_tile_size vae_decoder_tiled vae_decoder_tile_size cldm_tiled
cldm_tile_size cldm_tile_stride captioner pos_prompt neg_prompt cfg_scale rescale_cfg noise_aug s_churn s_tmin s_tmax
s_noise eta order strength batch_size guidance g_loss g_scale input n_samples output seed device precision
llava_bit""".split()


def args_for(tmp_path, *extra):
    return cli.parse_args(["--input", str(tmp_path / "in"), "--output", str(tmp_path / "out"), *extra])


def test_cli_has_every_reference_flag_with_its_default(tmp_path):
    a = vars(args_for(tmp_path))
    assert not [f for f in REFERENCE_FLAGS if f not in a]
    assert (a["task"], a["upscale"], a["version"], a["steps"], a["cfg_scale"], a["seed"], a["precision"]) == \
        ("sr", 4, "v2.1", 10, 6.0, 231, "fp16")
    assert (a["cldm_tile_size"], a["cldm_tile_stride"], a["s_tmax"], a["eta"], a["order"], a["strength"]) == (512, 256, 300, 1, 1, 1)
    assert a["pos_prompt"].startswith("Cinematic") and a["neg_prompt"].startswith("painting")


@pytest.mark.parametrize("extra,msg", [(["--captioner", "llava"], "captioner"),
                                        (["--guidance"], "guidance"), (["--precision", "fp32"], "fp32"),
                                        (["--device", "cpu"], "CUDA"), (["--vae_decoder_tiled"], "Tiled-VAE")])
def test_options_outside_the_path_are_refused(tmp_path, extra, msg):
    with pytest.raises(NotImplementedError, match=msg):
        loop_mod.check_supported(args_for(tmp_path, *extra))
    loop_mod.check_supported(args_for(tmp_path))          # the defaults are runnable
    # VAE tiles larger than any 512-class input: the reference runs un-tiled then (pipeline.py:106-110), so does the pipeline
    loop_mod.check_supported(args_for(tmp_path, "--vae_decoder_tiled", "--vae_decoder_tile_size", "4096"))


def test_checkpoints_resolve_locally_and_fail_loudly(tmp_path):
    assert set(MODELS) >= {"sd_v2.1", "sd_v2.1_zsnr", "v1_general", "v1_face", "v2", "v2.1", "swinir_general",
                           "swinir_face", "swinir_realesrgan"}
    with pytest.raises(FileNotFoundError, match="DiffBIR_v2.1.pt"):
        resolve("v2.1", str(tmp_path))
    (tmp_path / "DiffBIR_v2.1.pt").write_bytes(b"x")
    assert resolve("v2.1", str(tmp_path)).endswith("DiffBIR_v2.1.pt")
    with pytest.raises(FileNotFoundError, match="scunet_color_real_psnr.pth"):
        resolve("scunet_psnr", str(tmp_path))
    with pytest.raises(FileNotFoundError, match="BSRNet.pth"):
        resolve("bsrnet", str(tmp_path))


class StubPipeline:
    def __init__(self):
        self.calls = []

    def run(self, *a):
        self.calls.append(a)
        lq = a[0]
        return np.stack([np.full(lq.shape[1:], 10 * len(self.calls) + i, dtype=np.uint8) for i in range(lq.shape[0])])


def make_loop(args):
    lp = BSRInferenceLoop.__new__(BSRInferenceLoop)       # skip model loading (needs the GPU library)
    lp.args, lp.loop_ctx, lp.pipeline = args, {}, StubPipeline()
    return lp


def test_run_iterates_batches_and_saves(tmp_path):
    (tmp_path / "in").mkdir()
    Image.fromarray(np.zeros((6, 8, 3), np.uint8)).save(tmp_path / "in" / "b.png")
    Image.fromarray(np.zeros((5, 7, 3), np.uint8)).save(tmp_path / "in" / "a.jpg")
    (tmp_path / "in" / "notes.txt").write_text("skip me")
    args = args_for(tmp_path, "--n_samples", "3", "--batch_size", "2", "--upscale", "2", "--steps", "7", "--cfg_scale", "4.5",
                    "--sampler", "ddim", "--pos_prompt", "p", "--neg_prompt", "n", "--cldm_tiled")
    lp = make_loop(args)
    lp.run()
    calls = lp.pipeline.calls
    assert len(calls) == 4                                   # 2 images x ceil(3 / 2) batches
    assert [c[0].shape for c in calls] == [(2, 10, 14, 3), (1, 10, 14, 3), (2, 12, 16, 3), (1, 12, 16, 3)]   # sorted, x2 bicubic
    assert all(len(c) == 26 for c in calls)                  # pipeline.run's positional contract (pipeline.py:236-264)
    c = calls[0]
    assert c[1:4] == (7, 1, False) and c[10:13] == (True, 512, 256) and c[13:16] == ("p", "n", 4.5)
    assert c[16:20] == ("noise", "ddim", 0, False) and c[20:] == (0, 0, 300, 1, 1, 1)
    out = sorted(p.name for p in (tmp_path / "out").iterdir())
    assert out == ["a_0.png", "a_1.png", "a_2.png", "b_0.png", "b_1.png", "b_2.png", "prompt.csv"]
    assert np.array(Image.open(tmp_path / "out" / "a_2.png")).shape == (10, 14, 3)
    rows = list(csv.reader(open(tmp_path / "out" / "prompt.csv")))
    assert rows == [["file_name", "pos_prompt", "neg_prompt"], ["a", "p", "n"], ["b", "p", "n"]]


def test_single_sample_file_name_and_configs(tmp_path):
    (tmp_path / "in").mkdir()
    Image.fromarray(np.zeros((4, 4, 3), np.uint8)).save(tmp_path / "in" / "x.jpeg")
    lp = make_loop(args_for(tmp_path, "--upscale", "1"))
    lp.run()
    assert sorted(p.name for p in (tmp_path / "out").iterdir()) == ["prompt.csv", "x.png"]
    # the YAML files instantiate through the reference's `target:` names
    for name, target in (("cldm.yaml", "diffbir.model.ControlLDM"), ("swinir.yaml", "diffbir.model.SwinIR"),
                         ("diffusion.yaml", "diffbir.model.Diffusion"), ("diffusion_v2.1.yaml", "diffbir.model.Diffusion")):
        cfg = loop_mod.load_config(name)
        assert cfg["target"] == target
    from diffbir_b200.utils.common import instantiate_from_config
    d = instantiate_from_config(loop_mod.load_config("diffusion_v2.1.yaml"))
    assert d.parameterization == "v" and d.zero_snr and d.sqrt_alphas_cumprod[-1] == 0
    from diffbir_b200 import arch
    cfg = loop_mod.load_config("cldm.yaml")["params"]
    assert {k: tuple(v) if isinstance(v, list) else v for k, v in cfg["unet_cfg"].items()} == arch.UNET_CFG
    want = {k: v for k, v in arch.CONTROLNET_CFG.items() if k != "out_channels"}      # a ControlNet has no output conv
    assert {k: tuple(v) if isinstance(v, list) else v for k, v in cfg["controlnet_cfg"].items()} == want
    assert list(arch.unet_shapes(dict(cfg["controlnet_cfg"]), True)) == list(arch.unet_shapes(arch.CONTROLNET_CFG, True))
    sw = loop_mod.load_config("swinir.yaml")["params"]
    assert sw["embed_dim"] == arch.SWINIR_CFG["embed_dim"] and tuple(sw["depths"]) == arch.SWINIR_CFG["depths"]


def test_load_checkpoint_with_foreign_pickled_globals(tmp_path):
    """Lightning-style checkpoints carry non-tensor globals that weights_only=True rejects;
    the loader falls back to an unpickler that keeps tensors and stubs everything else."""
    import sys
    import types

    import torch

    from diffbir_b200.inference.loop import load_checkpoint
    mod = types.ModuleType("fake_lightning_callbacks")

    class ModelCheckpoint:
        def __init__(self):
            self.best = 0.5

    ModelCheckpoint.__module__, ModelCheckpoint.__qualname__ = "fake_lightning_callbacks", "ModelCheckpoint"
    mod.ModelCheckpoint = ModelCheckpoint
    sys.modules["fake_lightning_callbacks"] = mod
    path = tmp_path / "lightning.ckpt"
    try:
        torch.save({"state_dict": {"module.w": torch.arange(6.0).view(2, 3)}, "callbacks": {"ckpt": ModelCheckpoint()},
                    "epoch": 3}, path)
    finally:
        del sys.modules["fake_lightning_callbacks"]
    sd = load_checkpoint(str(path))
    assert list(sd) == ["w"] and torch.equal(sd["w"], torch.arange(6.0).view(2, 3))


def test_custom_loop_reads_the_training_yaml(tmp_path, monkeypatch):
    """CustomInferenceLoop (custom_loop.py:19-93): networks from the training YAML's model block, weights from its train
    block + --ckpt, SwinIRPipeline, bicubic pre-upscale. Model construction needs the GPU library, so the reflection and
    the checkpoint reader are recorded instead."""
    import yaml
    from diffbir_b200.inference import custom_loop as cl
    with pytest.raises(ValueError, match="train_cfg"):
        cl.CustomInferenceLoop(args_for(tmp_path, "--version", "custom"))
    cfg = dict(model=dict(cldm=dict(target="diffbir.model.cldm.ControlLDM", params=dict(latent_scale_factor=0.18215)),
                          swinir=dict(target="diffbir.model.swinir.SwinIR", params=dict(img_size=64)),
                          diffusion=dict(target="diffbir.model.gaussian_diffusion.Diffusion", params=dict(parameterization="v"))),
               train=dict(sd_path=str(tmp_path / "sd.ckpt"), swinir_path=str(tmp_path / "swinir.ckpt")))
    (tmp_path / "train.yaml").write_text(yaml.safe_dump(cfg))
    events = []

    class Net:
        def __init__(self, kind):
            self.kind = kind

        def load_pretrained_sd(self, sd):
            events.append(("sd", sd))
            return [], []

        def load_controlnet_from_ckpt(self, sd):
            events.append(("controlnet", sd))

        def load_state_dict(self, sd, strict=True):
            events.append(("swinir", sd, strict))

    def fake_instantiate(config, **extra):
        events.append(("new", config["target"], dict(config.get("params", {})), extra))
        return Net(config["target"])

    monkeypatch.setattr(cl, "instantiate_from_config", fake_instantiate)
    monkeypatch.setattr(cl, "load_checkpoint", lambda path: f"<{Path(path).name}>")
    monkeypatch.setattr(cl.CustomInferenceLoop, "_operand_check", lambda self: None)
    args = args_for(tmp_path, "--version", "custom", "--train_cfg", str(tmp_path / "train.yaml"), "--ckpt",
                    str(tmp_path / "controlnet.pt"), "--upscale", "2")
    lp = cl.CustomInferenceLoop(args)
    assert [e[:2] for e in events] == [("new", "diffbir.model.swinir.SwinIR"), ("swinir", "<swinir.ckpt>"),
                                       ("new", "diffbir.model.cldm.ControlLDM"), ("sd", "<sd.ckpt>"),
                                       ("controlnet", "<controlnet.pt>"), ("new", "diffbir.model.gaussian_diffusion.Diffusion")]
    assert events[1][2] is True and events[0][3] == {"device": "cuda"} and events[2][3]["synthetic_tokenizer"] is False
    assert type(lp.pipeline).__name__ == "SwinIRPipeline" and lp.pipeline.cleaner.kind.endswith("SwinIR")
    assert lp.after_load_lq(Image.new("RGB", (30, 20))).shape == (40, 60, 3)
    # an empty train.sd_path (the shipped training YAML leaves it blank) is reported, not passed to torch.load
    cfg["train"]["sd_path"] = None
    (tmp_path / "train.yaml").write_text(yaml.safe_dump(cfg))
    with pytest.raises(ValueError, match="train.sd_path"):
        cl.CustomInferenceLoop(args)


def test_load_model_from_url_reads_the_local_file(tmp_path, monkeypatch):
    """utils/common.py:113-120 semantics on a local weights directory: `state_dict` wrapper and `module.` prefix removed."""
    import torch
    from diffbir_b200.utils.common import load_file_from_url, load_model_from_url
    monkeypatch.setenv("DIFFBIR_WEIGHTS_DIR", str(tmp_path))
    torch.save({"state_dict": {"module.a.weight": torch.arange(3.0), "module.b": torch.ones(1)}}, tmp_path / "v2.pth")
    sd = load_model_from_url("https://huggingface.co/org/repo/resolve/main/" + MODELS["v2"])
    assert list(sd) == ["a.weight", "b"] and torch.equal(sd["a.weight"], torch.arange(3.0))
    assert load_file_from_url("whatever/x.bin", model_dir=str(tmp_path), file_name="v2.pth").endswith("v2.pth")
    with pytest.raises(FileNotFoundError, match="downloads are not performed"):
        load_file_from_url("https://host/none.ckpt")
