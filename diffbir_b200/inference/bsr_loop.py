"""BSRInferenceLoop (reference: diffbir/inference/bsr_loop.py): blind super-resolution with the
SwinIR stage-1 cleaner (v1: general SwinIR, v2.1: the Real-ESRGAN-degradation SwinIR). v2 pairs the
ControlNet with BSRNet, a cleaner outside the accelerated path -> refused."""
import numpy as np
from PIL import Image

from ..pipeline import SwinIRPipeline
from ..utils.common import instantiate_from_config
from .loop import InferenceLoop, load_checkpoint, load_config
from .pretrained_models import resolve


class BSRInferenceLoop(InferenceLoop):
    def load_cleaner(self) -> None:
        from ..model import SwinIR  # noqa: F401
        if self.args.version == "v2":
            raise NotImplementedError("--version v2 --task sr uses the BSRNet cleaner (SURVEY.md §8f); use v1 or v2.1")
        self.cleaner = instantiate_from_config(load_config("swinir.yaml"), device=self.args.device)
        if self.synthetic:
            from .. import arch
            from ..utils import synth
            self.cleaner.load_state_dict(synth.make_state_dict(arch.swinir_shapes(arch.SWINIR_CFG), self.args.seed + 4))
        else:
            key = "swinir_general" if self.args.version == "v1" else "swinir_realesrgan"
            self.cleaner.load_state_dict(load_checkpoint(resolve(key, self.weights_dir)), strict=True)

    def load_pipeline(self) -> None:
        self.pipeline = SwinIRPipeline(self.cleaner, self.cldm, self.diffusion, self.cond_fn, self.args.device)

    def after_load_lq(self, lq: Image.Image) -> np.ndarray:
        # v1 / v2.1: the LQ image is brought to the output size before stage 1 (bsr_loop.py:55-60)
        lq = lq.resize(tuple(int(x * self.args.upscale) for x in lq.size), Image.BICUBIC)
        return super().after_load_lq(lq)
