"""BSRInferenceLoop (reference: diffbir/inference/bsr_loop.py): blind super-resolution. v1 / v2.1 use a SwinIR
stage-1 cleaner on the pre-upscaled LQ image (v1: general SwinIR, v2.1: the Real-ESRGAN-degradation SwinIR);
v2 pairs the ControlNet with BSRNet (RRDBNet x4 on the LQ image itself, BSRNetPipeline)."""
import numpy as np
from PIL import Image

from ..pipeline import BSRNetPipeline, SwinIRPipeline
from ..utils.common import instantiate_from_config
from .loop import InferenceLoop, load_checkpoint, load_config
from .pretrained_models import resolve


class BSRInferenceLoop(InferenceLoop):
    def _swinir(self) -> bool:
        return self.args.version in ("v1", "v2.1")

    def load_cleaner(self) -> None:
        from .. import arch
        from ..utils import synth
        if self._swinir():
            self.cleaner = instantiate_from_config(load_config("swinir.yaml"), device=self.args.device)
            shapes, key = arch.swinir_shapes(arch.SWINIR_CFG), ("swinir_general" if self.args.version == "v1" else "swinir_realesrgan")
        else:
            self.cleaner = instantiate_from_config(load_config("bsrnet.yaml"), device=self.args.device)
            shapes, key = arch.rrdbnet_shapes(arch.RRDBNET_CFG), "bsrnet"
        if self.synthetic:
            self.cleaner.load_state_dict(synth.make_state_dict(shapes, self.args.seed + 4))
        else:
            self.cleaner.load_state_dict(load_checkpoint(resolve(key, self.weights_dir)), strict=True)

    def load_pipeline(self) -> None:
        if self._swinir():
            self.pipeline = SwinIRPipeline(self.cleaner, self.cldm, self.diffusion, self.cond_fn, self.args.device)
        else:
            self.pipeline = BSRNetPipeline(self.cleaner, self.cldm, self.diffusion, self.cond_fn, self.args.device,
                                           self.args.upscale)

    def after_load_lq(self, lq: Image.Image) -> np.ndarray:
        if self._swinir():       # the LQ image is brought to the output size before stage 1 (bsr_loop.py:55-60)
            lq = lq.resize(tuple(int(x * self.args.upscale) for x in lq.size), Image.BICUBIC)
        return super().after_load_lq(lq)
