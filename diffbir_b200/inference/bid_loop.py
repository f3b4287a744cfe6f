"""BIDInferenceLoop (reference: diffbir/inference/bid_loop.py): blind image denoising. v1 / v2.1 use a SwinIR
stage-1 cleaner, v2 pairs the ControlNet with SCUNet (SCUNetPipeline); the LQ image is always brought to the
output size before stage 1 (bid_loop.py:51-55)."""
import numpy as np
from PIL import Image

from ..pipeline import SCUNetPipeline, SwinIRPipeline
from ..utils.common import instantiate_from_config
from .loop import InferenceLoop, load_checkpoint, load_config
from .pretrained_models import resolve


class BIDInferenceLoop(InferenceLoop):
    def _swinir(self) -> bool:
        return self.args.version in ("v1", "v2.1")

    def load_cleaner(self) -> None:
        from .. import arch
        from ..utils import synth
        if self._swinir():
            self.cleaner = instantiate_from_config(load_config("swinir.yaml"), device=self.args.device)
            shapes, key = arch.swinir_shapes(arch.SWINIR_CFG), ("swinir_general" if self.args.version == "v1" else "swinir_realesrgan")
        else:
            self.cleaner = instantiate_from_config(load_config("scunet.yaml"), device=self.args.device)
            shapes, key = arch.scunet_shapes(arch.SCUNET_CFG), "scunet_psnr"
        if self.synthetic:
            self.cleaner.load_state_dict(synth.make_state_dict(shapes, self.args.seed + 4))
        else:
            self.cleaner.load_state_dict(load_checkpoint(resolve(key, self.weights_dir)), strict=True)

    def load_pipeline(self) -> None:
        cls = SwinIRPipeline if self._swinir() else SCUNetPipeline
        self.pipeline = cls(self.cleaner, self.cldm, self.diffusion, self.cond_fn, self.args.device)

    def after_load_lq(self, lq: Image.Image) -> np.ndarray:
        lq = lq.resize(tuple(int(x * self.args.upscale) for x in lq.size), Image.BICUBIC)
        return super().after_load_lq(lq)
