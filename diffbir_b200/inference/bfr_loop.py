"""BFRInferenceLoop (reference: diffbir/inference/bfr_loop.py): blind restoration of aligned faces —
the face SwinIR (same architecture as the general one) + the version's ControlNet."""
import numpy as np
from PIL import Image

from ..pipeline import SwinIRPipeline
from ..utils.common import instantiate_from_config
from .loop import InferenceLoop, load_checkpoint, load_config
from .pretrained_models import resolve


class BFRInferenceLoop(InferenceLoop):
    def load_cleaner(self) -> None:
        from ..model import SwinIR  # noqa: F401
        self.cleaner = instantiate_from_config(load_config("swinir.yaml"), device=self.args.device)
        if self.synthetic:
            from .. import arch
            from ..utils import synth
            self.cleaner.load_state_dict(synth.make_state_dict(arch.swinir_shapes(arch.SWINIR_CFG), self.args.seed + 4))
        else:
            self.cleaner.load_state_dict(load_checkpoint(resolve("swinir_face", self.weights_dir)), strict=True)

    def load_pipeline(self) -> None:
        self.pipeline = SwinIRPipeline(self.cleaner, self.cldm, self.diffusion, self.cond_fn, self.args.device)

    def after_load_lq(self, lq: Image.Image) -> np.ndarray:
        lq = lq.resize(tuple(int(x * self.args.upscale) for x in lq.size), Image.BICUBIC)
        return super().after_load_lq(lq)
