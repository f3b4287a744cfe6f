"""diffbir.inference counterpart: the inference loops whose networks are on the accelerated path.
The unaligned-face loop (face detector) is refused by the CLI."""
from .bfr_loop import BFRInferenceLoop
from .bid_loop import BIDInferenceLoop
from .bsr_loop import BSRInferenceLoop
from .custom_loop import CustomInferenceLoop
from .loop import InferenceLoop

__all__ = ["InferenceLoop", "BSRInferenceLoop", "BFRInferenceLoop", "BIDInferenceLoop", "CustomInferenceLoop"]
