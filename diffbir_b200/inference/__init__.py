"""diffbir.inference counterpart: the inference loops whose networks are on the accelerated path.
Unaligned faces (face detector) and custom loops are refused by the CLI."""
from .bfr_loop import BFRInferenceLoop
from .bid_loop import BIDInferenceLoop
from .bsr_loop import BSRInferenceLoop
from .loop import InferenceLoop

__all__ = ["InferenceLoop", "BSRInferenceLoop", "BFRInferenceLoop", "BIDInferenceLoop"]
