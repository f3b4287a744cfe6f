"""diffbir.inference (reference inference/__init__.py:1-5) -> diffbir_b200.inference."""
from diffbir_b200.inference import BFRInferenceLoop, BIDInferenceLoop, BSRInferenceLoop, InferenceLoop  # noqa: F401

from .._unsupported import unsupported

UnAlignedBFRInferenceLoop = unsupported("UnAlignedBFRInferenceLoop", "inference/unaligned_bfr_loop.py:14-152",
                                        "face detection / alignment is outside the path; use BFRInferenceLoop on aligned faces")
CustomInferenceLoop = unsupported("CustomInferenceLoop", "inference/custom_loop.py:12-64",
                                  "construct diffbir.model.{ControlLDM, SwinIR, Diffusion} and diffbir.pipeline.SwinIRPipeline directly")
