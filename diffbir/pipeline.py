"""diffbir.pipeline (reference pipeline.py:43-420) -> diffbir_b200.pipeline."""
from diffbir_b200.pipeline import BSRNetPipeline, Pipeline, SwinIRPipeline, pad_to_multiples_of, resize_short_edge_to  # noqa: F401

from ._unsupported import unsupported

SCUNetPipeline = unsupported("SCUNetPipeline", "pipeline.py:400-420", "use SwinIRPipeline")
