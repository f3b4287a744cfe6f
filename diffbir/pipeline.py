"""diffbir.pipeline (reference pipeline.py:43-420) -> diffbir_b200.pipeline."""
from diffbir_b200.pipeline import Pipeline, SwinIRPipeline, pad_to_multiples_of, resize_short_edge_to  # noqa: F401

from ._unsupported import unsupported

BSRNetPipeline = unsupported("BSRNetPipeline", "pipeline.py:324-366",
                             "use SwinIRPipeline (v1 / v2.1 BSR stage 1), the north-star cleaner")
SCUNetPipeline = unsupported("SCUNetPipeline", "pipeline.py:400-420", "use SwinIRPipeline")
