"""diffbir.pipeline (reference pipeline.py:43-420) -> diffbir_b200.pipeline."""
from diffbir_b200.pipeline import (BSRNetPipeline, Pipeline, SCUNetPipeline, SwinIRPipeline,  # noqa: F401
                                   pad_to_multiples_of, resize_short_edge_to)

