"""diffbir.inference.loop -> diffbir_b200.inference.loop."""
from diffbir_b200.inference.loop import *  # noqa: F401,F403
from diffbir_b200.inference.loop import InferenceLoop  # noqa: F401
