"""diffbir.model.swinir -> diffbir_b200.model.swinir."""
from diffbir_b200.model.swinir import SwinIR  # noqa: F401
