"""diffbir.model.bsrnet -> diffbir_b200.model.bsrnet."""
from diffbir_b200.model.bsrnet import RRDBNet  # noqa: F401
