"""diffbir.model.scunet -> diffbir_b200.model.scunet."""
from diffbir_b200.model.scunet import SCUNet  # noqa: F401
