"""diffbir.model.cldm -> diffbir_b200.model.cldm."""
from diffbir_b200.model.cldm import ControlLDM  # noqa: F401
