"""diffbir.utils.common (reference utils/common.py) -> diffbir_b200.utils.common."""
from diffbir_b200.utils.common import *  # noqa: F401,F403
from diffbir_b200.utils.common import (gaussian_weights, instantiate_from_config, make_tiled_fn, sliding_windows,  # noqa: F401
                                       wavelet_reconstruction)
