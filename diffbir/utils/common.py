"""diffbir.utils.common (reference utils/common.py) -> diffbir_b200.utils.common."""
from diffbir_b200.utils.common import *  # noqa: F401,F403
from diffbir_b200.utils.common import (VRAMPeakMonitor, gaussian_weights, instantiate_from_config, load_file_from_url,  # noqa: F401
                                       load_model_from_url, make_tiled_fn, sliding_windows, to, wavelet_decomposition,
                                       wavelet_reconstruction)
