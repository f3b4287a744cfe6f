"""diffbir.inference.pretrained_models (reference inference/pretrained_models.py) -> the local checkpoint registry:
MODELS maps the reference's keys to the file names its URLs end in; `diffbir.utils.common.load_model_from_url(MODELS[k])`
reads that file from the weights directory instead of downloading it."""
from diffbir_b200.inference.pretrained_models import MODELS, default_weights_dir, resolve  # noqa: F401
