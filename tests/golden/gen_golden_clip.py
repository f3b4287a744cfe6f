"""Golden fixture for the OpenCLIP text tower (penultimate layer), produced by the REFERENCE
(`diffbir.model.clip.FrozenOpenCLIPEmbedder`, imported read-only from /root/reference) on the reduced
text config with seeded synthetic weights and token ids.

    python tests/golden/gen_golden_clip.py        ->  tests/golden/clip_small.npz
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_shims"))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from _ref_import import use_reference  # noqa: E402

use_reference()      # `import diffbir` = the reference checkout, not this repo's alias package

from diffbir_b200 import arch  # noqa: E402
from diffbir_b200.utils.synth import make_state_dict  # noqa: E402
from tests.small_cfg import CLIP_SMALL  # noqa: E402

OUT = Path(__file__).resolve().parent


@torch.no_grad()
def main():
    from diffbir.model.clip import FrozenOpenCLIPEmbedder
    text_cfg = {k: CLIP_SMALL[k] for k in ("context_length", "vocab_size", "width", "heads", "layers")}
    vision_cfg = dict(image_size=32, layers=1, width=64, head_width=32, patch_size=16)     # deleted by the embedder
    emb = FrozenOpenCLIPEmbedder(CLIP_SMALL["embed_dim"], vision_cfg, text_cfg, layer="penultimate").eval()
    sd = make_state_dict(arch.clip_text_shapes(CLIP_SMALL), 7)
    missing, unexpected = emb.model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k == "attn_mask" for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(70)
    tokens = torch.randint(0, CLIP_SMALL["vocab_size"], (3, CLIP_SMALL["context_length"]), generator=g)
    out = emb(tokens)
    np.savez_compressed(OUT / "clip_small.npz", tokens=tokens.numpy(), out=out.numpy())
    print(out.shape, float(out.abs().mean()))


if __name__ == "__main__":
    main()
