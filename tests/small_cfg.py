"""Reduced-size architectures used by the golden fixtures and the fast parity tests: same block
structure as configs/inference/*.yaml, channel counts scaled down to multiples of 64."""
from diffbir_b200 import arch

UNET_SMALL = dict(arch.UNET_CFG, model_channels=64, context_dim=128)
CN_SMALL = dict(UNET_SMALL, hint_channels=4)
VAE_SMALL = dict(arch.VAE_CFG, ch=64)
SWIN_SMALL = dict(arch.SWINIR_CFG, depths=(2, 2), num_heads=(6, 6))
CLIP_SMALL = dict(arch.CLIP_TEXT_CFG, width=128, heads=4, layers=3, vocab_size=512, embed_dim=128)
RRDB_SMALL = dict(arch.RRDBNET_CFG, nb=2)          # same widths (the engine needs nf, nf + 4 gc multiples of 64), 2 of 23 blocks
SCUNET_SMALL = dict(arch.SCUNET_CFG, config=(2, 2, 1, 2, 1, 2, 2))   # same widths, 12 of 28 blocks (both W and SW types per stage where 2)
