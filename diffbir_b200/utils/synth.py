"""Seeded synthetic checkpoints in the reference's state_dict layout.

There is no network for real weights, so bench / tests use random-init weights of the real
architecture.  Every tensor is drawn from its own generator (seed = crc32(key) ^ seed), so
the result does not depend on enumeration order and is identical on every machine with the
same torch CPU RNG.  Tensors the reference zero-initialises (zero_module) are re-randomised
with N(0, 1/fan_in) — otherwise eps == 0 and control == 0 and every parity check is vacuous
(SURVEY.md §8d).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Callable, Dict, Optional, Tuple

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def make_state_dict(shapes: "OrderedDict[str, Tuple[int, ...]]", seed: int = 1234,
                    zero_init: Optional[Callable[[str], bool]] = None,
                    prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
    sd = OrderedDict()
    for key, shape in shapes.items():
        g = _gen(key, seed)
        if len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if key.endswith("relative_position_bias_table") or "embedding" in key:
                t = torch.randn(shape, generator=g) * 0.02
            elif zero_init is not None and zero_init(key):
                t = torch.randn(shape, generator=g) * fan_in ** -0.5
            else:
                b = fan_in ** -0.5
                t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif len(shape) == 1 and key.endswith("weight"):      # norm gains
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 0:
            t = torch.tensor(2.6593)
        else:                                                  # biases / norm shifts
            t = 0.02 * torch.randn(shape, generator=g)
        sd[prefix + key] = t
    return sd
