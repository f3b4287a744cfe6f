"""Makes `import diffbir...` resolve to the REFERENCE checkout (/root/reference) inside the fixture
generators. The reference's `diffbir` is a namespace package (no __init__.py), so the repo's own
`diffbir/` alias package (a regular package) would shadow it on any sys.path order: the generators
therefore pin `sys.modules["diffbir"]` to a package object whose search path is the reference's."""
import sys
import types
import typing
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
REFERENCE = Path("/root/reference")


def use_reference() -> None:
    if not (REFERENCE / "diffbir").is_dir():
        raise SystemExit("the fixture generators need the reference checkout at /root/reference")
    for p in (str(ROOT), str(ROOT / "oracle" / "_shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in [m for m in sys.modules if m == "diffbir" or m.startswith("diffbir.")]:
        del sys.modules[name]
    pkg = types.ModuleType("diffbir")
    pkg.__path__ = [str(REFERENCE / "diffbir")]
    sys.modules["diffbir"] = pkg
    torch.Tuple = typing.Tuple      # sampler/edm_sampler.py:145 annotation no longer exists in torch 2.11
