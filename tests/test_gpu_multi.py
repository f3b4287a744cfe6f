"""-m gpu, >= 2 GPUs: tiled sampling sharded over NCCL ranks is bit-identical to single-rank tiled
sampling and every rank holds the bit-identical latent (tools/run_tiled_multi.py under torchrun).
On a 1-GPU box the multi-rank part cannot run; test_batch_invariance_single_gpu covers the property
the sharded path relies on (a tile's eps does not depend on the batch it runs in)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_sharded_tiles_bit_identical():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(ROOT / "tools" / "run_tiled_multi.py")],
                       capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-3000:])
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "tiled_multi.log").write_text(r.stdout + "\n" + r.stderr)
    assert r.returncode == 0
