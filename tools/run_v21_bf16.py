"""Subprocess of tests/test_gpu_pipeline.py::test_full_config_v21_1024_batch2_fp16_and_bf16: runs the
v2.1 1024^2 batch-2 case on the bf16-operand build (DBIR_OPERANDS=bf16 must be set by the caller) and
prints {"psnr": ..} of its uint8 output against the fp32 oracle output saved by the parent."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from diffbir_b200 import lib  # noqa: E402
from diffbir_b200.utils.synth import build_synthetic_pipeline, synthetic_lq  # noqa: E402
from tests.test_gpu_pipeline import V21_KW, _psnr_u8  # noqa: E402


def main():
    ref = np.load(sys.argv[1])
    pipe = build_synthetic_pipeline("cuda", seed=1234, small=False, v_prediction=True)
    lq = synthetic_lq(1024, 1024, batch=2, seed=5)
    torch.manual_seed(231)
    out = pipe.run(lq, **V21_KW)
    print(json.dumps(dict(psnr=_psnr_u8(out, ref), operand_dtype=str(lib.operand_dtype()))))


if __name__ == "__main__":
    main()
