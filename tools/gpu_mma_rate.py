"""(needs the probe build: python -m diffbir_b200.build --tag=probes -DDBIR_DEBUG_PROBES -DDBIR_ATTN_PROBE;
run with DBIR_LIB_TAG=probes)
tcgen05.mma retire rate by tile shape / operand layout (calibrates the GEMM / attention models)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402

L = lib.load()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for a_tmem in (0, 1, 2):
    for b_mn in (0,):
        for n in (32, 64, 128, 256):
            res = []
            for iters in (64, 256):
                for _ in range(2):
                    lib.check(L.dbir_debug_mma_rate(n, b_mn, iters, a_tmem, C.c_void_p(out.data_ptr()), C.c_void_p(0)), "probe")
                    torch.cuda.synchronize()
                res.append(int(out.item()))
            per = (res[1] - res[0]) / (256 - 64)
            print(f"A {['smem', 'TMEM', 'smem->TMEM cp + TMEM'][a_tmem]}, B {'MN' if b_mn else 'K '}-major, M=128 N={n:3d} K=16: "
                  f"{per:6.1f} cycles per MMA  ({128 * n * 16 / per:7.0f} MAC/clk)  [64: {res[0]}, 256: {res[1]}]")
