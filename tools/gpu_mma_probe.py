"""Mainloop cycles per k-block with parts of the loop disabled (DBIR_GEMM_DBG: 1 half the MMAs,
2 no TMA loads, 4 no MMAs) -- separates tensor-pipe time from operand-feed time."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402

dt = lib.operand_dtype()
M, N, K = 2048, 1280, 2560
a = torch.randn(M, K, device="cuda").to(dt)
b = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
out = torch.empty(M, N, device="cuda")
for fbn in (32, 64, 128, 256):
    for pair in (2, 1):
        if pair == 1 and fbn < 64:
            continue
        for flags in (0, 2, 4, 6):
            os.environ["DBIR_GEMM_DBG"] = str(flags)
            dbg = torch.zeros(8192 * 8, dtype=torch.int64, device="cuda")
            for i in range(3):
                lib.gemm(a, b, out, M=M, N=N, K=K, force_bn=fbn, split_k=1, cta_pair=pair, debug_stamps=dbg)
            torch.cuda.synchronize()
            d = dbg.view(-1, 8).cpu()
            d = d[d[:, 3] > 0].double()
            main = d[:, 2] - d[:, 1]
            kb = K // 64
            print(f"bn={fbn:3d} pair={pair == 1} flags={flags}: ctas {len(d):3d} mainloop {main.mean():7.0f} = {main.mean() / kb:5.0f} cycles/kb | "
                  f"consumer wait {d[:, 4].mean() / kb:5.0f} issue {d[:, 5].mean() / kb:5.0f} | producer wait {d[:, 6].mean() / kb:5.0f} "
                  f"issue {d[:, 7].mean() / kb:5.0f}")
