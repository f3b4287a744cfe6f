"""Per-CTA phase timing (clock64) of dbir_gemm, single-CTA vs CTA-pair (cta_group::2) mainloops."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402

dt = lib.operand_dtype()
CASES = [("gemm", 2048, 640, 2560, None), ("gemm", 8192, 320, 1280, None), ("gemm", 8192, 1280, 1280, None),
         ("conv", 8192, 320, 2880, (2, 64, 64, 320)), ("conv", 512, 1280, 11520, (2, 16, 16, 1280))]
for kind, M, N, K, conv in CASES:
    a = torch.randn(*(conv if conv else (M, K)), device="cuda").to(dt)
    b = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
    out = torch.empty(M, N, device="cuda")
    bias = torch.randn(N, device="cuda")
    for fbn in (32, 64, 128, 160, 256):
        if N % fbn:
            continue
        for pair in (2, 1):
            if pair == 1 and fbn < 64:
                continue
            dbg = torch.zeros(8192 * 8, dtype=torch.int64, device="cuda")
            for i in range(3):
                lib.gemm(a, b, out, M=M, N=N, K=K, bias=bias, conv=(conv + (3,)) if conv else None, force_bn=fbn,
                         split_k=1, cta_pair=pair, debug_stamps=dbg)
            torch.cuda.synchronize()
            d = dbg.view(-1, 8).cpu()
            d = d[d[:, 3] > 0].double()
            setup, main, epi = (d[:, 1] - d[:, 0]), (d[:, 2] - d[:, 1]), (d[:, 3] - d[:, 2])
            span = d[:, 3].max() - d[:, 0].min()
            kb = K // 64
            print(f"{kind} M={M} N={N} K={K} bn={fbn} pair={pair == 1}: ctas {len(d)} | setup {setup.mean():.0f} | mainloop "
                  f"{main.mean():.0f} ({main.mean() / kb:.0f}/kb, max {main.max():.0f}) | epilogue {epi.mean():.0f} | "
                  f"span {span:.0f} cycles")
