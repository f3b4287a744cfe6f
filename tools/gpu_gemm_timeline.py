"""Per-CTA phase timing (clock64) of dbir_gemm for a few shapes: setup / mainloop / epilogue."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402

dt = lib.operand_dtype()
for (M, N, K, out16, res, geglu, fbn) in [(8192, 320, 320, False, True, False, 0), (8192, 320, 1280, False, True, False, 0),
                                            (8192, 960, 320, True, False, False, 0), (8192, 2560, 320, True, False, True, 256),
                                            (8192, 2560, 320, True, False, True, 128), (2048, 640, 640, False, True, False, 0),
                                            (8192, 320, 320, False, False, False, 0), (8192, 320, 320, True, False, False, 0)]:
    a = torch.randn(M, K, device="cuda").to(dt)
    b = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
    nout = N // 2 if geglu else N
    out = torch.empty(M, nout, device="cuda", dtype=dt if out16 else torch.float32)
    r = torch.randn(M, nout, device="cuda") if res else None
    bias = torch.randn(N, device="cuda")
    dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
    for i in range(3):
        lib.gemm(a, b, out, M=M, N=N, K=K, bias=bias, residual=r, geglu=geglu, force_bn=fbn, debug_stamps=dbg)
    torch.cuda.synchronize()
    d = dbg.view(-1, 8).cpu()
    d = d[d[:, 3] > 0].double()
    setup, main, epi = (d[:, 1] - d[:, 0]), (d[:, 2] - d[:, 1]), (d[:, 3] - d[:, 2])
    print(f"M={M} N={N} K={K} out16={out16} res={res} geglu={geglu} bn={fbn}: ctas {len(d)} | setup {setup.mean():.0f} "
          f"| mainloop {main.mean():.0f} (max {main.max():.0f}) | epilogue {epi.mean():.0f} (max {epi.max():.0f}) cycles")
