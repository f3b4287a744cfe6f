"""Sweeps the tile width (and split-K) of dbir_gemm for the forward's shapes: prints in-graph us."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402

dt = lib.operand_dtype()
dev = "cuda"
ws = torch.zeros(16 * 1024 * 1024 + 16384, device=dev)
SHAPES = [("gemm", 8192, 320, 320, None, True), ("gemm", 2048, 640, 640, None, True), ("gemm", 512, 1280, 1280, None, True),
          ("gemm", 8192, 960, 320, None, False), ("gemm", 2048, 1920, 640, None, False), ("gemm", 512, 3840, 1280, None, False),
          ("gemm", 8192, 320, 1280, None, True), ("gemm", 2048, 640, 2560, None, True), ("gemm", 512, 1280, 5120, None, True),
          ("gemm", 128, 1280, 1280, None, True),
          ("conv", 8192, 320, 2880, (2, 64, 64, 320), True), ("conv", 2048, 640, 5760, (2, 32, 32, 640), True),
          ("conv", 512, 1280, 11520, (2, 16, 16, 1280), True), ("conv", 128, 1280, 11520, (2, 8, 8, 1280), True)]


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for kind, M, N, K, conv, res in SHAPES:
    a = torch.randn(*conv, device=dev).to(dt) if conv else torch.randn(M, K, device=dev).to(dt)
    b = (torch.randn(N, K, device=dev) * K ** -0.5).to(dt)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if res else dt)
    r = torch.randn(M, N, device=dev) if res else None
    row = [f"{kind} {M}x{N}x{K}:"]
    auto = timeit(lambda: lib.gemm(a, b, out, M=M, N=N, K=K, bias=bias, residual=r, conv=(conv + (3,)) if conv else None, splitk_ws=ws))
    row.append(f"auto {auto:6.1f}")
    for bn in (32, 64, 128, 160, 256):
        if N % bn and bn > 64:
            continue
        for sk in (1, 2, 4, 8):
            kb = (K + 63) // 64
            if sk > 1 and (kb // sk < 6 or M > 512):
                continue
            try:
                us = timeit(lambda: lib.gemm(a, b, out, M=M, N=N, K=K, bias=bias, residual=r,
                                             conv=(conv + (3,)) if conv else None, splitk_ws=ws, force_bn=bn, split_k=sk))
                row.append(f"bn{bn}/s{sk} {us:6.1f}")
            except Exception as e:
                row.append(f"bn{bn}/s{sk} ERR")
    print("  ".join(row), flush=True)
