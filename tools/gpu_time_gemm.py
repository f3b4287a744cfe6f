"""Times dbir_gemm on the shapes of one ControlNet+UNet forward (batch 2, latent 64x64).
   python tools/gpu_time_gemm.py        (DBIR_GEMM_PAIR=0|1, DBIR_GEMM_AUTOTUNE=0 to constrain the plans)"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402

SHAPES = [  # kind, M, N, K, (n,h,w,c) for convs, count per forward
    ("conv", 8192, 320, 2880, (2, 64, 64, 320), 11), ("conv", 2048, 640, 5760, (2, 32, 32, 640), 9),
    ("conv", 512, 1280, 11520, (2, 16, 16, 1280), 10), ("conv", 128, 1280, 11520, (2, 8, 8, 1280), 19),
    ("conv", 128, 1280, 23040, (2, 8, 8, 2560), 3), ("conv", 512, 1280, 23040, (2, 16, 16, 2560), 2),
    ("conv", 8192, 320, 5760, (2, 64, 64, 640), 2), ("conv", 2048, 1280, 11520, (2, 32, 32, 1280), 1),
    ("conv", 8192, 640, 5760, (2, 64, 64, 640), 1), ("conv", 2048, 640, 17280, (2, 32, 32, 1920), 1),
    ("gemm", 8192, 320, 320, None, 38), ("gemm", 2048, 640, 640, None, 37), ("gemm", 512, 1280, 1280, None, 37),
    ("gemm", 8192, 960, 320, None, 7), ("gemm", 2048, 1920, 640, None, 7), ("gemm", 512, 3840, 1280, None, 7),
    ("gemm", 8192, 320, 1280, None, 7), ("gemm", 2048, 640, 2560, None, 7), ("gemm", 512, 1280, 5120, None, 7),
    ("gemm", 128, 1280, 1280, None, 14), ("geglu", 8192, 2560, 320, None, 7), ("geglu", 2048, 5120, 640, None, 7),
    ("geglu", 512, 10240, 1280, None, 7),
]


def main():
    dt = lib.operand_dtype()
    dev = "cuda"
    ws = torch.zeros(16 * 1024 * 1024 + 16384, device=dev)
    tot = 0.0
    for kind, M, N, K, conv, cnt in SHAPES:
        a = torch.randn(M, K if conv is None else conv[3], device=dev).to(dt)
        if conv is not None:
            a = torch.randn(*conv, device=dev).to(dt)
        b = (torch.randn(N, K, device=dev) * K ** -0.5).to(dt)
        bias = torch.randn(N, device=dev)
        geglu = kind == "geglu"
        nout = N // 2 if geglu else N
        out = torch.empty(M, nout, device=dev, dtype=dt if geglu or (kind == "gemm" and N >= 960) else torch.float32)
        res = torch.randn(M, nout, device=dev) if out.dtype == torch.float32 else None
        fbn = 128 if geglu else 0

        def call():
            lib.gemm(a, b, out, M=M, N=N, K=K, bias=bias, residual=res, conv=(conv + (3,)) if conv else None,
                     geglu=geglu, force_bn=fbn, splitk_ws=ws)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                call()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 10
        tot += us * cnt
        print(f"{kind:5s} M={M:5d} N={N:5d} K={K:5d} x{cnt:2d}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")
    print(f"weighted total per forward: {tot / 1e3:.2f} ms")


if __name__ == "__main__":
    main()
