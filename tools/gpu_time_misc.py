"""In-graph timing of the non-GEMM kernels at the shapes of one ControlNet+UNet forward (batch 2)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402

dt = lib.operand_dtype()
dev = "cuda"


def timeit(fn, n=10):
    for _ in range(3):
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


tot = 0.0
for b, h, sq, skv, cnt in [(2, 5, 4096, 4096, 7), (2, 10, 1024, 1024, 7), (2, 20, 256, 256, 7), (2, 20, 64, 64, 2),
                            (2, 5, 4096, 77, 7), (2, 10, 1024, 77, 7), (2, 20, 256, 77, 7), (2, 20, 64, 77, 2)]:
    c = h * 64
    if sq == skv:
        qkv = torch.randn(b * sq, 3 * c, device=dev).to(dt)
        q, k, v, ldq, ldk = qkv, qkv[:, c:], qkv[:, 2 * c:], 3 * c, 3 * c
    else:
        q = torch.randn(b * sq, c, device=dev).to(dt)
        kv = torch.randn(b * skv, 2 * c, device=dev).to(dt)
        k, v, ldq, ldk = kv, kv[:, c:], c, 2 * c
    out = torch.empty(b * sq, c, device=dev, dtype=dt)
    nws = lib.attention_ws_bytes(b, h, sq, skv)
    aws = torch.zeros(nws // 4, device=dev) if nws else None
    us = timeit(lambda: lib.attention(q, k, v, out, batch=b, heads=h, sq=sq, skv=skv, ldq=ldq, ldk=ldk, ldv=ldk, ldo=c, ws=aws))
    tot += us * cnt
    print(f"attention b={b} h={h:2d} sq={sq:4d} skv={skv:4d} x{cnt}: {us:7.1f} us  {4.0 * b * h * sq * skv * 64 / us / 1e6:6.1f} TFLOP/s")
print(f"attention weighted total per forward: {tot / 1e3:.2f} ms")

tot = 0.0
for n, hw, c1, c2, cnt in [(2, 4096, 320, 0, 30), (2, 1024, 640, 0, 20), (2, 256, 1280, 0, 18), (2, 64, 1280, 0, 16),
                            (2, 64, 1280, 1280, 3), (2, 256, 1280, 1280, 2), (2, 1024, 640, 640, 1), (2, 4096, 320, 320, 2)]:
    h = w = int(hw ** 0.5)
    s1 = torch.randn(n, hw, c1, device=dev)
    s2 = torch.randn(n, hw, c2, device=dev) if c2 else None
    c = c1 + c2
    stats = torch.empty(n * 64, device=dev)
    ws = torch.zeros(lib.gn_workspace_floats(n, hw, c), device=dev)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    out = torch.empty(n, hw, c, device=dev, dtype=dt)
    us1 = timeit(lambda: lib.gn_stats(s1, s2, c1, c2, n, hw, 1e-5, stats, ws))
    us2 = timeit(lambda: lib.gn_apply(s1, s2, c1, c2, n, h, w, stats, gamma, beta, out))
    tot += (us1 + us2) * cnt
    print(f"groupnorm n={n} hw={hw:5d} c={c1}+{c2} x{cnt}: stats {us1:6.1f} us  apply {us2:6.1f} us")
print(f"groupnorm weighted total per forward: {tot / 1e3:.2f} ms")
for rows, c, cnt in [(8192, 320, 21), (2048, 640, 21), (512, 1280, 21), (128, 1280, 6)]:
    x = torch.randn(rows, c, device=dev)
    g_, b_ = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    out = torch.empty(rows, c, device=dev, dtype=dt)
    us = timeit(lambda: lib.layernorm(x, c, rows, c, g_, b_, out, c))
    print(f"layernorm rows={rows} c={c} x{cnt}: {us:6.1f} us")
