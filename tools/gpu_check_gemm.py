"""Bring-up check of the tcgen05 GEMM / implicit conv kernel on a real B200.

    python tools/gpu_check_gemm.py            # runs every case in its own subprocess
    python tools/gpu_check_gemm.py <case>     # one case in-process

Checker = torch fp32 on the GPU with TF32 off, fed the same 16-bit-rounded operands.
"""
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def rel_err(x, ref):
    return ((x.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


def run_case(name):
    import torch
    import torch.nn.functional as F
    from diffbir_b200 import lib

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    dt = lib.operand_dtype()
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*s, scale=1.0):
        return (torch.randn(*s, device=dev, generator=g) * scale)

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    if name.startswith("plain"):
        cfg = {
            "plain_tiny": (256, 64, 64, 0),
            "plain_bn32": (300, 24, 128, 0),
            "plain_320": (4096 + 13, 320, 320, 0),
            "plain_640": (2048, 640, 2560, 0),
            "plain_1280": (512, 1280, 1280, 0),
            "plain_f160": (1024, 320, 640, 160),
            "plain_f256": (1024, 1280, 640, 256),
            "plain_f128": (1024, 640, 640, 128),
            "plain_k200": (777, 200, 200, 0),
        }[name]
        M, N, K, fbn = cfg
        a = rnd(M, K).to(dt)
        b = rnd(N, K, scale=K ** -0.5).to(dt)
        bias = rnd(N)
        res = rnd(M, N)
        out = torch.empty(M, N, device=dev)
        lib.gemm(a, b, out, M=M, N=N, K=K, bias=bias, residual=res, force_bn=fbn)
        torch.cuda.synchronize()
        ref = a.float() @ b.float().t() + bias + res
        e = rel_err(out, ref)
        # 16-bit output + activation + alpha
        out2 = torch.empty(M, N, device=dev, dtype=dt)
        lib.gemm(a, b, out2, M=M, N=N, K=K, bias=bias, act="gelu", alpha=0.5, force_bn=fbn)
        torch.cuda.synchronize()
        ref2 = 0.5 * F.gelu(a.float() @ b.float().t() + bias)
        e2 = rel_err(out2, ref2)
        ms = timeit(lambda: lib.gemm(a, b, out, M=M, N=N, K=K, bias=bias, residual=res, force_bn=fbn))
        print(f"{name}: M={M} N={N} K={K} rel_err fp32={e:.3e} op16/gelu={e2:.3e} "
              f"{ms*1e3:.1f} us {2*M*N*K/ms/1e9:.1f} TFLOP/s")
        ok = e < 2e-5 and e2 < 4e-3
    elif name.startswith("rowvec"):
        M, N, K = 2 * 1024, 640, 320
        a = rnd(M, K).to(dt)
        b = rnd(N, K, scale=K ** -0.5).to(dt)
        rv = rnd(2, N)
        out = torch.empty(M, N, device=dev)
        lib.gemm(a, b, out, M=M, N=N, K=K, rowvec=rv, rows_per_vec=1024)
        torch.cuda.synchronize()
        ref = (a.float() @ b.float().t()).view(2, 1024, N) + rv[:, None, :]
        e = rel_err(out, ref.view(M, N))
        print(f"{name}: rel_err={e:.3e}")
        ok = e < 2e-5
    elif name.startswith("geglu"):
        M, C, fbn = {"geglu_320": (4096, 320, 160), "geglu_1280": (300, 1280, 256),
                     "geglu_640": (1024, 640, 128)}[name]
        inner = 4 * C
        a = rnd(M, C).to(dt)
        w = rnd(2 * inner, C, scale=C ** -0.5).to(dt)     # rows: [values | gates] (torch chunk order)
        bias = rnd(2 * inner)
        hb = fbn // 2
        # pack per tile: tile j = [values j*hb..(j+1)*hb | gates ...]
        wv, wg = w[:inner].view(inner // hb, hb, C), w[inner:].view(inner // hb, hb, C)
        wp = torch.cat([wv, wg], dim=1).reshape(2 * inner, C).contiguous()
        bv, bg = bias[:inner].view(inner // hb, hb), bias[inner:].view(inner // hb, hb)
        bp = torch.cat([bv, bg], dim=1).reshape(2 * inner).contiguous()
        out = torch.empty(M, inner, device=dev, dtype=dt)
        lib.gemm(a, wp, out, M=M, N=2 * inner, K=C, bias=bp, geglu=True, force_bn=fbn)
        torch.cuda.synchronize()
        h = a.float() @ w.float().t() + bias
        ref = h[:, :inner] * F.gelu(h[:, inner:])
        e = rel_err(out, ref)
        ms = timeit(lambda: lib.gemm(a, wp, out, M=M, N=2 * inner, K=C, bias=bp, geglu=True, force_bn=fbn))
        print(f"{name}: rel_err={e:.3e} {ms*1e3:.1f} us {2*M*2*inner*C/ms/1e9:.1f} TFLOP/s")
        ok = e < 4e-3
    elif name.startswith("conv"):
        n, h, w, cin, cout, ks = {
            "conv_64_320": (2, 64, 64, 320, 320, 3),
            "conv_32_640": (2, 32, 32, 640, 640, 3),
            "conv_16_1280": (2, 16, 16, 1280, 1280, 3),
            "conv_8_1280": (2, 8, 8, 1280, 1280, 3),
            "conv_8_b1": (1, 8, 8, 1280, 1280, 3),
            "conv_cat": (2, 32, 32, 1920, 640, 3),
            "conv_odd": (1, 40, 72, 128, 192, 3),
            "conv_vae512": (1, 512, 512, 128, 128, 3),
            "conv_1x1": (2, 16, 16, 1280, 1280, 1),
        }[name]
        x = rnd(n, h, w, cin).to(dt)                       # NHWC
        wt = rnd(cout, cin, ks, ks, scale=(cin * ks * ks) ** -0.5).to(dt)
        bias = rnd(cout)
        rv = rnd(n, cout)
        wp = wt.permute(0, 2, 3, 1).reshape(cout, ks * ks * cin).contiguous()
        out = torch.empty(n * h * w, cout, device=dev)
        K = ks * ks * cin

        def call():
            lib.gemm(x, wp, out, M=n * h * w, N=cout, K=K, bias=bias, rowvec=rv,
                     conv=(n, h, w, cin, ks))
        call()
        torch.cuda.synchronize()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=ks // 2)
        ref = ref + rv[:, :, None, None]
        ref = ref.permute(0, 2, 3, 1).reshape(n * h * w, cout)
        e = rel_err(out, ref)
        ms = timeit(call)
        print(f"{name}: rel_err={e:.3e} {ms*1e3:.1f} us {2*n*h*w*cout*K/ms/1e9:.1f} TFLOP/s")
        ok = e < 2e-5
    else:
        raise SystemExit(f"unknown case {name}")
    print("PASS" if ok else "FAIL", name)
    return ok


CASES = ["plain_tiny", "plain_bn32", "plain_320", "plain_640", "plain_1280", "plain_f160",
         "plain_f256", "plain_f128", "plain_k200", "rowvec", "geglu_320", "geglu_640",
         "geglu_1280", "conv_64_320", "conv_32_640", "conv_16_1280", "conv_8_1280", "conv_8_b1",
         "conv_cat", "conv_odd", "conv_vae512", "conv_1x1"]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(0 if run_case(sys.argv[1]) else 1)
    bad = []
    for c in CASES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True,
                               timeout=180)
            out = (r.stdout + r.stderr).strip().splitlines()
            tail = "\n".join(out[-6:])
            print(f"[{c}] rc={r.returncode} {time.time()-t0:.1f}s\n{tail}", flush=True)
            if r.returncode != 0:
                bad.append(c)
        except subprocess.TimeoutExpired:
            print(f"[{c}] TIMEOUT", flush=True)
            bad.append(c)
    print("FAILED:", bad)
    sys.exit(1 if bad else 0)
