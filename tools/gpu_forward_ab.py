"""A/B timing of the ControlNet+UNet forward (batch 2 = cond/uncond of one 512^2 image) on the GPU box.

    [DBIR_LIB_TAG=<tag>] [DBIR_GEMM_WIDE=0] ... python tools/gpu_forward_ab.py [label] [--census]

Prints one line: CUDA-graph replay ms per forward (two streams and one stream), and with --census the
GEMM / attention family replay times + a per-shape table. Weights are drawn on the GPU (fast start;
numerics are irrelevant here), every run is a fresh process so plan caches / env switches are clean."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import arch, lib  # noqa: E402
from diffbir_b200.engine.cldm import CldmEngine  # noqa: E402


def fast_sd(shapes, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if len(shp) >= 2:
            fan = 1
            for d in shp[1:]:
                fan *= d
            sd[k] = (torch.rand(shp, device="cuda", generator=g) * 2 - 1) * fan ** -0.5
        elif k.endswith("weight"):
            sd[k] = 1.0 + 0.05 * torch.randn(shp, device="cuda", generator=g)
        else:
            sd[k] = 0.02 * torch.randn(shp, device="cuda", generator=g)
    return sd


def replay_ms(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    label = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "run"
    nb = 2
    for a in sys.argv:
        if a.startswith("--nb="):
            nb = int(a.split("=")[1])
    t0 = time.time()
    lib.load()
    usd = fast_sd(arch.unet_shapes(arch.UNET_CFG), 1)
    csd = fast_sd(arch.unet_shapes(arch.CONTROLNET_CFG, True), 2)
    eng = CldmEngine(usd, csd, None, None, "cuda")
    del usd, csd
    x = torch.randn(nb, 4, 64, 64, device="cuda")
    ci = torch.randn(nb, 4, 64, 64, device="cuda") * 0.5
    out = torch.empty_like(x)
    eng.set_context(torch.randn(nb, 77, 1024, device="cuda"))
    eng.set_timesteps([500], nb)
    eng.load_step(0)
    res = {"label": label, "nb": nb}
    for two in (False, True):
        eng.two_streams = two
        eng.forward(x, ci, [1.0] * 13, out=out)            # tunes plans (first pass), sizes buffers
        eng.forward(x, ci, [1.0] * 13, out=out)
        torch.cuda.synchronize()
        res["two_streams_ms" if two else "one_stream_ms"] = replay_ms(lambda: eng.forward(x, ci, [1.0] * 13, out=out))
    res["finite"] = bool(torch.isfinite(out).all())
    if "--census" in sys.argv:
        eng.two_streams = False
        lib.record_begin()
        eng.forward(x, ci, [1.0] * 13, out=out)
        calls = lib.record_end()
        fam = {}
        for kind in ("gemm+conv", "attention"):
            sel = [c for c in calls if (c[0] == "attention") == (kind == "attention")]
            fl = sum(c[2] for c in sel)
            ms = replay_ms(lambda: [c[3]() for c in sel], 5)
            fam[kind] = dict(launches=len(sel), gflop=fl / 1e9, ms=ms, tflops=fl / ms / 1e9)
        res["families"] = fam
        by = {}
        for c in calls:
            by.setdefault((c[0],) + tuple(c[1]), []).append(c)
        rows = []
        for k, sel in by.items():
            ms = replay_ms(lambda: [c[3]() for c in sel], 3)
            fl = sum(c[2] for c in sel)
            rows.append((ms, k, len(sel), fl))
        rows.sort(reverse=True)
        Path("gpurun_out").mkdir(exist_ok=True)
        with open(f"gpurun_out/census_{label}.csv", "w") as f:
            f.write("kind,shape,launches,gflop,ms,us_per_launch,tflops\n")
            for ms, k, n, fl in rows:
                f.write(f"{k[0]},{'x'.join(map(str, k[1:]))},{n},{fl / 1e9:.2f},{ms:.4f},{ms * 1e3 / n:.1f},{fl / ms / 1e9:.1f}\n")
    res["wall_s"] = round(time.time() - t0, 1)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
