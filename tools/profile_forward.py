"""Profiling helper (GPU box).

  python tools/profile_forward.py stages      # CUDA-event breakdown of one 512^2 image
  python tools/profile_forward.py forward N   # N eager ControlNet+UNet forwards (batch 2) for ncu:
      ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <count> --csv \
          --log-file gpurun_out/forward_launches.csv python tools/profile_forward.py forward 3
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402
from diffbir_b200.utils.synth import RUN_DEFAULTS, build_synthetic_pipeline, synthetic_lq  # noqa: E402


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def stages():
    pipe = build_synthetic_pipeline("cuda", 1234)
    lq = torch.from_numpy(synthetic_lq(512, 512)).cuda()
    kw = dict(RUN_DEFAULTS)
    for _ in range(2):
        torch.manual_seed(231)
        pipe.run_device(lq, **kw)
    torch.cuda.synchronize()
    # stage timing by calling the pieces like run_device does
    from diffbir_b200.pipeline import pad_to_multiples_of
    from diffbir_b200.sampler import SpacedSampler
    torch.manual_seed(231)
    t = {}
    e0 = ev()
    x = lq.float().div(255).clamp(0, 1).permute(0, 3, 1, 2).contiguous()
    clean = pipe.apply_cleaner(x, False, 512, 256)
    e1 = ev()
    cond_img = pad_to_multiples_of(clean, 64)
    cond = pipe.cldm.prepare_condition(cond_img, [""])
    uncond = dict(c_txt=pipe.cldm.clip(pipe.cldm.tokenize([kw["neg_prompt"]])), c_img=cond["c_img"].clone())
    e2 = ev()
    smp = SpacedSampler(pipe.diffusion.betas, "eps", False)
    w0 = time.perf_counter()
    z = smp.sample(pipe.cldm, "cuda", 50, (1, 4, 64, 64), cond, uncond, 4.0)
    e3 = ev()
    dec = pipe.cldm.vae_decode(z)
    e4 = ev()
    torch.cuda.synchronize()
    print(f"swinir {e0.elapsed_time(e1):.1f} ms | vae-encode+clip {e1.elapsed_time(e2):.1f} ms | "
          f"sampler(50) {e2.elapsed_time(e3):.1f} ms (host wall {1e3 * (time.perf_counter() - w0):.1f}) | "
          f"vae-decode {e3.elapsed_time(e4):.1f} ms")
    # inside the sampler: graph replay time alone
    eng = pipe.cldm.engine
    xin = torch.randn(2, 4, 64, 64, device="cuda")
    ci = torch.randn(2, 4, 64, 64, device="cuda")
    out = torch.empty_like(xin)
    eng.load_step(0)
    eng.forward(xin, ci, [1.0] * 13, out=out)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.forward(xin, ci, [1.0] * 13, out=out)
    torch.cuda.synchronize()
    cap = time.perf_counter() - w0
    a = ev()
    for _ in range(10):
        g.replay()
    b = ev()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    eng.set_timesteps(list(range(50)), 2)
    torch.cuda.synchronize()
    st = time.perf_counter() - w0
    print(f"graph capture {cap * 1e3:.1f} ms | graph replay {a.elapsed_time(b) / 10:.2f} ms per forward(batch 2) | "
          f"set_timesteps {st * 1e3:.1f} ms | launches/forward {lib.launches()}")


def forward(n):
    pipe = build_synthetic_pipeline("cuda", 1234)
    pipe.cldm._build()
    eng = pipe.cldm.engine
    xin = torch.randn(2, 4, 64, 64, device="cuda")
    ci = torch.randn(2, 4, 64, 64, device="cuda")
    eng.set_context(torch.randn(2, 77, 1024, device="cuda"))
    eng.set_timesteps([500], 2)
    eng.load_step(0)
    torch.cuda.synchronize()
    n0 = lib.launches()
    for i in range(n):
        if i == 1:
            torch.cuda.nvtx.range_push("dbir_fwd")
        eng.forward(xin, ci, [1.0] * 13)
        torch.cuda.synchronize()
        if i == 1:
            torch.cuda.nvtx.range_pop()
        if i == 0:
            print("launches per forward:", lib.launches() - n0, "launches before first forward:", n0)


if __name__ == "__main__":
    if sys.argv[1] == "stages":
        stages()
    else:
        forward(int(sys.argv[2]))
