"""bench.py — headline benchmark of the DiffBIR hot path (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 512|tiled2048]

A "step" is one full restoration of one synthetic image per GPU through SwinIRPipeline: SwinIR
stage 1 -> VAE encode + CLIP -> 50-step spaced sampler (cfg 4.0, ControlNet + UNet, batched
cond/uncond) -> VAE decode -> colour fix -> uint8.  Random-init SD-2.1 / SwinIR weights
(no network for checkpoints), synthetic image-like input.

  value : MPix/s with the uint8 input already resident in HBM (CUDA-event time, max over ranks)
  e2e   : the same through the public API Pipeline.run(host uint8) -> host uint8 (H2D + D2H inside)
  roofline : tcgen05 GEMM/conv kernel family — algorithmic FLOPs of every launch of one
             ControlNet+UNet forward / CUDA-event time of those launches, vs the measured bf16 peak
  cpu_baseline : the fp32 oracle port of the reference on the host cores, bounded sample
                 (1 of 50 sampler steps; SwinIR, VAE encode/decode once), extrapolated.

Multi-GPU (torchrun, one rank per GPU): the 512x512 workload has no tiles, so ranks are
independent replicas (weak scaling, no collective); --workload tiled2048 shards the 49 latent
tiles of a 2048x2048 image over the ranks with one NCCL all-gather per step (strong scaling).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "MPix/s end-to-end 50-step restore @512px"
SAMPLER_STEPS = 50


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_times(n_steps: int = 1, warm: int = 0, size: int = 512):
    """Times the reference algorithm (oracle port, fp32, all host threads) on a bounded sample of
    the 512x512 workload: SwinIR forward, VAE encode, `n_steps` sampler steps (2 forwards each, as
    spaced_sampler.py:156-157), VAE decode.  Returns component seconds."""
    import torch
    from diffbir_b200 import arch
    from diffbir_b200.utils.synth import make_state_dict, synthetic_lq
    from oracle import cldm as ocl
    from oracle import swinir as osw
    torch.manual_seed(231)
    L = size // 8
    t = {}
    with torch.no_grad():
        ssd = make_state_dict(arch.swinir_shapes(arch.SWINIR_CFG), 1238)
        x = torch.tensor(synthetic_lq(size, size)).float().div(255).permute(0, 3, 1, 2).contiguous()
        t0 = time.perf_counter(); clean = osw.swinir_forward(ssd, x); t["swinir"] = time.perf_counter() - t0
        del ssd
        vsd = make_state_dict(arch.vae_shapes(arch.VAE_CFG), 1235)
        t0 = time.perf_counter(); c_img = ocl.vae_encode_mode(vsd, clean * 2 - 1); t["vae_encode"] = time.perf_counter() - t0
        t0 = time.perf_counter(); ocl.vae_decode(vsd, c_img / 0.18215); t["vae_decode"] = time.perf_counter() - t0
        del vsd
        usd = make_state_dict(arch.unet_shapes(arch.UNET_CFG), 1234, arch.is_zero_init)
        csd = make_state_dict(arch.unet_shapes(arch.CONTROLNET_CFG, True), 1237, arch.is_zero_init)
        xt = torch.randn(1, 4, L, L)
        ctx = torch.randn(1, 77, 1024)
        tt = torch.full((1,), 999)
        steps = []
        for i in range(warm + n_steps):
            t0 = time.perf_counter()
            ec = ocl.cldm_forward(usd, csd, xt, tt, ctx, c_img, [1.0] * 13)
            eu = ocl.cldm_forward(usd, csd, xt, tt, ctx, c_img, [1.0] * 13)
            xt = xt - 0.01 * (eu + 4.0 * (ec - eu))
            dt = time.perf_counter() - t0
            if i >= warm:
                steps.append(dt)
        t["sampler_step"] = steps
    return t, torch.get_num_threads()


def cpu_image_seconds(t, step_s):
    # the reference encodes the condition image twice (pipeline.py:117-128)
    return t["swinir"] + 2 * t["vae_encode"] + SAMPLER_STEPS * step_s + t["vae_decode"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t, cores = cpu_reference_times(n_steps=args.steps, warm=args.warmup)
    step_s = statistics.median(t["sampler_step"])
    total = cpu_image_seconds(t, step_s)
    mpix = 512 * 512 / 1e6 / total
    sample = (f"oracle port of the reference (fp32, {cores} threads): SwinIR 512^2 {t['swinir']:.2f}s, VAE encode "
              f"{t['vae_encode']:.2f}s (x2), decode {t['vae_decode']:.2f}s measured once; {args.steps} of 50 sampler "
              f"steps measured (median {step_s:.2f}s, 2 forwards each), image time extrapolated to 50 steps")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": "MPix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BSR pipeline 512x512, 50-step spaced sampler, cfg 4.0, random-init SD2.1 UNet+ControlNet (configs[1])"},
        "cpu_baseline": {"value": mpix, "unit": "MPix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mpix, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def kernel_census(pipe, torch, lib):
    """Records the tensor-core launches of one ControlNet+UNet forward (batch 2 = cond/uncond of a
    512^2 image), then replays each kernel family back-to-back inside a CUDA graph and times the
    replay with CUDA events on the launching stream (steady state, same buffers as the real forward).
    Returns per-family totals and the per-shape table (per-shape numbers from per-launch events)."""
    eng = pipe.cldm.engine
    dev = eng.dev
    x = torch.randn(2, 4, 64, 64, device=dev)
    ci = torch.randn(2, 4, 64, 64, device=dev) * 0.5
    eng.set_context(torch.randn(2, 77, 1024, device=dev))
    eng.set_timesteps([500], 2)
    eng.load_step(0)
    two = eng.two_streams
    eng.two_streams = False                  # record on one stream
    eng.forward(x, ci, [1.0] * 13)
    torch.cuda.synchronize()
    lib.record_begin()
    eng.forward(x, ci, [1.0] * 13)
    calls = lib.record_end()
    eng.two_streams = two
    fam, shapes = {}, {}
    names = {"gemm": "gemm+conv (gemm_tc_kernel)", "conv": "gemm+conv (gemm_tc_kernel)",
             "attention": "attention (attn_fwd_kernel)"}

    def replay_ms(sel, reps=5):
        for c in sel:
            c[3]()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for c in sel:
                c[3]()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for fname in set(names.values()):
        sel = [c for c in calls if names[c[0]] == fname]
        if sel:
            fam[fname] = [sum(c[2] for c in sel), replay_ms(sel), len(sel)]
    by_shape = {}
    for c in calls:
        by_shape.setdefault((c[0],) + tuple(c[1]), []).append(c)
    for k, sel in by_shape.items():      # per-shape: the launches of that shape back-to-back in a graph
        shapes[k] = [sum(c[2] for c in sel), replay_ms(sel, 3), len(sel)]
    return fam, shapes


def run_ours(args):
    import numpy as np
    import torch
    from diffbir_b200 import lib
    from diffbir_b200.utils.synth import RUN_DEFAULTS, build_synthetic_pipeline, synthetic_lq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lib.load()
    tiled = args.workload == "tiled2048"
    size = 2048 if tiled else 512
    kw = dict(RUN_DEFAULTS)
    if tiled:
        kw.update(cldm_tiled=True, cldm_tile_size=512, cldm_tile_stride=256)
    t_build = time.time()
    pipe = build_synthetic_pipeline(dev, seed=1234)
    # replicas restore different images; the tiled workload restores ONE image on all ranks
    lq = synthetic_lq(size, size, seed=0 if tiled else rank)
    lq_pinned = torch.from_numpy(lq).pin_memory()
    lq_dev = lq_pinned.to(dev)
    log(f"[rank {rank}] pipeline built in {time.time() - t_build:.1f}s")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one(device_resident: bool):
        torch.manual_seed(231)
        if device_resident:
            return pipe.run_device(lq_dev, **kw)
        return pipe.run(lq_pinned, **kw)

    for _ in range(max(args.warmup, 1)):
        one(True)
    one(False)
    barrier()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = lib.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.nvtx.range_push("dbir_timed")      # ncu --nvtx --nvtx-include "dbir_timed/" profiles exactly this region
    e0.record()
    for _ in range(args.steps):
        one(True)
    e1.record()
    torch.cuda.nvtx.range_pop()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = lib.launches() - n0
    # end-to-end through the public API (pinned host uint8 in, host uint8 out)
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        out = one(False)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - w0) * 1e3
    barrier()
    clk = clocks.stop() if rank == 0 else None

    times = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = times.tolist()
    images = (1 if tiled else world) * args.steps
    mpix_total = images * size * size / 1e6
    value = mpix_total / (dev_ms / 1e3)
    e2e = mpix_total / (e2e_ms / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel family (live CUDA events) --------------------------
    peak_tf, _, peak_src = measured_peaks()
    fam, shapes = kernel_census(pipe, torch, lib)
    gname = "gemm+conv (gemm_tc_kernel)"
    gf, gms, gn = fam[gname]
    achieved = gf / (gms * 1e-3) / 1e12
    forward_ms = sum(v[1] for v in fam.values())
    traffic = None
    tpath = ROOT / "profiles" / "r01_ncu_full_summary.json"
    if tpath.exists():       # dram bytes (read + write) per launch from the committed `ncu --set full` capture
        rows = [r for r in json.loads(tpath.read_text()) if "gemm_tc_kernel" in r["kernel"]]
        if rows:
            traffic = {"dram_bytes_per_launch_mean": sum(r["dram_bytes"] for r in rows) / len(rows),
                       "l2_to_sm_bytes_per_launch_mean": sum(r["l2_to_sm_bytes"] for r in rows) / len(rows),
                       "launches": len(rows), "source": "profiles/r01_ncu_full_summary.json"}
    roof = {"bound": "tensor", "kernel": gname, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved / peak_tf, "traffic": traffic, "peak_source": peak_src,
            "launches_per_forward": gn, "algorithmic_gflop_per_forward": gf / 1e9,
            "kernel_ms_per_forward": gms,
            "attention": {"achieved": fam.get("attention (attn_fwd_kernel)", [0, 1, 0])[0] /
                          (fam.get("attention (attn_fwd_kernel)", [0, 1, 0])[1] * 1e-3) / 1e12,
                          "ms_per_forward": fam.get("attention (attn_fwd_kernel)", [0, 0, 0])[1]},
            "tensor_kernel_ms_per_forward": forward_ms}
    prof_dir = ROOT / "gpurun_out"
    prof_dir.mkdir(exist_ok=True)
    with open(prof_dir / "kernel_census.csv", "w") as f:
        f.write("kind,shape,launches,gflop,ms,tflops\n")
        for k, (fl, ms, n) in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[0]},{'x'.join(map(str, k[1:]))},{n},{fl / 1e9:.2f},{ms:.4f},{fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:.1f}\n")
    # ---- CPU baseline (bounded sample) on the host cores, N = 1 only ------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        t, cores = cpu_reference_times(n_steps=1)
        step_s = t["sampler_step"][0]
        total = cpu_image_seconds(t, step_s)
        cpu = {"value": 512 * 512 / 1e6 / total, "unit": "MPix/s", "cores": cores, "kind": "port",
               "sample": (f"oracle port, fp32: SwinIR {t['swinir']:.2f}s + VAE encode 2x{t['vae_encode']:.2f}s + "
                          f"1 of 50 sampler steps ({step_s:.2f}s, 2 forwards) x50 + VAE decode {t['vae_decode']:.2f}s "
                          f"= {total:.1f}s per 512^2 image (extrapolated)")}
    line = {
        "metric": METRIC if not tiled else "MPix/s end-to-end 50-step restore, tiled 2048px",
        "value": value, "unit": "MPix/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if tiled else "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate" if lib.operand_dtype() == torch.float16 else "bf16 operands / f32 accumulate",
        "data": "synthetic",
        "config": {
            "workload": ("Tiled BSR 2048x2048, tile 512 stride 256 (49 latent tiles sharded over ranks, NCCL all-gather per step) (configs[3])"
                         if tiled else
                         "BSR pipeline 512x512, 50-step spaced sampler, cfg 4.0, random-init SD2.1 UNet+ControlNet (configs[1])"),
            "images_per_gpu_per_step": 1, "sampler_steps": SAMPLER_STEPS,
            "parallelism": "tiles sharded round-robin + all-gather" if tiled else "independent replicas, no collective",
            "l2": "no flush needed: each forward streams 2.6 GB of weights >> 126 MB L2",
        },
        "e2e": {"value": e2e, "unit": "MPix/s", "h2d_bytes_per_step": int(lq.nbytes), "d2h_bytes_per_step": int(out.nbytes),
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="512", choices=["512", "tiled2048"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
