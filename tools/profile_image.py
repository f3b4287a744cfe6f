"""One steady-state 512^2 image between cudaProfilerStart/Stop (GPU box), for

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/image_launches.csv python tools/profile_image.py [sampler_steps]

Three warm-up images first (GEMM plans tuned, every stage graph captured), so the profiled launches
are exactly the graph replays a timed bench step consists of."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402
from diffbir_b200.utils.synth import RUN_DEFAULTS, build_synthetic_pipeline, synthetic_lq  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    lib.load()
    pipe = build_synthetic_pipeline("cuda", 1234)
    lq = torch.from_numpy(synthetic_lq(512, 512)).cuda()
    kw = dict(RUN_DEFAULTS, steps=steps)
    for _ in range(3):
        torch.manual_seed(231)
        pipe.run_device(lq, **kw)
    torch.cuda.synchronize()
    n0 = lib.launches()
    torch.cuda.profiler.start()
    torch.manual_seed(231)
    pipe.run_device(lq, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"profiled launches through the library: {lib.launches() - n0}")


if __name__ == "__main__":
    main()
