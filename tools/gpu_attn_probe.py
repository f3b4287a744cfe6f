"""(needs the probe build: python -m diffbir_b200.build --tag=probes -DDBIR_DEBUG_PROBES -DDBIR_ATTN_PROBE;
run with DBIR_LIB_TAG=probes)
Where an attention CTA's time goes (clock64 sums per CTA: softmax-warp waits vs work, MMA-warp waits)."""
import ctypes as C
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from diffbir_b200 import lib  # noqa: E402

L = lib.load()
dt = lib.operand_dtype()
for b, h, sq, skv, use_ws in [(2, 5, 4096, 4096, True), (2, 5, 4096, 4096, False), (1, 1, 128, 4096, False), (2, 10, 1024, 1024, False)]:
    c = h * 64
    qkv = torch.randn(b * sq, 3 * c, device="cuda").to(dt)
    k = torch.randn(b * skv, 2 * c, device="cuda").to(dt)
    out = torch.empty(b * sq, c, device="cuda", dtype=dt)
    nws = lib.attention_ws_bytes(b, h, sq, skv) if use_ws else 0
    ws = torch.zeros(nws // 4, device="cuda") if nws else None
    dbg = torch.zeros(1024 * 8, dtype=torch.int64, device="cuda")
    for i in range(3):
        if i == 2:
            L.dbir_debug_attn_stamps(C.c_void_p(dbg.data_ptr()))
        lib.attention(qkv, k, k[:, c:], out, batch=b, heads=h, sq=sq, skv=skv, ldq=3 * c, ldk=2 * c, ldv=2 * c, ldo=c, ws=ws)
    torch.cuda.synchronize()
    L.dbir_debug_attn_stamps(C.c_void_p(0))
    d = dbg.view(-1, 8).cpu().double()
    d = d[d[:, 5] > 0]
    n = d[:, 5]
    print(f"b={b} h={h} sq={sq} skv={skv} ws={use_ws}: ctas {len(d)} steps/cta {n.mean():.1f} | softmax warp per step: total {(d[:,0]/n).mean():.0f} "
          f"= wait S {(d[:,1]/n).mean():.0f} + ld/max/rescale {(d[:,2]/n).mean():.0f} + wait PV(j-2) {(d[:,3]/n).mean():.0f} + exp/store {(d[:,4]/n).mean():.0f}"
          f" | MMA warp wait P per step {(d[:,6]/n).mean():.0f} cycles")
