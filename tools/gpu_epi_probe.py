import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from diffbir_b200 import lib
dt = lib.operand_dtype()
M, N, K = 8192, 320, 320
a = torch.randn(M, K, device="cuda").to(dt); b = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
out = torch.empty(M, N, device="cuda"); r = torch.randn(M, N, device="cuda"); bias = torch.randn(N, device="cuda")
for name, fl in [("full", 0), ("no stores", 1), ("no bias", 2), ("no residual ld", 4), ("no res+no stores", 5), ("no bias/res/stores", 7), ("skip transposed phase", 8), ("skip all", 15)]:
    dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
    for i in range(3):
        lib.gemm(a, b, out, M=M, N=N, K=K, bias=bias, residual=r, debug_stamps=dbg, debug_flags=fl)
    torch.cuda.synchronize()
    d = dbg.view(-1, 8).cpu(); d = d[d[:, 3] > 0].double()
    print(f"{name:24s}: epilogue {(d[:,3]-d[:,2]).mean():8.0f} cycles | chunk0 finish {(d[:,6]-d[:,5]).mean():7.0f}")
