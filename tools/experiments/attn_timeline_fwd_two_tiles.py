"""clock64 timeline of CTA 0 of the two-tile attention forward (attention_fwd2.cu) at the benchmark shape."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from harness import ROOT, time_cuda  # noqa: F401,E402
import torch  # noqa: E402
from qflux_b200 import lib  # noqa: E402

B, H, S, T = 4, 24, 2400, 352
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()
Q, K, V = (mk(B, H, S, 128) for _ in range(3))
lse = torch.zeros(B, H, S, device="cuda")
ot, oi = torch.zeros(B * T, H * 128, device="cuda", dtype=torch.bfloat16), torch.zeros(B * (S - T), H * 128, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
ms = time_cuda(lambda: lib.attn_fwd(Q, K, V, ot, oi, T, lse), flush=flush)
print(f"attn_fwd: {ms:.4f} ms ({4.0 * B * H * S * S * 128 / ms / 1e9:.1f} TFLOP/s) env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("QFX_")))
dbg = torch.zeros(40 * 16, device="cuda", dtype=torch.int64)
lib._lib.qfx_attn_bwd_set_debug.argtypes = [C.c_void_p]
lib._lib.qfx_attn_bwd_set_debug(C.c_void_p(dbg.data_ptr()))
lib.attn_fwd(Q, K, V, ot, oi, T, lse)
torch.cuda.synchronize()
lib._lib.qfx_attn_bwd_set_debug(C.c_void_p(0))
d = dbg.view(40, 16).cpu()
t0 = int(d[0, 0])
cols = [(0, "mma:S0 s"), (1, "mma:S0 e"), (2, "mma:PV1 e"), (3, "mma:S1 e"), (4, "mma:PV0 e"), (8, "sm0:S rdy"), (9, "sm0:P done"), (12, "sm1:S rdy"), (13, "sm1:P done")]
print("tile " + " ".join(f"{n:>11s}" for _, n in cols))
for j in range(19):
    print(f"{j:4d} " + " ".join(f"{int(d[j, k]) - t0:11d}" for k, _ in cols))
print("period:", [int(d[j + 1, 0] - d[j, 0]) for j in range(18)])
