"""Per-phase clock64 timeline of one pipelined attention-backward CTA (attention_bwd3.cu, key tile 1 of head 0) at the benchmark shape.
Columns are clocks since the first stamp; `mma:*` are the issuing warp (s = after its wait, e = after the last issue of the group)."""
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
Q, K, V, dO = (mk(B, H, S, 128) for _ in range(4))
lse = torch.zeros(B, H, S, device="cuda")
ot, oi = torch.zeros(B * T, H * 128, device="cuda", dtype=torch.bfloat16), torch.zeros(B * (S - T), H * 128, device="cuda", dtype=torch.bfloat16)
lib.attn_fwd(Q, K, V, ot, oi, T, lse)
delta = torch.zeros(B, H, S, device="cuda")
dQ = torch.zeros(B, H, S, 128, device="cuda")
dK, dV = torch.empty_like(K), torch.empty_like(V)


def run():
    dQ.zero_()
    lib.attn_bwd(Q, K, V, dO, lse, delta, dQ, dK, dV)


flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
ms = time_cuda(run, flush=flush)
print(f"attn_bwd: {ms:.4f} ms  ({10.0 * B * H * S * S * 128 / ms / 1e9:.1f} TFLOP/s)  env: "
      + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("QFX_")))
dbg = torch.zeros(19 * 16, device="cuda", dtype=torch.int64)
lib._lib.qfx_attn_bwd_set_debug.argtypes = [C.c_void_p]
lib._lib.qfx_attn_bwd_set_debug(C.c_void_p(dbg.data_ptr()))
run()
torch.cuda.synchronize()
lib._lib.qfx_attn_bwd_set_debug(C.c_void_p(0))
d = dbg.view(19, 16).cpu()
t0 = int(d[0, 0])
names = ["mma:S s", "mma:S e", "mma:dQdK s", "mma:dQdK e", "mma:dP s", "mma:dP e", "mma:dV s", "mma:dV e", "cmp:S rdy", "cmp:P done", "cmp:dP rdy",
         "cmp:dS done", "drn:dQ rdy", "drn:drained", "cmp:S in reg", "cmp:dP in reg"]
print("iter " + " ".join(f"{n:>12s}" for n in names))
for i in range(19):
    print(f"{i:4d} " + " ".join(f"{int(d[i, k]) - t0:12d}" for k in range(16)))
print("period (mma:S s):", [int(d[i + 1, 0] - d[i, 0]) for i in range(18)])
