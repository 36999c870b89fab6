"""Per-phase clock64 timeline of one attention-backward CTA (key tile 1 of head 0) at the benchmark shape."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from harness import ROOT  # noqa: F401,E402
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
for _ in range(2):
    lib.attn_bwd(Q, K, V, dO, lse, delta, dQ, dK, dV)
dbg = torch.zeros(19 * 16, device="cuda", dtype=torch.int64)
lib._lib.qfx_attn_bwd_set_debug.argtypes = [C.c_void_p]
lib._lib.qfx_attn_bwd_set_debug(C.c_void_p(dbg.data_ptr()))
lib.attn_bwd(Q, K, V, dO, lse, delta, dQ, dK, dV)
torch.cuda.synchronize()
lib._lib.qfx_attn_bwd_set_debug(C.c_void_p(0))
d = dbg.view(19, 16).cpu()
t0 = int(d[0, 0])
names = {7: "cmp:dQ_drained", 0: "mma:loads_ready", 1: "mma:Scols_free", 2: "mma:P_ready", 3: "mma:dS_ready", 8: "cmp:iter_start", 9: "cmp:S_ready", 10: "cmp:stage_free",
         11: "cmp:P_written", 12: "cmp:dS_written", 13: "cmp:dQ_ready", 14: "cmp:reduce_issued", 15: "cmp:reduce_read_done"}
print("iter " + " ".join(f"{names[k]:>20s}" for k in sorted(names)))
for i in range(19):
    print(f"{i:4d} " + " ".join(f"{int(d[i, k]) - t0:20d}" for k in sorted(names)))
print("per-iteration period (mma:loads_ready):", [int(d[i + 1, 0] - d[i, 0]) for i in range(18)])
