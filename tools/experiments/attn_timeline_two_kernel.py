"""Per-phase clock64 timeline of one dK/dV CTA of the two-kernel attention backward (QFX_ATTN_BWD2=1) at the benchmark shape."""
import ctypes as C
import os
import sys

os.environ["QFX_ATTN_BWD2"] = "1"
os.environ.setdefault("QFX_ATTN_BWD2_ONLY", "dkv")
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
NQ = 38
dbg = torch.zeros(NQ * 16, device="cuda", dtype=torch.int64)
lib._lib.qfx_attn_bwd_set_debug.argtypes = [C.c_void_p]
lib._lib.qfx_attn_bwd_set_debug(C.c_void_p(dbg.data_ptr()))
lib.attn_bwd(Q, K, V, dO, lse, delta, dQ, dK, dV)
torch.cuda.synchronize()
lib._lib.qfx_attn_bwd_set_debug(C.c_void_p(0))
d = dbg.view(NQ, 16).cpu()
t0 = int(d[0, 0])
names = {0: "mma:q_ready(S i)", 1: "mma:P_ready", 2: "mma:dS_ready", 4: "cmp:start", 5: "cmp:LD_ready", 6: "cmp:S_ready", 7: "cmp:P_computed",
         8: "cmp:dv_done(i-1)", 9: "cmp:P_stored", 10: "cmp:dS_computed", 11: "cmp:dk_done(i-1)", 12: "cmp:end"}
print("iter " + " ".join(f"{names[k]:>17s}" for k in sorted(names)))
for i in range(4, 14):
    print(f"{i:4d} " + " ".join(f"{int(d[i, k]) - t0:17d}" for k in sorted(names)))
print("period (cmp:end):", [int(d[i + 1, 12] - d[i, 12]) for i in range(NQ - 2)])
