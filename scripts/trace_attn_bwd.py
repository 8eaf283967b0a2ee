"""In-kernel phase trace of the attention backward (CTA (1,0,0), B8 N4096 d40): prints per-tile clock deltas."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sid_lsg_b200 import ops
from sid_lsg_b200._lib import lib, ptr, stream

B, N, C, H = 8, 4096, 320, 8
d = C // H
dev = "cuda"
q, k, v = (torch.randn(B, N, C, device=dev, dtype=torch.bfloat16) for _ in range(3))
o = torch.empty_like(q)
lse = torch.empty(B, H, N, device=dev, dtype=torch.float32)
lib.call("attention_fwd", ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), B, N, N, H, d, C, C, C, stream())
do = torch.randn_like(o)
delta = torch.empty_like(lse)
dq_acc = torch.empty(B, N, C, device=dev, dtype=torch.float32)
dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
trace = torch.zeros(32 * 16, device=dev, dtype=torch.int64)
for _ in range(2):
    lib.call("debug_attention_bwd_trace", ptr(q), ptr(k), ptr(v), ptr(o), ptr(do), ptr(lse), ptr(delta), ptr(dq_acc),
             ptr(dq), ptr(dk), ptr(dv), B, N, N, H, d, C, C, C, C, C, C, ptr(trace), stream())
torch.cuda.synchronize()
t = trace.cpu().view(32, 16)
names = ["c:loop_top", "c:s_full", "c:dq_full(i-1)", "c:exp_done", "c:drain_done", "c:P_stored", "c:dp_full", "c:ds_done",
         "m:wait_pds", "m:pds_full", "m:issued"]
base = int(t[2, 0])
print("tile " + " ".join("%14s" % n for n in names))
for i in range(2, 14):
    print("%4d " % i + " ".join("%14d" % (int(t[i, j]) - base) for j in range(len(names))))
print("period (c:s_full) per tile:", [int(t[i + 1, 1]) - int(t[i, 1]) for i in range(2, 13)])
