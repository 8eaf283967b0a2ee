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
delta = torch.empty((2,) + tuple(lse.shape), device=dev, dtype=torch.float32)
dq_acc = torch.empty(B, N, C, device=dev, dtype=torch.float32)
dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
trace = torch.zeros(2 * 32 * 16, device=dev, dtype=torch.int64)
for _ in range(2):
    lib.call("debug_attention_bwd_trace", ptr(q), ptr(k), ptr(v), ptr(o), ptr(do), ptr(lse), ptr(delta), ptr(dq_acc),
             ptr(dq), ptr(dk), ptr(dv), B, N, N, H, d, C, C, C, C, C, C, ptr(trace), stream())
torch.cuda.synchronize()
t = trace.cpu()[:512].view(32, 16)
td = trace.cpu()[512:].view(32, 16)
names = {0: "c:top", 1: "c:s_full", 2: "c:P_done", 3: "c:dp_full", 4: "c:dS_done", 6: "d:dq_full", 7: "d:reduced",
         8: "m:top", 9: "m:p_full", 10: "m:dV,S+", 11: "m:ds_full", 12: "m:dK", 13: "m:dq_empty", 14: "m:dQ,dP+"}
cols = sorted(names)
base = int(t[2, 0])
print("pipelined kernel (attn_bwd2_kernel); SIDLSG_ATTN_BWD2=0 traces the first kernel with different slot meanings")
print("tile " + " ".join("%10s" % names[c] for c in cols))
for i in range(2, 14):
    print("%4d " % i + " ".join("%10d" % (int(t[i, c]) - base) for c in cols))
print("period (c:s_full) per tile:", [int(t[i + 1, 1]) - int(t[i, 1]) for i in range(2, 13)])
print("compute: wait S, exps+P, wait dP, dS:", [(int(t[i, 1]) - int(t[i, 0]), int(t[i, 2]) - int(t[i, 1]), int(t[i, 3]) - int(t[i, 2]), int(t[i, 4]) - int(t[i, 3])) for i in range(2, 8)])

if int(td.abs().sum()):
    print("per compute warp 0-7: P^T done / dS^T done, relative to warp 0's P^T done of the tile")
    for i in range(3, 12):
        b0 = int(td[i, 0])
        print("%4d  P " % i + " ".join("%6d" % (int(td[i, w]) - b0) for w in range(8)) +
              "   dS " + " ".join("%6d" % (int(td[i, 8 + w]) - b0) for w in range(8)))
