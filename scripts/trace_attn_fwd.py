"""In-kernel phase trace of the two-tile attention forward (CTA (1,0,0), B8 N4096 d40): per key tile clock deltas."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sid_lsg_b200._lib import lib, ptr, stream

B, N, C, H = 8, 4096, 320, 8
d = C // H
dev = "cuda"
q, k, v = (torch.randn(B, N, C, device=dev, dtype=torch.bfloat16) for _ in range(3))
o = torch.empty_like(q)
lse = torch.empty(B, H, N, device=dev, dtype=torch.float32)
trace = torch.zeros(2 * 32 * 16, device=dev, dtype=torch.int64)
for _ in range(2):
    lib.call("debug_attention_fwd_trace", ptr(q), ptr(k), ptr(v), ptr(o), ptr(lse), B, N, N, H, d, C, C, C, ptr(trace), stream())
torch.cuda.synchronize()
t = trace.cpu()[:512].view(32, 16)
td = trace.cpu()[512:].view(32, 16)
names = {0: "A:top", 1: "A:s_full", 2: "A:exp_done", 3: "A:w1_done", 15: "A:w3_done", 4: "B:top", 5: "B:s_full", 6: "B:exp_done",
         14: "m:top", 8: "m:waitPA", 9: "m:PA", 7: "m:QKA_iss", 10: "m:issuedA", 11: "m:waitPB", 12: "m:PB", 13: "m:issuedB"}
base = int(t[2, 0])
cols = [0, 1, 2, 3, 15, 4, 5, 6, 14, 8, 9, 7, 10, 11, 12, 13]
print("tile " + " ".join("%11s" % names[c] for c in cols))
for j in range(2, 14):
    print("%4d " % j + " ".join("%11d" % (int(t[j, c]) - base) for c in cols))
print("period (A:s_full):", [int(t[j + 1, 1]) - int(t[j, 1]) for j in range(2, 13)])
print("A: wait S / exps :", [(int(t[j, 1]) - int(t[j, 0]), int(t[j, 2]) - int(t[j, 1])) for j in range(2, 10)])
print("B: wait S / exps :", [(int(t[j, 5]) - int(t[j, 4]), int(t[j, 6]) - int(t[j, 5])) for j in range(2, 10)])

if int(td.abs().sum()) == 0:
    sys.exit(0)
print("per softmax warp (A: 0-3, B: 4-7): scores seen / exponentials done, relative to warp 0's scores-seen of the tile")
for j in range(3, 12):
    b0 = int(td[j, 8])
    print("%4d  seen " % j + " ".join("%6d" % (int(td[j, 8 + w]) - b0) for w in range(8)) +
          "   done " + " ".join("%6d" % (int(td[j, w]) - b0) for w in range(8)))
