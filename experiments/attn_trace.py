"""Phase timeline of one attention-forward CTA (clock64 stamps written by the kernel, sidlsg_attention_trace)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sid_lsg_b200 import ops
from sid_lsg_b200._lib import lib

B, N, C, h = 8, 4096, 320, 8
q = torch.randn(B, N, C, device="cuda", dtype=torch.bfloat16)
k = torch.randn_like(q); v = torch.randn_like(q)
for _ in range(2):
    ops.attention(q, k, v, h)
buf = torch.zeros(32 * 16, dtype=torch.int64, device="cuda")
lib.call("attention_trace", buf.data_ptr())
ops.attention(q, k, v, h)
torch.cuda.synchronize()
lib.call("attention_trace", None)
t = buf.cpu().view(32, 16)
t0 = int(t[0, 0])
print("tile | softmax: waitS  waitO   exp  fence arrive | period || mma: waitP  issuePV issueQK | (cycles)")
for j in range(1, 31):
    s = [int(x) - t0 for x in t[j, :6]]
    m = [int(x) - t0 for x in t[j, 8:12]]
    per = int(t[j + 1, 0] - t[j, 0])
    print("%4d | %6d %6d %6d %6d %6d | %6d || %6d %6d %6d | start %d" % (
        j, s[1] - s[0], s[2] - s[1], s[3] - s[2], s[4] - s[3], s[5] - s[4], per, m[1] - m[0], m[2] - m[1], m[3] - m[2], s[0]))
