"""ncu capture target: each hot kernel launched ONCE at its SD1.5 batch-32 shape through the C ABI
(ncu --set full -k regex:'gemm_tc|attn_|gn_' python scripts/prof_target.py)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sid_lsg_b200 import ops

dev = "cuda"
what = sys.argv[1] if len(sys.argv) > 1 else "attn,gemm,conv,gn"


def bfp(t, cl=False):
    t = t.to(dev)
    if cl:
        t = t.contiguous(memory_format=torch.channels_last)
    p = torch.nn.Parameter(t)
    p.grad = torch.zeros_like(p)
    p._shadow = p.detach().to(torch.bfloat16)
    return p


if "attn" in what:
    B, N, C, h = 8, 4096, 320, 8
    a = torch.randn(B, N, 3 * C, device=dev, dtype=torch.bfloat16, requires_grad=True)
    o = ops.packed_attention(a, None, h)
    o.backward(torch.randn_like(o))
    B, N, C, h = 8, 1024, 640, 8
    a = torch.randn(B, N, 3 * C, device=dev, dtype=torch.bfloat16, requires_grad=True)
    o = ops.packed_attention(a, None, h)
    o.backward(torch.randn_like(o))
if "gemm" in what:
    for (M, K, N, res) in ((131072, 320, 320, True), (131072, 320, 960, False), (131072, 320, 2560, False),
                           (131072, 1280, 320, True), (32768, 640, 640, True)):
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16, requires_grad=True)
        w = bfp(torch.randn(N, K) / math.sqrt(K))
        b = torch.nn.Parameter(torch.randn(N, device=dev)); b.grad = torch.zeros_like(b)
        r = torch.randn(M, N, device=dev, dtype=torch.bfloat16) if res else None
        y = ops.linear(x, w, b, r)           # fwd
        y.backward(torch.randn_like(y))      # dgrad + wgrad
if "conv" in what:
    for (B, H, C, N) in ((32, 64, 320, 320), (32, 16, 1280, 1280)):
        x = torch.randn(B, H, H, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        w = bfp(torch.randn(N, C, 3, 3) / math.sqrt(9 * C), cl=True)
        b = torch.nn.Parameter(torch.randn(N, device=dev)); b.grad = torch.zeros_like(b)
        y = ops.conv3x3(x, w, b)
        y.backward(torch.randn_like(y))
if "gn" in what:
    for (B, HW, C) in ((32, 4096, 320), (32, 1024, 640)):
        x = torch.randn(B, HW, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        g = torch.nn.Parameter(torch.ones(C, device=dev)); bt = torch.nn.Parameter(torch.zeros(C, device=dev))
        y = ops.group_norm(x, g, bt, 32, 1e-5, True)
        y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("done")
