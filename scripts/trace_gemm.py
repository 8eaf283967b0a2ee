"""In-kernel phase trace of the tensor-core GEMM (CTA 1, first tiles): linear forward M x K -> N [+ bias + residual].
Needs an instrumented build: SIDLSG_NVCC_EXTRA=-DSIDLSG_GEMM_TRACE python -m sid_lsg_b200.build --force"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sid_lsg_b200 import ops
from sid_lsg_b200._lib import lib, ptr

dev = "cuda"
M, K, N = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (262144, 320, 320)))
res = len(sys.argv) > 4 and sys.argv[4] == "res"
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
w = torch.nn.Parameter((torch.randn(N, K) / math.sqrt(K)).to(dev)); w._shadow = w.detach().to(torch.bfloat16)
b = torch.nn.Parameter(torch.randn(N, device=dev)) if res else None
r = torch.randn(M, N, device=dev, dtype=torch.bfloat16) if res else None
for _ in range(3):
    ops.linear(x, w, b, r)
trace = torch.zeros(32 * 16, device=dev, dtype=torch.int64)
lib.call("debug_gemm_trace", ptr(trace))
ops.linear(x, w, b, r)
torch.cuda.synchronize()
lib.call("debug_gemm_trace", None)
t = trace.cpu().view(32, 16)
names = ["e:top", "e:tfull", "c0:buf", "c0:conv", "c0:store", "c1:buf", "c1:conv", "c1:store", "c2:buf", "c2:conv", "c2:store",
         "e:end", "m:tempty", "m:full0", "m:issued", "p:start"]
base = int(t[2, 0])
print("M %d K %d N %d%s   (clk relative to tile 2's epilogue top; 0 = not stamped)" % (M, K, N, " +bias+res" if res else ""))
print("tile " + " ".join("%9s" % n for n in names))
for i in range(2, 12):
    print("%4d " % i + " ".join("%9d" % ((int(t[i, c]) - base) if int(t[i, c]) else 0) for c in range(16)))
print("period (e:end):", [int(t[i + 1, 11]) - int(t[i, 11]) for i in range(2, 11)])
print("epilogue: wait tfull, chunks, total:", [(int(t[i, 1]) - int(t[i, 0]), int(t[i, 11]) - int(t[i, 1]), int(t[i, 11]) - int(t[i, 0])) for i in range(2, 10)])
print("mma: tempty->full0, full0->issued:", [(int(t[i, 13]) - int(t[i, 12]), int(t[i, 14]) - int(t[i, 13])) for i in range(2, 10)])
