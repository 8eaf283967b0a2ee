"""Dev micro-benchmarks of single kernels through the C ABI (CUDA-event timing; also the ncu capture target)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sid_lsg_b200 as S
from sid_lsg_b200 import ops

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5


def timeit(fn, flops=None, nbytes=None, name=""):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    s = "%-44s %8.3f ms" % (name, ms)
    if flops:
        s += "  %7.1f TFLOP/s" % (flops / ms / 1e9)
    if nbytes:
        s += "  %7.1f GB/s" % (nbytes / ms / 1e6)
    print(s, flush=True)


def bfp(t, cl=False):
    t = t.to(dev)
    if cl:
        t = t.contiguous(memory_format=torch.channels_last)
    p = torch.nn.Parameter(t)
    p.grad = torch.zeros_like(p)
    p._shadow = p.detach().to(torch.bfloat16)
    return p


if which in ("all", "attn"):
    for (B, N, M, C, h) in ((8, 4096, 4096, 320, 8), (8, 1024, 1024, 640, 8), (8, 4096, 77, 320, 8), (8, 1024, 1024, 640, 10)):
        q = torch.randn(B, N, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        k = torch.randn(B, M, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        v = torch.randn(B, M, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        fl = 4.0 * B * N * M * C
        timeit(lambda: ops.attention(q.detach(), k.detach(), v.detach(), h), fl, name="attn fwd B%d N%d M%d C%d h%d" % (B, N, M, C, h))
        o = ops.attention(q, k, v, h)
        do = torch.randn_like(o)
        timeit(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True), 2.5 * fl, name="attn bwd")

if which in ("all", "gemm"):
    for (M, K, N) in ((131072, 320, 320), (131072, 320, 2560), (131072, 1280, 320), (32768, 640, 640), (32768, 640, 5120),
                      (8192, 1280, 1280), (8192, 1280, 10240), (8192, 5120, 1280)):
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16, requires_grad=True)
        w = bfp(torch.randn(N, K) / math.sqrt(K))
        fl = 2.0 * M * N * K
        timeit(lambda: ops.linear(x.detach(), w), fl, nbytes=2.0 * (M * K + M * N + N * K), name="linear fwd M%d K%d N%d" % (M, K, N))
        y = ops.linear(x, w)
        dy = torch.randn_like(y)
        timeit(lambda: torch.autograd.grad(y, (x,), dy, retain_graph=True), fl, name="  dgrad")
        w.requires_grad_(True)
        x2 = x.detach()
        y2 = ops.linear(x2, w)
        timeit(lambda: y2.backward(dy, retain_graph=True), fl, name="  wgrad")

if which in ("all", "conv"):
    for (B, H, C, N) in ((32, 64, 320, 320), (32, 64, 640, 320), (32, 32, 640, 640), (32, 16, 1280, 1280), (32, 8, 2560, 1280)):
        x = torch.randn(B, H, H, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        w = bfp(torch.randn(N, C, 3, 3) / math.sqrt(9 * C), cl=True)
        b = torch.nn.Parameter(torch.randn(N, device=dev))
        fl = 2.0 * B * H * H * N * 9 * C
        timeit(lambda: ops.conv3x3(x.detach(), w, b), fl, name="conv fwd B%d H%d C%d N%d" % (B, H, C, N))
        w.requires_grad_(False)
        y = ops.conv3x3(x, w, b)
        dy = torch.randn_like(y)
        timeit(lambda: torch.autograd.grad(y, (x,), dy, retain_graph=True), fl, name="  dgrad")
        w.requires_grad_(True)
        y2 = ops.conv3x3(x.detach(), w, None)
        timeit(lambda: y2.backward(dy, retain_graph=True), fl, name="  wgrad")

if which in ("all", "gn"):
    for (B, HW, C) in ((64, 4096, 320), (64, 1024, 640), (64, 256, 1280), (64, 4096, 960), (32, 4096, 320)):
        x = torch.randn(B, HW, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        g = torch.nn.Parameter(torch.ones(C, device=dev)); bt = torch.nn.Parameter(torch.zeros(C, device=dev))
        n = B * HW * C * 2.0
        timeit(lambda: ops.group_norm(x.detach(), g, bt, 32, 1e-5, True), nbytes=3 * n, name="gn fwd B%d HW%d C%d" % (B, HW, C))
        y = ops.group_norm(x, g, bt, 32, 1e-5, True)
        dy = torch.randn_like(y)
        timeit(lambda: torch.autograd.grad(y, (x,), dy, retain_graph=True), nbytes=5 * n, name="  gn bwd")

if which in ("all", "ln"):
    for (M, C) in ((262144, 320), (65536, 640), (16384, 1280)):
        x = torch.randn(M, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
        g = torch.nn.Parameter(torch.ones(C, device=dev)); bt = torch.nn.Parameter(torch.zeros(C, device=dev))
        g.grad = torch.zeros_like(g); bt.grad = torch.zeros_like(bt)
        n = M * C * 2.0
        timeit(lambda: ops.layer_norm(x.detach(), g, bt, 1e-5), nbytes=2 * n, name="ln fwd M%d C%d" % (M, C))
        y = ops.layer_norm(x, g, bt, 1e-5)
        dy = torch.randn_like(y)
        timeit(lambda: torch.autograd.grad(y, (x,), dy, retain_graph=True), nbytes=3 * n, name="  ln bwd")
    for (M, F) in ((262144, 1280), (65536, 2560), (16384, 5120)):
        x = torch.randn(M, 2 * F, device=dev, dtype=torch.bfloat16, requires_grad=True)
        n = M * F * 2.0
        timeit(lambda: ops.GegluFn.apply(x.detach()), nbytes=3 * n, name="geglu fwd M%d F%d" % (M, F))
        y = ops.GegluFn.apply(x)
        dy = torch.randn_like(y)
        timeit(lambda: torch.autograd.grad(y, (x,), dy, retain_graph=True), nbytes=5 * n, name="  geglu bwd")
