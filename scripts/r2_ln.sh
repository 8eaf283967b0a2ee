#!/bin/bash
OUT=gpurun_out/r2s; mkdir -p $OUT
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_step_gpu.py -m gpu -q -x -p no:cacheprovider -k "layernorm or unet or iteration" > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 python - <<'PY' 2>&1 | tee $OUT/micro_ln.txt
import torch, sys
sys.path.insert(0, '.')
from sid_lsg_b200 import ops
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
for rows, C in ((64*4096, 320), (64*1024, 640), (64*256, 1280)):
    x = torch.randn(rows, C, device='cuda', dtype=torch.bfloat16, requires_grad=True)
    g = torch.nn.Parameter(torch.ones(C, device='cuda')); b = torch.nn.Parameter(torch.zeros(C, device='cuda'))
    g.grad = torch.zeros_like(g); b.grad = torch.zeros_like(b)
    y = ops.layer_norm(x, g, b)
    dy = torch.randn_like(y)
    f = t(lambda: ops.layer_norm(x.detach(), g, b))
    bw = t(lambda: torch.autograd.grad(y, x, dy, retain_graph=True))
    nb = rows*C*2
    print("LN rows %d C %d: fwd %.3f ms %.0f GB/s   bwd %.3f ms %.0f GB/s" % (rows, C, f, 2*nb/f/1e6, bw, 3*nb/bw/1e6))
PY
