"""Dev probe: time one SD1.5 UNet forward / forward+backward on the GPU kernels (not a bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sid_lsg_b200 as S

b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dtype = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else torch.float32
dev = "cuda"
t0 = time.time()
with torch.device(dev):
    m = S.UNet2DConditionModel(S.SD15, compute_dtype=dtype)
m.flatten_()
torch.cuda.synchronize()
print("build %.1fs, params %.1fM" % (time.time() - t0, m.flat.numel / 1e6))
x = torch.randn(b, 4, 64, 64, device=dev, requires_grad=True)
t = torch.randint(20, 980, (b,), device=dev)
e = torch.randn(b, 77, 768, device=dev)
def run(bwd):
    y = m(x, t, encoder_hidden_states=e).sample
    if bwd:
        y.backward(torch.ones_like(y))
    return y
for bwd in (False, True):
    run(bwd); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = S.lib.launches
    ev0.record(); y = run(bwd); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    fl = 0.8033e12 * b * (3 if bwd else 1)
    import ctypes
    cnt = (ctypes.c_long * 2)()
    S.lib.call("counters", ctypes.addressof(cnt))
    print("tc launches %d simt launches %d" % (cnt[0], cnt[1]))
    print("b=%d %s %s: %.1f ms  %.1f TFLOP/s  calls %d  finite %s  mem %.1f GB" % (b, dtype, "fwd+bwd" if bwd else "fwd", ms, fl / ms / 1e9, S.lib.launches - n0, bool(torch.isfinite(y).all()), torch.cuda.max_memory_allocated() / 1e9))
