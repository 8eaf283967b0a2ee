"""Does one whole SiD-LSG iteration capture into a CUDA graph, and what does replay buy over eager launches?"""
import sys, os, copy, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sid_lsg_b200 as S
from sid_lsg_b200.training.step import synth_microbatch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = S.SD15
torch.manual_seed(0)
with torch.device(dev):
    base = S.UNet2DConditionModel(cfg, compute_dtype=torch.bfloat16)
base.flatten_()
fake, G, G_ema = copy.deepcopy(base), copy.deepcopy(base), copy.deepcopy(base)
st = S.SiDLSGStep(base, fake, G, G_ema, S.DDPMScheduler(device=dev), cfg_train_fake=1.5, cfg_eval_fake=1.5, cfg_eval_real=1.5)
mf = [synth_microbatch(B, cfg, 1, dev, dropout=True)]
mg = [synth_microbatch(B, cfg, 2, dev)]


def timed(fn, n=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host * 1e3


for _ in range(2):
    st.iteration(mf, mg, batch_size=B)
ms, host = timed(lambda: st.iteration(mf, mg, batch_size=B))
print("eager : %.1f ms per iteration on the device, %.1f ms of host enqueue" % (ms, host), flush=True)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    st.iteration(mf, mg, batch_size=B)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
torch.cuda.empty_cache()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = st.iteration(mf, mg, batch_size=B)
torch.cuda.synchronize()
print("captured; peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9), flush=True)
ms, host = timed(lambda: g.replay())
print("graph : %.1f ms per iteration on the device, %.1f ms of host enqueue" % (ms, host))
print("losses", [float(x[0]) for x in out])
