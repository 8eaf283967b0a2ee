"""Reference-surface host logic on CPU: sampler order, RNG stream, stats, dnnlib helpers, parameter order, pickling,
gloo process groups (no GPU, no kernels)."""
import os
import pickle
import sys

import pytest
import torch

from golden_util import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_infinite_sampler_matches_reference_order():
    """tests/golden/sampler_order.pt = the reference's own InfiniteSampler (torch_utils/misc.py:110-141)."""
    from sid_lsg_b200.torch_utils import misc
    fx = load("sampler_order.pt")
    assert len(fx) == 4
    for key, want in fx.items():
        r, w = key[1:].split("w")
        it = iter(misc.InfiniteSampler(list(range(37)), rank=int(r), num_replicas=int(w), seed=3))
        got = torch.tensor([int(next(it)) for _ in range(len(want))])
        assert torch.equal(got, want), key
    # ranks partition the global stream
    its = [iter(misc.InfiniteSampler(list(range(10)), rank=r, num_replicas=3, seed=1)) for r in range(3)]
    one = iter(misc.InfiniteSampler(list(range(10)), rank=0, num_replicas=1, seed=1))
    merged = [int(next(its[k % 3])) for k in range(30)]
    assert merged == [int(next(one)) for _ in range(30)]
    with pytest.raises(AssertionError):
        misc.InfiniteSampler([], rank=0)
    with pytest.raises(AssertionError):
        misc.InfiniteSampler([1], rank=2, num_replicas=2)


@pytest.mark.parametrize("name", ["loop_1step.pt", "loop_2step_alpha12.pt"])
def test_draw_stream_reproduces_reference_draws_from_seed(name):
    """Every z / noise / timestep / dropout draw the REFERENCE loop made (recorded in the fixture) comes out of
    DrawStream(seed) when it is walked in the loop's order with the scheduler's discarded draws burned."""
    from sid_lsg_b200.training.draws import DrawStream
    fx = load(name)
    b, n = fx["batch_gpu"], fx["num_steps"]
    rounds = fx["batch"] // b
    shape = (b, 4, 16, 16)
    d = DrawStream(seed=3, rank=0, world=1, device="cpu", rng_device="cpu", compat=True)
    torch.empty((), dtype=torch.int64).random_(generator=d.cpu_gen)   # DataLoader base seed (iterator creation)
    d.burn(shape, 14)                                                   # fakes_init export: 28 grid prompts / 2, one step() each
    got = []

    def sampler():
        for i in range(n):
            if i > 0:
                got.append(d.randn_like(torch.empty(shape)))
            d.burn(shape, 1)

    for _it in range(2):
        for _r in range(rounds):
            got.append(d.rand_cpu(b))
            z = d.randn(shape)
            got += [z, d.randn_like(z)]
            sampler()
            got.append(d.randint(20, 980, (b,)))
        for _r in range(rounds):
            z = d.randn(shape)
            got += [z, d.randn_like(z), d.randint(20, 980, (b,))]
            sampler()
            d.burn(shape[1:], 2 * b)                                    # predict_x0: per-sample step() in both denoise calls
    want = [t for (_n, _s, line, t) in fx["draws"] if line != 267]
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape and torch.equal(g, w), i


def test_draw_stream_per_rank_seeds():
    from sid_lsg_b200.training.draws import DrawStream
    import numpy as np
    a, b = DrawStream(5, 0, 2, "cpu"), DrawStream(5, 1, 2, "cpu")
    assert a.np_seed == 10 and b.np_seed == 11                        # (seed * W + rank) % 2^31, :238
    np.random.seed(10)
    assert a.torch_seed == np.random.randint(1 << 31)                 # torch.manual_seed(np.random.randint(1 << 31)), :239
    assert not torch.equal(a.randn((4,)), b.randn((4,)))
    c = DrawStream(5, 0, 2, "cpu")
    c.burn((3,), 5)                                                   # not compat: burn is a no-op
    assert torch.equal(c.randn((4,)), DrawStream(5, 0, 2, "cpu").randn((4,)))


def test_dnnlib_and_training_stats():
    from sid_lsg_b200 import dnnlib
    from sid_lsg_b200.torch_utils import training_stats as ts
    c = dnnlib.EasyDict(a=1)
    c.b = 2
    assert c["b"] == 2 and c.a == 1
    with pytest.raises(AttributeError):
        c.missing
    p = [torch.nn.Parameter(torch.zeros(3))]
    opt = dnnlib.util.construct_class_by_name(params=p, class_name="torch.optim.Adam", lr=0.5, betas=[0.0, 0.999])
    assert isinstance(opt, torch.optim.Adam) and opt.param_groups[0]["lr"] == 0.5
    assert dnnlib.util.format_time(59) == "59s" and dnnlib.util.format_time(3661) == "1h 01m 01s"
    assert dnnlib.util.format_time(90061) == "1d 01h 01m"
    ts.report("x/loss", [1.0, 3.0])
    ts.report0("x/tick", 7)
    col = ts.Collector(regex="x/.*")
    col.update()
    assert col.mean("x/loss") == 2.0 and col.num("x/loss") == 2 and abs(col.std("x/loss") - 1.0) < 1e-12
    assert col["x/tick"] == 7.0
    ts.report("x/loss", 5.0)
    col.update()
    assert col.mean("x/loss") == 5.0                                  # window between the last two updates
    col.update()
    assert col.mean("x/loss") == 5.0                                  # keep_previous
    assert set(col.as_dict()["x/loss"]) == {"num", "mean", "std"}


def diffusers_parameter_order(cfg):
    """named_parameters() order of diffusers 0.27.2 UNet2DConditionModel, written down from its module registration
    order (unet_2d_condition.py creates `down_blocks` and `up_blocks` before `mid_block`; the cross-attention blocks
    register `attentions` before `resnets`; Transformer2DModel: norm, proj_in, transformer_blocks, proj_out).
    diffusers is not installable here, so this list is a restatement (as is oracle/unet.py), not a recording."""
    def aff(p):
        return [p + ".weight", p + ".bias"]

    def res(p, shortcut):
        out = aff(p + ".norm1") + aff(p + ".conv1") + aff(p + ".time_emb_proj") + aff(p + ".norm2") + aff(p + ".conv2")
        return out + (aff(p + ".conv_shortcut") if shortcut else [])

    def attn(p):
        return [p + ".to_q.weight", p + ".to_k.weight", p + ".to_v.weight"] + aff(p + ".to_out.0")

    def tr(p):
        b = p + ".transformer_blocks.0"
        blk = (aff(b + ".norm1") + attn(b + ".attn1") + aff(b + ".norm2") + attn(b + ".attn2") + aff(b + ".norm3") +
               aff(b + ".ff.net.0.proj") + aff(b + ".ff.net.2"))
        return aff(p + ".norm") + aff(p + ".proj_in") + blk + aff(p + ".proj_out")

    ch = cfg.block_out_channels
    nb = len(ch)
    names = aff("conv_in") + aff("time_embedding.linear_1") + aff("time_embedding.linear_2")
    cout = ch[0]
    for i in range(nb):
        cin, cout = cout, ch[i]
        p = "down_blocks.%d" % i
        if i < nb - 1:
            for j in range(cfg.layers_per_block):
                names += tr(p + ".attentions.%d" % j)
        for j in range(cfg.layers_per_block):
            names += res(p + ".resnets.%d" % j, (cin if j == 0 else cout) != cout)
        if i < nb - 1:
            names += aff(p + ".downsamplers.0.conv")
    rev = tuple(reversed(ch))
    cout = rev[0]
    for i in range(nb):
        cprev, cout = cout, rev[i]
        cin = rev[min(i + 1, nb - 1)]
        p = "up_blocks.%d" % i
        n = cfg.layers_per_block + 1
        if i > 0:
            for j in range(n):
                names += tr(p + ".attentions.%d" % j)
        for j in range(n):
            names += res(p + ".resnets.%d" % j, True)   # every up-block resnet changes width (skip concat)
        if i < nb - 1:
            names += aff(p + ".upsamplers.0.conv")
    names += tr("mid_block.attentions.0") + res("mid_block.resnets.0", False) + res("mid_block.resnets.1", False)
    return names + aff("conv_norm_out") + aff("conv_out")


@pytest.mark.parametrize("cfg_name", ["SD15", "SD21_BASE", "TINY"])
def test_parameter_order_is_diffusers_order(cfg_name):
    """Index-keyed optimiser state dicts (torch.optim.Adam.state_dict, exchanged with the reference through
    training-state-*.pt) are only meaningful if parameters() enumerates in the same order on both sides."""
    import sid_lsg_b200 as S
    cfg = getattr(S, cfg_name)
    with torch.device("meta"):
        m = S.UNet2DConditionModel(cfg)
    names = [n for n, _ in m.named_parameters()]
    assert len(names) == 686
    assert names == diffusers_parameter_order(cfg)


def test_module_pickle_roundtrip_and_snapshot_format(tmp_path):
    """`pickle.dump({'ema': G_ema})` / `pickle.load(f)['ema']` (sid_training_loop.py:641-650, generate_onestep.py:247-248)
    works on the module itself; the payload is a plain state dict (no flat-bucket storage)."""
    import sid_lsg_b200 as S
    torch.manual_seed(1)
    m = S.UNet2DConditionModel(S.TINY, compute_dtype=torch.bfloat16).eval().requires_grad_(False)
    f = tmp_path / "network-snapshot-1.000000-000001.pkl"
    with open(f, "wb") as fh:
        pickle.dump(dict(ema=m), fh)
    assert os.path.getsize(f) < 1.2 * sum(p.numel() * 4 for p in m.parameters())
    with open(f, "rb") as fh:
        m2 = pickle.load(fh)["ema"]
    assert isinstance(m2, S.UNet2DConditionModel) and m2.compute_dtype == torch.bfloat16 and not m2.training
    assert not any(p.requires_grad for p in m2.parameters())
    for (k, a), (k2, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k == k2 and torch.equal(a, b)
    # S.checkpoint.load_network_snapshot accepts the same file
    m3 = S.load_network_snapshot(str(f), compute_dtype=torch.float32, allow_pickle=True)
    assert torch.equal(m3.state_dict()["conv_in.weight"], m.state_dict()["conv_in.weight"])


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    from sid_lsg_b200.torch_utils import distributed as dist, training_stats as ts, misc
    from sid_lsg_b200.ddp import FlatDDP
    dist.init()                                    # no CUDA here -> gloo (the reference cannot, distributed.py:26-28)
    assert dist.get_world_size() == world and dist.get_rank() == rank
    ts.report("l", float(rank + 1))
    col = ts.Collector(regex="l")
    col.update()                                   # one all_reduce over ranks

    # --- FlatDDP on a toy network with a hand-made flat bucket (host logic only; the real buckets are CUDA) -----
    class Flat:
        def __init__(self, params):
            self.params, self.offsets, n = params, [], 0
            for p in params:
                self.offsets.append(n)
                n += p.numel()
            self.numel = n
            self.master = torch.zeros(n)
            self.grad = torch.zeros(n)
            for p, o in zip(params, self.offsets):
                self.master[o:o + p.numel()] = p.data.flatten()
                p.data = self.master[o:o + p.numel()].view(p.shape)
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
            self.reducer = None

        def refresh_shadow(self):
            pass

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(100 + rank)          # ranks start DIFFERENT: the constructor broadcast must fix that
            self.a, self.b, self.c = torch.nn.Linear(4, 4), torch.nn.Linear(4, 4), torch.nn.Linear(4, 2)
            self.flat = Flat(list(self.parameters()))
            self._grad_ready = None

        def grad_stages(self):
            f, out, i = self.flat, [], 0
            for lin in (self.a, self.b, self.c):
                n = sum(p.numel() for p in lin.parameters())
                out.append([(i, i + n)])
                i += n
            return out

        def forward(self, x):
            ready, self._grad_ready = self._grad_ready, None
            h = self.a(x)
            if ready is not None:
                h.register_hook(lambda g: ready(1))
            h = self.b(torch.tanh(h))
            if ready is not None:
                h.register_hook(lambda g: ready(2))
            return self.c(torch.tanh(h))

    net = Toy()
    ddp = FlatDDP(net)
    w0 = net.flat.master.clone()
    torch.manual_seed(7)
    xs = torch.randn(2 * world, 3, 4)              # [micro-batch, rows, features]; rank r owns micro-batches r, r + world
    mine = [xs[rank], xs[rank + world]]
    for i, x in enumerate(mine):
        with misc.ddp_sync(ddp, i == len(mine) - 1):
            ddp(x).square().sum().backward()
    ddp.finish()
    g = net.flat.grad.clone() / world
    # single-process truth on the concatenated batch with rank 0's weights
    ref = Toy()
    with torch.no_grad():
        ref.flat.master.copy_(w0)
    ref.flat.grad.zero_()
    for x in xs:
        ref(x).square().sum().backward()
    q.put((rank, col.mean("l"), bool(torch.allclose(g, ref.flat.grad / world, rtol=1e-5, atol=1e-6)),
           ddp.reduced_elems, net.flat.numel, w0.tolist()))
    torch.distributed.destroy_process_group()


def test_gloo_world2_dist_shim_stats_and_flat_ddp():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mean0, ok0, red0, n0, w0), (r1, mean1, ok1, red1, n1, w1) = out
    assert mean0 == mean1 == 1.5                   # (1 + 2) / 2 through the stats all_reduce
    assert ok0 and ok1                             # 2 ranks x 2 accumulation rounds == 1 rank on the concatenated batch
    assert red0 == n0 and red1 == n1               # every element reduced exactly once (no_sync rounds reduce nothing)
    assert w0 == w1                                # constructor broadcast: both ranks hold rank 0's weights
