#!/usr/bin/env python
"""Generate the committed parity fixtures by running the REFERENCE's own code.

Runs in the build container only (needs /root/reference; the GPU box never runs this).

diffusers / xformers are not installable here, so the reference's glue and training loop are
imported unmodified from /root/reference with stub `diffusers` / `metrics` modules, and are handed
the oracle's UNet + DDPM schedule (oracle/unet.py, oracle/scheduler.py) plus a deterministic
table-lookup tokenizer / text encoder.  What this pins:

  glue.pt  : outputs of the reference's sid_sd_sampler / sid_sd_denoise
             (/root/reference/training/sid_sd_util.py:163-274) on seeded inputs
  loop.pt  : two full iterations of the reference's training_loop
             (/root/reference/training/sid_training_loop.py:383-567) on CPU/gloo, with every RNG
             draw it made recorded, the losses it reported and slices of the weights it produced
  sampler_order.pt : first indices of the reference's InfiniteSampler (torch_utils/misc.py:110-141)

    python tests/golden/make_golden.py
"""
import inspect
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import DDPMSchedule, UNet2DCondition, TINY  # noqa: E402
from oracle.scheduler import compute_snr  # noqa: E402

D = TINY.cross_attention_dim
PROMPTS = [f"prompt number {i}" for i in range(1, 41)]
WATCH = ["conv_in.weight", "conv_out.weight", "time_embedding.linear_1.bias",
         "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight",
         "mid_block.resnets.1.conv2.bias", "up_blocks.3.resnets.2.conv_shortcut.weight",
         "up_blocks.1.attentions.0.transformer_blocks.0.ff.net.0.proj.bias"]


# ---- stubs -----------------------------------------------------------------------------------
def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dummy:
        pass

    mod("diffusers", AutoencoderKL=_Dummy, DDPMScheduler=_Dummy, DiffusionPipeline=_Dummy,
        UNet2DConditionModel=_Dummy, __version__="0.27.2")
    mod("diffusers.loaders", StableDiffusionXLLoraLoaderMixin=_Dummy)
    mod("diffusers.optimization", get_scheduler=None)
    mod("diffusers.utils", check_min_version=lambda *_: None, convert_state_dict_to_diffusers=None)
    mod("diffusers.utils.import_utils", is_xformers_available=lambda: False)
    mod("diffusers.models", )
    mod("diffusers.models.attention_processor", AttnProcessor2_0=_Dummy, XFormersAttnProcessor=_Dummy,
        LoRAXFormersAttnProcessor=_Dummy, LoRAAttnProcessor2_0=_Dummy, FusedAttnProcessor2_0=_Dummy)
    mod("diffusers.training_utils", compute_snr=compute_snr)
    metrics = mod("metrics")
    metrics.sid_metric_main = mod("metrics.sid_metric_main")


class StubTokenizer:
    """prompt -> id row; '' -> 0, PROMPTS[i] -> i+1.  Same call signature the reference uses."""
    model_max_length = 77

    def __call__(self, prompt, padding=None, max_length=None, truncation=None, return_tensors=None):
        ids = [0 if p == "" else PROMPTS.index(p) + 1 for p in prompt]
        return types.SimpleNamespace(input_ids=torch.tensor(ids, dtype=torch.long)[:, None].repeat(1, 77))


def embedding_table():
    g = torch.Generator().manual_seed(777)
    return torch.randn([len(PROMPTS) + 1, 77, D], generator=g)


class StubTextEncoder:
    def __init__(self):
        self.table = embedding_table()
        self.calls = []

    def __call__(self, input_ids):
        self.calls.append(input_ids[:, 0].clone())
        return (self.table[input_ids[:, 0]],)


class StubVAE:
    dtype = torch.float32
    config = types.SimpleNamespace(force_upcast=False, scaling_factor=0.18215, block_out_channels=(1, 1, 1, 1))

    def decode(self, x, return_dict=False):
        return (x[:, :3].clamp(-1, 1),)


class PromptSet(torch.utils.data.Dataset):
    name = "stub"
    resolution = 512

    def __len__(self):
        return len(PROMPTS)

    def __getitem__(self, idx):
        return np.zeros((3, 8, 8), dtype=np.float32), PROMPTS[idx]


def fresh_unet():
    torch.manual_seed(0)
    return UNet2DCondition(TINY)


# ---- part A: glue ------------------------------------------------------------------------------
def make_glue(ref_util):
    unet = fresh_unet().eval().requires_grad_(False)
    sched = DDPMSchedule()
    tok, te = StubTokenizer(), StubTextEncoder()
    g = torch.Generator().manual_seed(11)
    b = 3
    z = torch.randn([b, 4, 16, 16], generator=g)
    noise = torch.randn([b, 4, 16, 16], generator=g)
    t = torch.tensor([20, 500, 979])
    ctx = [PROMPTS[4], "", PROMPTS[17]]
    out = dict(z=z, noise=noise, t=t, ctx_ids=torch.tensor([5, 0, 18]))
    init_t = 625 * torch.ones((b,), dtype=torch.long)
    kw = dict(noise_scheduler=sched, text_encoder=te, tokenizer=tok, resolution=128, dtype=torch.float32)
    with torch.no_grad():
        out["sampler_1step"] = ref_util.sid_sd_sampler(unet=unet, latents=z, contexts=ctx, init_timesteps=init_t, **kw)
        images = out["sampler_1step"]
        for kappa in (1, 1.5, 4.5):
            out[f"denoise_x0_k{kappa}"] = ref_util.sid_sd_denoise(unet=unet, images=images, noise=noise, contexts=ctx,
                                                                  timesteps=t, guidance_scale=kappa, **kw)
            out[f"denoise_eps_k{kappa}"] = ref_util.sid_sd_denoise(unet=unet, images=images, noise=noise, contexts=ctx,
                                                                   timesteps=t, guidance_scale=kappa, predict_x0=False, **kw)
        # 4-step sampler: randn_like draws are recorded through the patched torch.randn_like
        rec = []
        orig = torch.randn_like

        def rl(x, *a, **k):
            r = orig(x, *a, **k)
            rec.append(r.clone())
            return r
        torch.randn_like = rl
        try:
            torch.manual_seed(5)
            out["sampler_4step"] = ref_util.sid_sd_sampler(unet=unet, latents=z, contexts=ctx, init_timesteps=init_t,
                                                           num_steps=4, **kw)
        finally:
            torch.randn_like = orig
        out["sampler_4step_sub_noise"] = torch.stack(rec)
    out["unet_checksum"] = torch.tensor([sum(float(p.double().sum()) for p in unet.parameters()),
                                         sum(float(p.double().abs().sum()) for p in unet.parameters())],
                                        dtype=torch.float64)
    torch.save(out, os.path.join(HERE, "glue.pt"))
    print("glue.pt", {k: tuple(v.shape) for k, v in out.items()})


# ---- part B: the whole training loop ---------------------------------------------------------
def make_loop(ref_loop, num_steps, fname, kappa=1.5, alpha=1.0, batch=4, batch_gpu=2, lr=1e-3):
    import dnnlib  # the reference's
    unet = fresh_unet()
    sched = DDPMSchedule()
    tok, te, vae = StubTokenizer(), StubTextEncoder(), StubVAE()
    ref_loop.load_sd15 = lambda **kw: (unet, vae, sched, te, tok)
    sys.modules["golden_stub_data"] = types.ModuleType("golden_stub_data")
    sys.modules["golden_stub_data"].PromptSet = PromptSet

    # record every RNG draw made from the reference's own files
    draws = []
    orig = dict(randn=torch.randn, randn_like=torch.randn_like, randint=torch.randint, rand=torch.rand)

    def wrap(name):
        def f(*a, **k):
            r = orig[name](*a, **k)
            fr = inspect.stack()[1]
            src = os.path.basename(fr.filename)
            if src in ("sid_training_loop.py", "sid_sd_util.py"):
                draws.append((name, src, fr.lineno, r.detach().clone().cpu()))
            return r
        return f
    losses = []
    from torch_utils import training_stats
    orig_report = training_stats.report

    def report(name, value):
        losses.append((name, float(value)))
        return orig_report(name, value)

    real_ddp = torch.nn.parallel.DistributedDataParallel

    class CpuDDP(real_ddp):
        def __init__(self, module, device_ids=None, **kw):
            super().__init__(module, **kw)

    te_mark = {}
    patches = [(torch, n, wrap(n)) for n in orig]
    patches += [(training_stats, "report", report),
                (torch.nn.parallel, "DistributedDataParallel", CpuDDP),
                (torch.cuda, "max_memory_allocated", lambda *a: 0),
                (torch.cuda, "max_memory_reserved", lambda *a: 0),
                (torch.cuda, "reset_peak_memory_stats", lambda *a: None)]
    saved = [(o, n, getattr(o, n)) for o, n, _ in patches]
    for o, n, v in patches:
        setattr(o, n, v)
    run_dir = tempfile.mkdtemp()
    try:
        opt = dict(class_name="torch.optim.Adam", lr=lr, betas=[0.0, 0.999], eps=1e-8)
        ref_loop.training_loop(
            run_dir=run_dir, dataset_kwargs={}, data_loader_kwargs=dict(pin_memory=False, num_workers=0),
            network_kwargs=dnnlib.EasyDict(use_fp16=False), loss_kwargs={},
            fake_score_optimizer_kwargs=dict(opt), g_optimizer_kwargs=dict(opt), seed=3,
            batch_size=batch, batch_gpu=batch_gpu, total_kimg=2 * batch / 1000, ema_halflife_kimg=50,
            ema_rampup_ratio=0.05, loss_scaling=1, loss_scaling_G=100, kimg_per_tick=0, snapshot_ticks=None,
            state_dump_ticks=1, alpha=alpha, tmax=980, tmin=20, device=torch.device("cpu"), metrics=None,
            init_timestep=625, dataset_prompt_text_kwargs=dict(class_name="golden_stub_data.PromptSet"),
            cfg_train_fake=kappa, cfg_eval_fake=kappa, cfg_eval_real=kappa, num_steps=num_steps,
            enable_xformers=False, resolution=128)
        te_mark["calls"] = [c.clone() for c in te.calls]
    finally:
        for o, n, v in saved:
            setattr(o, n, v)
    state = torch.load(os.path.join(run_dir, f"training-state-{0:06d}.pt"), weights_only=False)
    out = dict(num_steps=num_steps, kappa=kappa, alpha=alpha, batch=batch, batch_gpu=batch_gpu, lr=lr,
               loss_scaling=1, loss_scaling_G=100, ema_halflife_kimg=50,
               draws=[(n, s, l, t) for n, s, l, t in draws],
               losses=losses, te_calls=te_mark["calls"])
    for key in ("G", "fake_score", "G_ema"):
        sd = state[key].state_dict()
        out[key] = {k: sd[k].clone() for k in WATCH}
        out[key + "_sum"] = torch.tensor([sum(float(v.double().sum()) for v in sd.values()),
                                          sum(float(v.double().abs().sum()) for v in sd.values())], dtype=torch.float64)
    torch.save(out, os.path.join(HERE, fname))
    print(fname, "draws", len(draws), "losses", losses, "te_calls", len(out["te_calls"]))


def make_sampler_order():
    from torch_utils import misc  # the reference's
    ds = list(range(37))
    out = {}
    for rank, world in ((0, 1), (0, 2), (1, 2), (3, 8)):
        it = iter(misc.InfiniteSampler(ds, rank=rank, num_replicas=world, seed=3))
        out[f"r{rank}w{world}"] = torch.tensor([int(next(it)) for _ in range(64)])
    torch.save(out, os.path.join(HERE, "sampler_order.pt"))
    print("sampler_order.pt")


def main():
    install_stubs()
    # torch>=2.2 compat: the reference's InfiniteSampler calls Sampler.__init__(dataset) (torch 2.3 accepted it)
    torch.utils.data.Sampler.__init__ = lambda self, *a, **k: None
    sys.path.insert(0, REF)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.distributed.init_process_group("gloo", rank=0, world_size=1)
    from training import sid_sd_util as ref_util
    from training import sid_training_loop as ref_loop
    make_glue(ref_util)
    make_sampler_order()
    make_loop(ref_loop, num_steps=1, fname="loop_1step.pt")
    make_loop(ref_loop, num_steps=2, fname="loop_2step_alpha12.pt", kappa=2.0, alpha=1.2)
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
