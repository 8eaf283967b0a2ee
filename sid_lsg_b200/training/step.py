"""One SiD-LSG iteration on the B200 kernels: fake-score update, generator update, EMA.

Host orchestration of /root/reference/training/sid_training_loop.py:383-567; every arithmetic step is a kernel
of libsidlsg.so.  Differences from the reference that do not change results:
  * NaN rows are zero-weighted inside the fused loss kernels instead of being filtered by shape (App. B-6);
  * nan_to_num + clip + Adam + EMA (+ bf16 shadow refresh) are ONE pass over the flat buckets;
  * the data-parallel gradient mean runs through ddp.FlatDDP: the trainable networks are wrapped exactly where the
    reference wraps them in DistributedDataParallel (:316-323), every forward sits under `misc.ddp_sync(net, last
    accumulation round)` like the reference's (:406, 416, 487, 494; torch_utils/misc.py:168-175), and the reduction of
    each stage's gradient range starts inside backward (side stream) instead of in 25 MiB autograd buckets.
"""
import torch
import torch.distributed as dist

from .. import ops
from ..ddp import FlatDDP
from ..torch_utils.misc import ddp_sync
from .sid_sd_util import PromptBatch, sid_sd_sampler, sid_sd_denoise


def ema_beta(batch_size, cur_nimg, ema_halflife_kimg, ema_rampup_ratio=0.05):
    """sid_training_loop.py:553-558."""
    ema_halflife_nimg = ema_halflife_kimg * 1000
    if ema_rampup_ratio is not None:
        ema_halflife_nimg = min(ema_halflife_nimg, cur_nimg * ema_rampup_ratio)
    return 0.5 ** (batch_size / max(ema_halflife_nimg, 1e-8))


def _world():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def _val(m, key):
    """a micro-batch entry, or a callable producing it when it is first needed (draws made in the reference's order:
    the f_psi phase draws its timesteps AFTER the generator forward, :413)."""
    v = m[key]
    if callable(v):
        v = v()
        m[key] = v
    return v


class SiDLSGStep:
    """State + the two update phases.  `true_score`, `fake_score`, `G`, `G_ema` are UNet2DConditionModel
    instances with flat buckets (the reference builds them with deepcopy, :284-287, 327)."""

    def __init__(self, true_score, fake_score, G, G_ema, scheduler, *, lr=1e-6, glr=1e-6, betas=(0.0, 0.999),
                 eps=1e-8, fp16=False, alpha=1.0, init_timestep=625, tmin=20, tmax=980, num_steps=1,
                 cfg_train_fake=1.0, cfg_eval_fake=1.0, cfg_eval_real=1.0, loss_scaling=1.0, loss_scaling_G=1.0,
                 ema_halflife_kimg=50, ema_rampup_ratio=0.05, overlap_allreduce=True):
        self.true_score, self.fake_score, self.G, self.G_ema = true_score, fake_score, G, G_ema
        self.sched = scheduler
        self.lr, self.glr, self.betas, self.eps, self.fp16 = lr, glr, betas, eps, fp16
        self.alpha, self.init_timestep, self.tmin, self.tmax, self.num_steps = alpha, init_timestep, tmin, tmax, num_steps
        self.kappa1, self.kappa2, self.kappa4 = cfg_train_fake, cfg_eval_fake, cfg_eval_real
        self.loss_scaling, self.loss_scaling_G = loss_scaling, loss_scaling_G
        self.ema_halflife_kimg, self.ema_rampup_ratio = ema_halflife_kimg, ema_rampup_ratio
        self.cur_nimg = 0
        self._hyper = None       # (f_psi, G_theta) device float[4] buffers while a GraphedIteration drives the step
        for net in (true_score, fake_score, G, G_ema):
            if net is not None and net.flat is None:
                net.flatten_()
        true_score.eval().requires_grad_(False)
        if G_ema is not None and G_ema is not G:
            G_ema.eval().requires_grad_(False)
        fake_score.flat.init_adam(betas[0])
        G.flat.init_adam(betas[0])
        # sid_training_loop.py:316-323 (parameters broadcast from rank 0, gradient mean on synchronised backwards)
        self.fake_score_ddp = FlatDDP(fake_score, overlap=overlap_allreduce)
        self.G_ddp = FlatDDP(G, overlap=overlap_allreduce)
        if _world() > 1 and G_ema is not None and G_ema is not G:
            G_ema.flat.copy_from(G.flat)

    # -- f_psi update: sid_training_loop.py:389-462 --------------------------------------------------------
    def fake_score_phase(self, microbatches, batch_gpu_total=None, rounds=None):
        """microbatches: iterable of dicts {z, noise, t, cond, uncond[, sub_noise]} (device tensors, or callables
        evaluated at the point the reference makes the corresponding draw)."""
        f = self.fake_score
        f.train().requires_grad_(True)
        f.flat.zero_grad()
        if rounds is None:
            microbatches = list(microbatches)
            rounds = len(microbatches)
        total = batch_gpu_total or sum(_val(m, "z").shape[0] for m in microbatches)
        out = None
        for r, m in enumerate(microbatches):
            z = _val(m, "z")
            noise = _val(m, "noise")
            b = z.shape[0]
            prompts = PromptBatch(m["cond"], m["uncond"])
            init_t = torch.full((b,), self.init_timestep, dtype=torch.long, device=z.device)
            with ddp_sync(self.G_ddp, False), torch.no_grad():
                images = sid_sd_sampler(self.G_ddp, z, prompts, init_t, self.sched, num_steps=self.num_steps,
                                        sub_noise=m.get("sub_noise"))
            t = _val(m, "t")
            with ddp_sync(self.fake_score_ddp, r == rounds - 1):
                eps_hat = sid_sd_denoise(self.fake_score_ddp, images, noise, prompts, t, self.sched, predict_x0=False,
                                         guidance_scale=self.kappa1)
                loss, out = ops.fake_loss(eps_hat, noise, self.loss_scaling / total)
                loss.backward()
        f.eval().requires_grad_(False)
        f.flat.adam_step(self.lr, self.betas, self.eps, grad_scale=1.0 / _world(),
                         hyper=self._hyper[0] if self._hyper else None)
        return out  # device float[2] {loss of the last micro-batch, valid rows}: no host sync here

    # -- G_theta update: sid_training_loop.py:468-549, EMA :553-565 -----------------------------------------
    def generator_phase(self, microbatches, batch_gpu_total=None, batch_size=None, return_images=False, rounds=None):
        G, phi = self.G, self.true_score
        G.train().requires_grad_(True)
        G.flat.zero_grad()
        if rounds is None:
            microbatches = list(microbatches)
            rounds = len(microbatches)
        total = batch_gpu_total or sum(_val(m, "z").shape[0] for m in microbatches)
        out = None
        imgs = []
        for r, m in enumerate(microbatches):
            z = _val(m, "z")
            noise = _val(m, "noise")
            t = _val(m, "t")
            b = z.shape[0]
            prompts = PromptBatch(m["cond"], m["uncond"])
            init_t = torch.full((b,), self.init_timestep, dtype=torch.long, device=z.device)
            with ddp_sync(self.G_ddp, r == rounds - 1):
                images = sid_sd_sampler(self.G_ddp, z, prompts, init_t, self.sched, num_steps=self.num_steps,
                                        sub_noise=m.get("sub_noise"))
                with ddp_sync(self.fake_score_ddp, False):
                    y_fake = sid_sd_denoise(self.fake_score_ddp, images, noise, prompts, t, self.sched,
                                            guidance_scale=self.kappa2)
                    y_real = sid_sd_denoise(phi, images, noise, prompts, t, self.sched, guidance_scale=self.kappa4)
                    loss, out = ops.lsg_loss(images, y_real, y_fake, self.alpha, self.loss_scaling_G / total)
                    loss.backward()
            if return_images:
                imgs.append(images.detach())
        G.eval().requires_grad_(False)
        world = _world()
        bs = batch_size or total * world
        beta = 0.0
        ema = None
        if self.G_ema is not None and self.G_ema is not G and self.ema_halflife_kimg > 0:
            beta = ema_beta(bs, self.cur_nimg, self.ema_halflife_kimg, self.ema_rampup_ratio)
            ema = self.G_ema.flat
        G.flat.adam_step(self.glr, self.betas, self.eps, grad_scale=1.0 / world, clip=1.0 if self.fp16 else 0.0,
                         ema=ema, ema_beta=beta, hyper=self._hyper[1] if self._hyper else None)
        self.cur_nimg += bs
        if return_images:
            return out, torch.cat(imgs)
        return out

    def iteration(self, mb_f, mb_g, batch_size=None):
        lf = self.fake_score_phase(mb_f)
        lg = self.generator_phase(mb_g, batch_size=batch_size)
        return lf, lg


class GraphedIteration:
    """One whole iteration (both phases, optimiser passes, EMA, the gradient allreduces) captured ONCE as a CUDA graph and
    replayed per step: ~4,800 kernel launches become one graph launch (measured on B200, SD1.5 batch 32: host enqueue
    846 -> 170 ms per iteration, device time 867 -> 852 ms).  Inputs are copied into static device buffers before every
    replay; the per-step scalars of the optimiser passes (Adam bias corrections, EMA beta, learning rates) are
    recomputed ON THE DEVICE by the graph's first node from device-side counters (sidlsg_hyper_advance), so a replay
    follows the step count instead of repeating the captured one and the host never has to wait for the device.
    Shapes must not change between steps.  Run at least one eager iteration before constructing this (lazy gradient
    buckets, the autograd worker thread's CUDA context)."""

    def __init__(self, step, mb_f, mb_g, batch_size=None):
        from .._lib import lib, ptr, stream
        self.step = step
        dev = mb_f[0]["z"].device

        def static(mbs):
            return [{k: ([x.clone() for x in v] if isinstance(v, list) else v.clone()) for k, v in m.items()} for m in mbs]
        self.static_f, self.static_g = static(mb_f), static(mb_g)
        self.batch_size = batch_size or sum(m["z"].shape[0] for m in self.static_g) * _world()
        st = step
        self.counters = torch.tensor([st.fake_score.flat.step_count, st.G.flat.step_count, st.cur_nimg], dtype=torch.int64,
                                     device=dev)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        ema_on = st.G_ema is not None and st.G_ema is not st.G and st.ema_halflife_kimg > 0
        rampup = -1.0 if st.ema_rampup_ratio is None else float(st.ema_rampup_ratio)
        host_counters = (st.fake_score.flat.step_count, st.G.flat.step_count, st.cur_nimg)
        st._hyper = (self.hyper[:4], self.hyper[4:])
        torch.cuda.synchronize(dev)
        torch.cuda.empty_cache()                      # the graph's private pool should not have to sit beside the eager one
        self.graph = torch.cuda.CUDAGraph()
        calls0 = lib.launches
        try:
            with torch.cuda.graph(self.graph):
                lib.call("hyper_advance", ptr(self.hyper), ptr(self.counters), float(st.lr), float(st.glr),
                         float(st.betas[0]), float(st.betas[1]), float(self.batch_size),
                         float(st.ema_halflife_kimg * 1000), rampup, 1 if ema_on else 0, stream())
                self.out = st.iteration(self.static_f, self.static_g, batch_size=self.batch_size)
        finally:
            st._hyper = None
            # capture ran the host side once without executing anything: undo its bookkeeping
            st.fake_score.flat.step_count, st.G.flat.step_count, st.cur_nimg = host_counters
        self.calls_per_replay = lib.launches - calls0     # C-ABI calls (kernel launches) one replay stands for

    def __call__(self, mb_f, mb_g):
        for dst_list, src_list in ((self.static_f, mb_f), (self.static_g, mb_g)):
            for dst, src in zip(dst_list, src_list):
                for k, v in dst.items():
                    if isinstance(v, list):
                        for a, b in zip(v, src[k]):
                            a.copy_(b, non_blocking=True)
                    else:
                        v.copy_(src[k], non_blocking=True)
        st = self.step                                # host mirrors of the device-side counters (checkpoints, logging)
        st.fake_score.flat.step_count += 1
        st.G.flat.step_count += 1
        st.cur_nimg += self.batch_size
        self.graph.replay()
        return self.out


def synth_microbatch(b, cfg, seed, device, dropout=False, num_steps=1, tmin=20, tmax=980, pinned=False):
    """Synthetic inputs of SURVEY.md §8d, drawn on the HOST with the same generator recipe as the oracle's
    synth_microbatch (so the CUDA path and the CPU oracle see identical numbers), then copied to `device`."""
    g = torch.Generator().manual_seed(seed)
    hw, d = cfg.sample_size, cfg.cross_attention_dim
    ug = torch.Generator().manual_seed(1234567)
    uncond1 = torch.randn([1, 77, d], generator=ug)
    cond = torch.randn([b, 77, d], generator=g)
    uncond = uncond1.expand(b, 77, d).contiguous()
    if dropout:
        drop = torch.rand(b, generator=g) < 0.1
        cond = torch.where(drop[:, None, None], uncond, cond)
    m = dict(cond=cond, uncond=uncond,
             z=torch.randn([b, cfg.in_channels, hw, hw], generator=g),
             noise=torch.randn([b, cfg.in_channels, hw, hw], generator=g),
             t=torch.randint(tmin, tmax, (b,), generator=g, dtype=torch.long))
    if num_steps > 1:
        m["sub_noise"] = [torch.randn([b, cfg.in_channels, hw, hw], generator=g) for _ in range(num_steps - 1)]
    if pinned:
        m = {k: ([x.pin_memory() for x in v] if isinstance(v, list) else v.pin_memory()) for k, v in m.items()}
    if device is None:
        return m
    return to_device(m, device)


def device_microbatch(b, cfg, draws, dropout=False, uncond=None, tmin=20, tmax=980):
    """Synthetic micro-batch drawn ON THE DEVICE from a training.draws.DrawStream (the reference's on-device z /
    noise / t sampling, sid_training_loop.py:398-402, 413, 479-484, with synthetic prompt embeddings in place of the
    text encoder): no host->device traffic at all."""
    d = cfg.cross_attention_dim
    hw = cfg.sample_size
    if uncond is None:
        uncond = draws.randn([1, 77, d])
    uncond = uncond.expand(b, 77, d)
    cond = draws.randn([b, 77, d])
    if dropout:
        drop = (draws.rand_cpu(b) < 0.1).to(cond.device)
        cond = torch.where(drop[:, None, None], uncond, cond)
    z = draws.randn([b, cfg.in_channels, hw, hw])
    return dict(cond=cond, uncond=uncond, z=z, noise=draws.randn_like(z), t=draws.randint(tmin, tmax, (b,)))


def to_device(m, device):
    return {k: ([x.to(device, non_blocking=True) for x in v] if isinstance(v, list) else v.to(device, non_blocking=True))
            for k, v in m.items()}
