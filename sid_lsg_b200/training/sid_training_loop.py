"""`training_loop(**c)`: the reference's distillation entry point on the B200 kernels.

Same keyword arguments as /root/reference/training/sid_training_loop.py:148-194 (called as `training_loop(**c)` from
sid_train.py:372), same phases, bookkeeping and files: tick status lines through `training_stats`,
`stats_<alpha>.jsonl`, `network-snapshot-<alpha>-<kimg>.pkl` = `pickle.dump({'ema': G_ema})`,
`training-state-<kimg>.pt` with the reference's five keys, resume through `copy_params_and_buffers` +
`optimizer.load_state_dict`.  What differs is what runs underneath:

  * networks are sid_lsg_b200.UNet2DConditionModel (hand-written sm_100a kernels), wrapped in ddp.FlatDDP where the
    reference wraps torch DDP; optimisers are params.FlatAdam (fused nan_to_num + clip + Adam + EMA pass);
  * prompts are tokenised / encoded ONCE per micro-batch (training/prompts.PromptEncoder) instead of 8 times;
  * z / noise / timesteps come from training/draws.DrawStream: the reference's seed recipe (:238-239) and draw order,
    on the device.  `rng_compat=True` (extra keyword, default False) additionally reproduces the RNG consumption of
    diffusers' scheduler.step(), so the numbers drawn equal the reference's for the same seed;
  * `fp16` requests (network_kwargs.use_fp16) select the bf16 tensor-core mode with fp32 master weights - the
    reference's pure-half weights have no counterpart here; use_fp16=False selects the fp32-exact mode.
  * evaluation mode (`train_mode=False`, FID/CLIP metrics) and `metrics=` are outside the hot path:
    requesting them raises NotImplementedError rather than silently skipping.
"""
import copy
import gc
import json
import os
import pickle
import time

import numpy as np
import psutil
import torch

from .. import dnnlib
from ..ddp import FlatDDP
from ..params import FlatAdam
from ..torch_utils import distributed as dist
from ..torch_utils import misc
from ..torch_utils import training_stats
from . import step as _step
from .draws import DrawStream
from .prompts import PromptEncoder
from .sid_sd_util import load_sd15, sid_sd_sampler, sid_sd_denoise, PromptBatch  # noqa: F401  (load_sd15: patch point)


# ---- helpers with the reference's names (:39-145) ---------------------------------------------------------------
def setup_snapshot_image_grid(training_set, random_seed=0):
    rnd = np.random.RandomState(random_seed)
    gw = int(np.clip(3840 // training_set.resolution, 7, 32))
    gh = int(np.clip(2160 // training_set.resolution, 4, 32))
    order = list(range(len(training_set)))
    rnd.shuffle(order)
    picks = [order[i % len(order)] for i in range(gw * gh)]
    images, contexts = zip(*[training_set[i] for i in picks])
    return (gw, gh), np.stack(images), contexts


def split_list(lst, split_sizes):
    if isinstance(split_sizes, int):
        n = split_sizes
        return [list(lst[i:i + n]) for i in range(0, len(lst), n)]
    out, i = [], 0
    for n in split_sizes:
        out.append(list(lst[i:i + n]))
        i += n
    return out


def save_image_grid(img, fname, drange, grid_size):
    import PIL.Image
    lo, hi = drange
    img = np.asarray(img, dtype=np.float32)
    img = np.rint((img - lo) * (255 / (hi - lo))).clip(0, 255).astype(np.uint8)
    gw, gh = grid_size
    _n, C, H, W = img.shape
    img = img.reshape(gh, gw, C, H, W).transpose(0, 3, 1, 4, 2).reshape(gh * H, gw * W, C)
    assert C in (1, 3)
    PIL.Image.fromarray(img[:, :, 0] if C == 1 else img, "L" if C == 1 else "RGB").save(fname)


def save_data(data, fname):
    with open(fname, "wb") as f:
        pickle.dump(data, f)


def save_pt(pt, fname):
    torch.save(pt, fname)


def append_line(jsonl_line, fname):
    with open(fname, "at") as f:
        f.write(jsonl_line + "\n")


def _make_optimizer(net, kwargs):
    """`dnnlib.util.construct_class_by_name(params=net.parameters(), **kwargs)` of the reference (:291-292) for the two
    optimiser classes sid_train.py can ask for (:219-226): both map onto the fused flat-bucket pass."""
    kw = dict(kwargs)
    name = kw.pop("class_name", "torch.optim.Adam")
    if name not in ("torch.optim.Adam", "torch.optim.AdamW"):
        raise NotImplementedError("optimizer %r: the fused pass implements torch.optim.Adam / AdamW" % name)
    if name == "torch.optim.AdamW":
        kw.setdefault("weight_decay", 0.01)
    return FlatAdam(net.flat, decoupled=(name == "torch.optim.AdamW"), **kw)


def _export_grid(G_ema, grid_z, grid_c, enc, sched, vae, init_timestep, device, resolution, dtype, num_steps,
                 num_steps_eval, fname, grid_size):
    """the reference's sample-image export (:347-354, 603-615): always runs the sampler (it is part of the RNG stream);
    writes a PNG when a VAE is available, the latents otherwise."""
    outs = []
    for z, c in zip(grid_z, grid_c):
        init_t = init_timestep * torch.ones((len(c),), device=device, dtype=torch.long)
        outs.append(sid_sd_sampler(unet=G_ema, latents=z.float(), contexts=enc.encode(c), init_timesteps=init_t,
                                   noise_scheduler=sched, resolution=resolution, dtype=dtype,
                                   return_images=vae is not None, vae=vae, num_steps=num_steps, train_sampler=False,
                                   num_steps_eval=num_steps_eval))
    images = torch.cat(outs).cpu()
    if vae is not None:
        save_image_grid(img=images.numpy(), fname=fname, drange=[-1, 1], grid_size=grid_size)
    else:
        torch.save(images, os.path.splitext(fname)[0] + "_latents.pt")


def training_loop(
    run_dir=".", dataset_kwargs={}, data_loader_kwargs={}, network_kwargs={}, loss_kwargs={},
    fake_score_optimizer_kwargs={}, g_optimizer_kwargs={}, augment_kwargs=None, seed=0, batch_size=512,
    batch_gpu=None, total_kimg=200000, ema_halflife_kimg=500, ema_rampup_ratio=0.05, loss_scaling=1,
    loss_scaling_G=1, kimg_per_tick=50, snapshot_ticks=50, state_dump_ticks=500, resume_pkl=None,
    resume_training=None, resume_kimg=0, alpha=1, tmax=980, tmin=20, cudnn_benchmark=True,
    device=torch.device("cuda"), metrics=None, init_timestep=None, metric_pt_path=None, metric_open_clip_path=None,
    metric_clip_path=None, pretrained_model_name_or_path="runwayml/stable-diffusion-v1-5",
    pretrained_vae_model_name_or_path="runwayml/stable-diffusion-v1-5", fake_score_use_lora=False,
    dataset_prompt_text_kwargs={}, cfg_train_fake=1, cfg_eval_fake=1, cfg_eval_real=1, num_steps=1, train_mode=True,
    network_pkl=None, enable_xformers=True, gradient_checkpointing=False, resolution=512,
    rng_compat=False, rng_device=None, on_iteration=None,
):
    del dataset_kwargs, loss_kwargs, augment_kwargs, resume_pkl, cudnn_benchmark, metric_pt_path
    del metric_open_clip_path, metric_clip_path, pretrained_vae_model_name_or_path, fake_score_use_lora, network_pkl
    if not train_mode or metrics is not None:
        raise NotImplementedError("evaluation mode / FID-CLIP metrics are outside the distillation hot path "
                                  "(SURVEY.md §8): run them with the reference's tooling on the exported snapshots")
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    rank, world = dist.get_rank(), dist.get_world_size()
    use_fp16 = bool(getattr(network_kwargs, "use_fp16", False) if not isinstance(network_kwargs, dict)
                    else network_kwargs.get("use_fp16", False))
    dtype = torch.float16 if use_fp16 else torch.float32

    dist.print0("Loading dataset...")
    dataset_obj = dnnlib.util.construct_class_by_name(**dataset_prompt_text_kwargs)
    use_context_dropout_train_fake = (cfg_train_fake != 1 or cfg_eval_fake != 1)   # :208-211

    # weights are loaded rank-0 first behind barriers (:219-231)
    if rank != 0:
        dist.barrier()
    unet, vae, noise_scheduler, text_encoder, tokenizer = load_sd15(
        pretrained_model_name_or_path=pretrained_model_name_or_path, pretrained_vae_model_name_or_path=None,
        device=device, weight_dtype=dtype, variant="fp16" if use_fp16 else None, enable_xformers=enable_xformers,
        lora_config=None)
    if rank == 0:
        dist.barrier()
    dist.print0("Loading network completed")

    start_time = time.time()
    draws = DrawStream(seed, rank, world, device=device, rng_device=rng_device, compat=rng_compat)   # :238-239
    noise_scheduler.rng = draws
    enc = PromptEncoder(tokenizer, text_encoder, device=device)

    batch_gpu_total = batch_size // world                                            # :246-250
    if batch_gpu is None or batch_gpu > batch_gpu_total:
        batch_gpu = batch_gpu_total
    num_accumulation_rounds = batch_gpu_total // batch_gpu
    assert batch_size == batch_gpu * num_accumulation_rounds * world

    latent_img_channels = 4
    vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1) if vae is not None else 8
    latent_resolution = resolution // vae_scale_factor

    grid_size = grid_z = grid_c = None
    if rank == 0:                                                                    # :258-271 (seed 2024, restored after)
        grid_size, grid_images, contexts = setup_snapshot_image_grid(training_set=dataset_obj)
        g2024 = torch.Generator(device=draws.rng_device).manual_seed(2024)
        grid_z = torch.randn([len(contexts), latent_img_channels, latent_resolution, latent_resolution],
                             generator=g2024, device=draws.rng_device, dtype=torch.float32).to(device)
        grid_z = grid_z.split(batch_gpu)
        grid_c = split_list(contexts, batch_gpu)

    sampler = misc.InfiniteSampler(dataset=dataset_obj, rank=rank, num_replicas=world, seed=seed)
    prompt_iterator = iter(torch.utils.data.DataLoader(dataset=dataset_obj, sampler=sampler, batch_size=batch_gpu,
                                                       generator=draws.cpu_gen, **data_loader_kwargs))
    dist.print0("Example text prompts used for distillation:")
    for _i in range(16):                                                             # :277-281 (advances the prompt stream)
        _, contexts = next(prompt_iterator)
        dist.print0(_i, contexts)

    true_score = unet                                                                # :284-287
    true_score.eval().requires_grad_(False).to(device)
    fake_score = copy.deepcopy(true_score).train().requires_grad_(True).to(device)
    G = copy.deepcopy(true_score).train().requires_grad_(True).to(device)
    for net in (true_score, fake_score, G):
        net.flatten_() if net.flat is None else None

    dist.print0("Setting up optimizer...")
    fake_score_optimizer = _make_optimizer(fake_score, fake_score_optimizer_kwargs)  # :291-292
    g_optimizer = _make_optimizer(G, g_optimizer_kwargs)

    G_ema = None
    if resume_training is not None:                                                  # :296-317
        dist.print0("checkpoint path:", resume_training)
        data = torch.load(resume_training, map_location=torch.device("cpu"), weights_only=False)
        misc.copy_params_and_buffers(src_module=data["fake_score"], dst_module=fake_score, require_all=True)
        misc.copy_params_and_buffers(src_module=data["G"], dst_module=G, require_all=True)
        if ema_halflife_kimg > 0:
            G_ema = copy.deepcopy(G).eval().requires_grad_(False)
            misc.copy_params_and_buffers(src_module=data["G_ema"], dst_module=G_ema, require_all=True)
        fake_score_optimizer.load_state_dict(data["fake_score_optimizer_state"])
        g_optimizer.load_state_dict(data["g_optimizer_state"])
        del data
        dist.print0("Loading checkpoint completed")
        dist.barrier()
    dist.print0("Setting up GPU parallel computing")
    fake_score_ddp = FlatDDP(fake_score, device_ids=[device], broadcast_buffers=False, find_unused_parameters=False)
    G_ddp = FlatDDP(G, device_ids=[device], broadcast_buffers=False, find_unused_parameters=False)
    if G_ema is None:
        G_ema = copy.deepcopy(G).eval().requires_grad_(False) if ema_halflife_kimg > 0 else G   # :324-327
    fake_score_ddp.eval().requires_grad_(False)
    G_ddp.eval().requires_grad_(False)

    dist.print0(f"Training for {total_kimg} kimg...")
    dist.print0()
    cur_nimg = resume_kimg * 1000
    cur_tick = 0
    tick_start_nimg = cur_nimg
    tick_start_time = time.time()
    maintenance_time = tick_start_time - start_time
    dist.update_progress(cur_nimg // 1000, total_kimg)

    sampler_kw = dict(noise_scheduler=noise_scheduler, resolution=resolution, dtype=dtype)
    if resume_training is None and rank == 0:                                        # :343-355
        print("Exporting sample fake images at initialization...")
        _export_grid(G_ema, grid_z, grid_c, enc, noise_scheduler, vae, init_timestep, device, resolution, dtype,
                     num_steps, 1, os.path.join(run_dir, "fakes_init.png"), grid_size)
    dist.barrier()

    fp16_clip = 1.0 if dtype == torch.float16 else 0.0
    inv_world = 1.0 / world
    dist.print0("Start Running")
    while True:
        # ---- fake-score update (:389-462) ----------------------------------------------------------------------
        G_ddp.eval().requires_grad_(False)
        fake_score_ddp.train().requires_grad_(True)
        fake_score_optimizer.zero_grad(set_to_none=True)
        for round_idx in range(num_accumulation_rounds):
            _, contexts = next(prompt_iterator)
            prompts = enc.encode(contexts)
            if use_context_dropout_train_fake:
                drop = draws.rand_cpu(len(contexts)) < 0.1                           # :393-396, on embeddings
                prompts = PromptBatch(torch.where(drop.to(device)[:, None, None], prompts.uncond, prompts.cond),
                                      prompts.uncond)
            z = draws.randn([len(contexts), latent_img_channels, latent_resolution, latent_resolution])
            noise = draws.randn_like(z)
            init_timesteps = init_timestep * torch.ones((len(contexts),), device=device, dtype=torch.long)
            with misc.ddp_sync(G_ddp, False), torch.no_grad():
                images = sid_sd_sampler(unet=G_ddp, latents=z, contexts=prompts, init_timesteps=init_timesteps,
                                        num_steps=num_steps, **sampler_kw)
            timesteps = draws.randint(tmin, tmax, (len(contexts),))
            with misc.ddp_sync(fake_score_ddp, round_idx == num_accumulation_rounds - 1):
                noise_fake = sid_sd_denoise(unet=fake_score_ddp, images=images, noise=noise, contexts=prompts,
                                            timesteps=timesteps, predict_x0=False, guidance_scale=cfg_train_fake,
                                            **sampler_kw)
                # NaN rows are zero-weighted inside the kernel (shape-stable form of :423-436)
                loss, loss_fake_dev = _step.ops.fake_loss(noise_fake, noise, loss_scaling / batch_gpu_total)
                del images
                loss.backward()
        fake_score_ddp.eval().requires_grad_(False)
        fake_score_optimizer.step(grad_scale=inv_world)           # nan_to_num (:458-460) + Adam (:462), one pass

        # ---- generator update (:468-549) -----------------------------------------------------------------------
        G_ddp.train().requires_grad_(True)
        g_optimizer.zero_grad(set_to_none=True)
        for round_idx in range(num_accumulation_rounds):
            _, contexts = next(prompt_iterator)
            prompts = enc.encode(contexts)
            z = draws.randn([len(contexts), latent_img_channels, latent_resolution, latent_resolution])
            noise = draws.randn_like(z)
            init_timesteps = init_timestep * torch.ones((len(contexts),), device=device, dtype=torch.long)
            timesteps = draws.randint(tmin, tmax, (len(contexts),))
            with misc.ddp_sync(G_ddp, round_idx == num_accumulation_rounds - 1):
                images = sid_sd_sampler(unet=G_ddp, latents=z, contexts=prompts, init_timesteps=init_timesteps,
                                        num_steps=num_steps, **sampler_kw)
                with misc.ddp_sync(fake_score_ddp, False):
                    y_fake = sid_sd_denoise(unet=fake_score_ddp, images=images, noise=noise, contexts=prompts,
                                            timesteps=timesteps, guidance_scale=cfg_eval_fake, **sampler_kw)
                    y_real = sid_sd_denoise(unet=true_score, images=images, noise=noise, contexts=prompts,
                                            timesteps=timesteps, guidance_scale=cfg_eval_real, **sampler_kw)
                    loss, loss_g_dev = _step.ops.lsg_loss(images, y_real, y_fake, alpha, loss_scaling_G / batch_gpu_total)
                    loss.backward()
        G_ddp.eval().requires_grad_(False)
        ema_target, beta = None, 0.0
        if ema_halflife_kimg > 0:                                                    # :553-565, fused into the step
            beta = _step.ema_beta(batch_size, cur_nimg, ema_halflife_kimg, ema_rampup_ratio)
            ema_target = G_ema.flat
        g_optimizer.step(grad_scale=inv_world, clip=fp16_clip, ema=ema_target, ema_beta=beta)   # :541-549

        # the two host reads of the iteration (:452, :535), after both phases have been enqueued
        loss_fake_score_print = float(loss_fake_dev[0].item())
        lossG_print = float(loss_g_dev[0].item())
        training_stats.report("fake_score_Loss/loss", loss_fake_score_print)
        training_stats.report("G_Loss/loss", lossG_print)
        if on_iteration is not None:
            on_iteration(dict(cur_nimg=cur_nimg, loss_fake=loss_fake_score_print, loss_G=lossG_print, G=G,
                              fake_score=fake_score, G_ema=G_ema))

        cur_nimg += batch_size
        done = cur_nimg >= total_kimg * 1000
        if (not done) and (cur_tick != 0) and (cur_nimg < tick_start_nimg + kimg_per_tick * 1000):
            continue

        # ---- tick (:573-588) ------------------------------------------------------------------------------------
        tick_end_time = time.time()
        r0 = training_stats.report0
        gpu = device.type == "cuda"
        fields = [
            f"tick {r0('Progress/tick', cur_tick):<5d}",
            f"kimg {r0('Progress/kimg', cur_nimg / 1e3):<9.1f}",
            f"time {dnnlib.util.format_time(r0('Timing/total_sec', tick_end_time - start_time)):<12s}",
            f"sec/tick {r0('Timing/sec_per_tick', tick_end_time - tick_start_time):<7.1f}",
            f"sec/kimg {r0('Timing/sec_per_kimg', (tick_end_time - tick_start_time) / (cur_nimg - tick_start_nimg) * 1e3):<7.2f}",
            f"maintenance {r0('Timing/maintenance_sec', maintenance_time):<6.1f}",
            f"cpumem {r0('Resources/cpu_mem_gb', psutil.Process(os.getpid()).memory_info().rss / 2**30):<6.2f}",
            f"gpumem {r0('Resources/peak_gpu_mem_gb', (torch.cuda.max_memory_allocated(device) if gpu else 0) / 2**30):<6.2f}",
            f"reserved {r0('Resources/peak_gpu_mem_reserved_gb', (torch.cuda.max_memory_reserved(device) if gpu else 0) / 2**30):<6.2f}",
            f"loss_fake_score {r0('fake_score_Loss/loss', loss_fake_score_print):<6.2f}",
            f"loss_G {r0('G_Loss/loss', lossG_print):<6.2f}",
        ]
        if gpu:
            torch.cuda.reset_peak_memory_stats()
        dist.print0(" ".join(fields))

        if (not done) and dist.should_stop():
            done = True
            dist.print0()
            dist.print0("Aborting...")

        if (snapshot_ticks is not None) and (done or cur_tick % snapshot_ticks == 0 or
                                             cur_tick in [2, 4, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100]):   # :597-652
            dist.print0("Exporting sample images...")
            if rank == 0:
                for num_steps_eval in [1, 2, 4]:
                    _export_grid(G_ema, grid_z, grid_c, enc, noise_scheduler, vae, init_timestep, device, resolution,
                                 dtype, num_steps, num_steps_eval,
                                 os.path.join(run_dir, f"fakes_{alpha:03f}_{cur_nimg//1000:06d}_{num_steps_eval:d}.png"),
                                 grid_size)
            data = dict(ema=copy.deepcopy(G_ema).eval().requires_grad_(False).cpu())
            if rank == 0:
                save_data(data=data, fname=os.path.join(run_dir, f"network-snapshot-{alpha:03f}-{cur_nimg//1000:06d}.pkl"))
            del data
            gc.collect()

        if (state_dump_ticks is not None) and (done or cur_tick % state_dump_ticks == 0) and cur_tick != 0 and rank == 0:
            dist.print0(f"saving checkpoint: training-state-{cur_nimg//1000:06d}.pt")                         # :654-656
            save_pt(pt=dict(fake_score=fake_score, G=G, G_ema=G_ema,
                            fake_score_optimizer_state=fake_score_optimizer.state_dict(),
                            g_optimizer_state=g_optimizer.state_dict()),
                    fname=os.path.join(run_dir, f"training-state-{cur_nimg//1000:06d}.pt"))

        training_stats.default_collector.update()                                                              # :659-662
        if rank == 0:
            append_line(jsonl_line=json.dumps(dict(training_stats.default_collector.as_dict(), timestamp=time.time())),
                        fname=os.path.join(run_dir, f"stats_{alpha:03f}.jsonl"))
        dist.update_progress(cur_nimg // 1000, total_kimg)

        cur_tick += 1
        tick_start_nimg = cur_nimg
        tick_start_time = time.time()
        maintenance_time = tick_start_time - tick_end_time
        if done:
            break

    dist.print0()
    dist.print0("Exiting...")
    return dict(G=G, G_ema=G_ema, fake_score=fake_score, cur_nimg=cur_nimg)
