"""One SiD-LSG iteration restated on CPU (oracle; TEST INFRASTRUCTURE).

Follows, behaviour for behaviour:
  sampler          /root/reference/training/sid_sd_util.py:176-185
  denoise (CFG,x0) /root/reference/training/sid_sd_util.py:242-274
  f_psi phase      /root/reference/training/sid_training_loop.py:389-462
  G_theta phase    /root/reference/training/sid_training_loop.py:468-549
  EMA              /root/reference/training/sid_training_loop.py:553-565
  Adam settings    /root/reference/sid_train.py:219-226

Differences that are deliberate: the text front end (tokenizer + CLIP) is replaced by injected
prompt embeddings (SURVEY.md §8d synthetic inputs), and z / noise / timesteps are injected rather
than drawn (RNG order is implementation-defined upstream).
"""
import torch


def sampler(unet, sched, z, cond_emb, init_timesteps, num_steps=1, sub_noise=None):
    """sid_sd_util.py:176-185 (train_sampler=True). sub_noise[i-1] replaces randn_like for i>=1."""
    d_x = torch.zeros_like(z)
    latents = z
    for i in range(num_steps):
        noise = z if i == 0 else sub_noise[i - 1]
        t_i = (init_timesteps * (1 - i / num_steps)).to(torch.long)
        latents = sched.add_noise(d_x, noise, t_i).to(torch.float32)
        eps = unet(latents, t_i, encoder_hidden_states=cond_emb).sample.to(torch.float32)
        d_x = sched.step(eps, t_i[0], latents).pred_original_sample.to(torch.float32)
    return d_x


def denoise(unet, sched, images, noise, cond_emb, uncond_emb, timesteps, predict_x0=True, guidance_scale=1.0):
    """sid_sd_util.py:242-274."""
    latents = sched.add_noise(images, noise, timesteps)
    if guidance_scale == 1:
        eps = unet(latents, timesteps, encoder_hidden_states=cond_emb).sample.to(torch.float32)
    else:
        emb = torch.cat([uncond_emb, cond_emb])
        t2 = torch.cat([timesteps, timesteps])
        x2 = torch.cat([latents] * 2)
        out = unet(x2, t2, encoder_hidden_states=emb).sample.to(torch.float32)
        e_u, e_c = out.chunk(2)
        eps = e_u + guidance_scale * (e_c - e_u)
    if predict_x0:
        xs = [sched.step(n, t, x).pred_original_sample.to(torch.float32)
              for n, t, x in zip(eps, timesteps, latents.to(torch.float32))]
        return torch.stack(xs).to(torch.float32)
    return eps.to(torch.float32)


def fake_score_loss(noise_fake, noise, loss_scaling, batch_gpu_total):
    """sid_training_loop.py:423-445 (epsilon prediction)."""
    nan_mask = torch.isnan(noise_fake).flatten(start_dim=1).any(dim=1)
    if nan_mask.any():
        keep = ~nan_mask
        noise_fake = noise_fake[keep]
        noise = noise[keep]
    loss = (noise_fake - noise) ** 2
    return loss.sum().mul(loss_scaling / batch_gpu_total), len(noise)


def generator_loss(images, y_real, y_fake, alpha, loss_scaling_G, batch_gpu_total):
    """sid_training_loop.py:508-530."""
    nan_mask = (torch.isnan(images).flatten(start_dim=1).any(dim=1)
                | torch.isnan(y_real).flatten(start_dim=1).any(dim=1)
                | torch.isnan(y_fake).flatten(start_dim=1).any(dim=1))
    if nan_mask.any():
        keep = ~nan_mask
        images, y_real, y_fake = images[keep], y_real[keep], y_fake[keep]
    with torch.no_grad():
        w = abs(images.to(torch.float32) - y_real.to(torch.float32)).mean(dim=[1, 2, 3], keepdim=True).clip(min=0.00001)
    if alpha == 1:
        loss = (y_real - y_fake) * (y_fake - images) / w
    else:
        loss = (y_real - y_fake) * ((y_real - images) - alpha * (y_real - y_fake)) / w
    return loss.sum().mul(loss_scaling_G / batch_gpu_total), len(y_real)


def sanitize_grads(params, clip_value=None):
    """sid_training_loop.py:458-460, 541-547."""
    for p in params:
        if p.grad is not None:
            torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
    if clip_value is not None:
        torch.nn.utils.clip_grad_value_(params, clip_value)


def make_optimizer(params, lr=1e-6, eps=1e-8):
    """sid_train.py:219-226: Adam(betas=[0.0, 0.999])."""
    return torch.optim.Adam(params, lr=lr, betas=(0.0, 0.999), eps=eps)


def ema_beta(batch_size, cur_nimg, ema_halflife_kimg, ema_rampup_ratio=0.05):
    """sid_training_loop.py:553-558."""
    ema_halflife_nimg = ema_halflife_kimg * 1000
    if ema_rampup_ratio is not None:
        ema_halflife_nimg = min(ema_halflife_nimg, cur_nimg * ema_rampup_ratio)
    return 0.5 ** (batch_size / max(ema_halflife_nimg, 1e-8))


def ema_update(g_ema, g, beta):
    """sid_training_loop.py:560-563."""
    with torch.no_grad():
        for p_ema, p in zip(g_ema.parameters(), g.parameters()):
            p_ema.copy_(p.detach().lerp(p_ema, beta))


def fake_score_phase(G, fake_score, sched, opt_f, mb, *, kappa1, init_timestep=625, num_steps=1,
                     loss_scaling=1.0, batch_gpu_total=None):
    """One f_psi update over the micro-batches in `mb` (each a dict with z, noise, t, cond, uncond
    [, sub_noise]).  Returns the last micro-batch's loss value (what the reference prints)."""
    fake_score.train().requires_grad_(True)
    opt_f.zero_grad(set_to_none=True)
    total = batch_gpu_total or sum(m["z"].shape[0] for m in mb)
    loss = None
    for m in mb:
        b = m["z"].shape[0]
        init_t = init_timestep * torch.ones((b,), dtype=torch.long)
        with torch.no_grad():
            images = sampler(G, sched, m["z"], m["cond"], init_t, num_steps, m.get("sub_noise"))
        eps = denoise(fake_score, sched, images, m["noise"], m["cond"], m["uncond"], m["t"],
                      predict_x0=False, guidance_scale=kappa1)
        loss, n = fake_score_loss(eps, m["noise"], loss_scaling, total)
        if n > 0:
            loss.backward()
    fake_score.eval().requires_grad_(False)
    sanitize_grads(list(fake_score.parameters()))
    opt_f.step()
    return float(loss.item())


def generator_phase(G, fake_score, true_score, sched, opt_g, mb, *, kappa2, kappa4, alpha=1.0,
                    init_timestep=625, num_steps=1, loss_scaling_G=1.0, batch_gpu_total=None, fp16=False,
                    return_images=False):
    G.train().requires_grad_(True)
    opt_g.zero_grad(set_to_none=True)
    total = batch_gpu_total or sum(m["z"].shape[0] for m in mb)
    loss = None
    imgs = []
    for m in mb:
        b = m["z"].shape[0]
        init_t = init_timestep * torch.ones((b,), dtype=torch.long)
        images = sampler(G, sched, m["z"], m["cond"], init_t, num_steps, m.get("sub_noise"))
        y_fake = denoise(fake_score, sched, images, m["noise"], m["cond"], m["uncond"], m["t"], guidance_scale=kappa2)
        y_real = denoise(true_score, sched, images, m["noise"], m["cond"], m["uncond"], m["t"], guidance_scale=kappa4)
        loss, n = generator_loss(images, y_real, y_fake, alpha, loss_scaling_G, total)
        if n > 0:
            loss.backward()
        imgs.append(images.detach())
    G.eval().requires_grad_(False)
    sanitize_grads(list(G.parameters()), clip_value=1 if fp16 else None)
    opt_g.step()
    if return_images:
        return float(loss.item()), torch.cat(imgs)
    return float(loss.item())


def iteration(G, G_ema, fake_score, true_score, sched, opt_f, opt_g, mb_f, mb_g, *, kappa, alpha=1.0,
              batch_size=None, cur_nimg=0, ema_halflife_kimg=50, num_steps=1, init_timestep=625,
              loss_scaling=1.0, loss_scaling_G=1.0):
    """sid_training_loop.py:383-567 with kappa1=kappa2=kappa3=kappa4=kappa."""
    lf = fake_score_phase(G, fake_score, sched, opt_f, mb_f, kappa1=kappa, init_timestep=init_timestep,
                          num_steps=num_steps, loss_scaling=loss_scaling)
    lg = generator_phase(G, fake_score, true_score, sched, opt_g, mb_g, kappa2=kappa, kappa4=kappa, alpha=alpha,
                         init_timestep=init_timestep, num_steps=num_steps, loss_scaling_G=loss_scaling_G)
    bs = batch_size or sum(m["z"].shape[0] for m in mb_g)
    if G_ema is not None and ema_halflife_kimg > 0:
        ema_update(G_ema, G, ema_beta(bs, cur_nimg, ema_halflife_kimg))
    return lf, lg


def synth_microbatch(b, cfg, seed, dropout=False, num_steps=1, tmin=20, tmax=980):
    """Synthetic inputs of SURVEY.md §8d: randn embeddings, fixed uncond, z/noise, t in [tmin,tmax)."""
    g = torch.Generator().manual_seed(seed)
    hw = cfg.sample_size
    d = cfg.cross_attention_dim
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
    return m
