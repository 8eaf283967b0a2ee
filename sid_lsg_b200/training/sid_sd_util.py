"""`sid_sd_sampler` / `sid_sd_denoise` with the reference's signatures
(/root/reference/training/sid_sd_util.py:163-164, 214-215) on the sm_100a kernels.

`contexts` is either a list[str] (then `tokenizer` / `text_encoder` are called exactly as the reference does,
:170-172, :221-240) or a `PromptBatch` of precomputed embeddings (the synthetic-embedding metric, SURVEY.md §8d;
the '' embedding is a constant the reference recomputes on every call).
"""
from collections import namedtuple

import torch

from .. import ops

PromptBatch = namedtuple("PromptBatch", ["cond", "uncond"])  # [b,77,D] each; len() == 2, so use .cond.shape[0]


def _embed(contexts, device, text_encoder, tokenizer, want_uncond):
    if isinstance(contexts, PromptBatch):
        return contexts.cond, (contexts.uncond if want_uncond else None)
    if torch.is_tensor(contexts):
        if want_uncond:
            raise ValueError("classifier-free guidance needs PromptBatch(cond, uncond) or string prompts")
        return contexts, None
    prompt = list(contexts)
    ti = tokenizer(prompt, padding="max_length", max_length=tokenizer.model_max_length, truncation=True,
                   return_tensors="pt")
    with torch.no_grad():
        cond = text_encoder(ti.input_ids.to(device))[0]
        uncond = None
        if want_uncond:
            ui = tokenizer([""] * len(prompt), padding="max_length", max_length=ti.input_ids.shape[-1],
                           return_tensors="pt")
            uncond = text_encoder(ui.input_ids.to(device))[0]
    return cond, uncond


def sid_sd_sampler(unet, latents, contexts, init_timesteps, noise_scheduler, text_encoder=None, tokenizer=None,
                   resolution=512, dtype=torch.float16, return_images=False, vae=None, guidance_scale=1, num_steps=1,
                   train_sampler=True, num_steps_eval=1, sub_noise=None):
    """Reference :176-196.  `dtype` is accepted for signature compatibility; the UNet computes in its own
    compute_dtype and returns fp32.  `sub_noise` (list, optional) injects the i>=1 sub-step noises that the
    reference draws with randn_like (tests need explicit draws: RNG order is implementation-defined, App. B-3)."""
    cond, _ = _embed(contexts, latents.device, text_encoder, tokenizer, False)
    n = num_steps if train_sampler else num_steps_eval
    z = latents
    d_x = None
    ctx = torch.enable_grad() if (train_sampler and torch.is_grad_enabled()) else torch.no_grad()
    with ctx:
        for i in range(n):
            if i == 0:
                noise = z
            elif sub_noise is not None:
                noise = sub_noise[i - 1]
            else:
                noise = torch.randn_like(z)
            t_i = (init_timesteps * (1 - i / n)).to(torch.long)
            x_t = noise_scheduler.add_noise(d_x, noise, t_i)
            eps = unet(x_t, t_i, encoder_hidden_states=cond).sample
            d_x = noise_scheduler.pred_x0(eps, None, x_t, t_i, 1.0, True)
    if return_images:
        if vae is None:
            raise ValueError("return_images=True needs a VAE (outside the distillation hot path)")
        return vae.decode(d_x / vae.config.scaling_factor, return_dict=False)[0].to(torch.float32)
    return d_x


def sid_sd_denoise(unet, images, noise, contexts, timesteps, noise_scheduler, text_encoder=None, tokenizer=None,
                   resolution=512, dtype=torch.float16, predict_x0=True, guidance_scale=1):
    """Reference :242-274: add_noise -> (batched CFG) UNet -> eps [-> x0], one launch per algebra step."""
    cond, uncond = _embed(contexts, images.device, text_encoder, tokenizer, guidance_scale != 1)
    x_t = noise_scheduler.add_noise(images, noise, timesteps)
    if guidance_scale == 1:
        eps_u = unet(x_t, timesteps, encoder_hidden_states=cond).sample
        eps_c = None
    else:
        emb = torch.cat([uncond, cond])
        t2 = torch.cat([timesteps, timesteps])
        out = unet(torch.cat([x_t, x_t]), t2, encoder_hidden_states=emb).sample
        eps_u, eps_c = out.chunk(2)
    return noise_scheduler.pred_x0(eps_u, eps_c, x_t, timesteps, float(guidance_scale), predict_x0)
