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
                noise = sub_noise(i - 1) if callable(sub_noise) else sub_noise[i - 1]
            elif getattr(noise_scheduler, "rng", None) is not None:
                noise = noise_scheduler.rng.randn_like(z)
            else:
                noise = torch.randn_like(z)
            t_i = (init_timesteps * (1 - i / n)).to(torch.long)
            x_t = noise_scheduler.add_noise(d_x, noise, t_i)
            eps = unet(x_t, t_i, encoder_hidden_states=cond).sample
            d_x = noise_scheduler.pred_x0(eps, None, x_t, t_i, 1.0, True)
            noise_scheduler.burn(eps.shape, 1)          # the reference's scheduler.step() draw (:185), compat streams only
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
    if predict_x0:
        noise_scheduler.burn(eps_u.shape[1:], eps_u.shape[0])   # the per-sample scheduler.step() loop (:268-272)
    return noise_scheduler.pred_x0(eps_u, eps_c, x_t, timesteps, float(guidance_scale), predict_x0)


# ---- load_sd15 (reference :51-118) -------------------------------------------------------------------------------
class SyntheticTokenizer:
    """Offline stand-in with the tokenizer call signature the loop uses (:170, :221-236): a prompt becomes 77 ids
    derived from its bytes; '' maps to id 0 everywhere."""
    model_max_length = 77

    def __init__(self, vocab_size=4096):
        self.vocab_size = vocab_size

    def __call__(self, prompt, padding=None, max_length=None, truncation=None, return_tensors=None):
        import zlib
        from types import SimpleNamespace
        L = max_length or self.model_max_length
        rows = []
        for p in prompt:
            if p == "":
                rows.append([0] * L)
            else:
                h = zlib.crc32(p.encode())
                rows.append([1 + (h + 2654435761 * i) % (self.vocab_size - 1) for i in range(L)])
        return SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.long))


class SyntheticTextEncoder(torch.nn.Module):
    """token + position embedding lookup -> [b, 77, D] (random, fixed by `seed`): stands in for CLIPTextModel when no
    pretrained text encoder is on disk (synthetic prompt embeddings)."""

    def __init__(self, dim, vocab_size=4096, length=77, seed=1234567):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.tok = torch.nn.Parameter(torch.randn(vocab_size, dim, generator=g), requires_grad=False)
        self.pos = torch.nn.Parameter(torch.randn(length, dim, generator=g) * 0.5, requires_grad=False)

    def forward(self, input_ids):
        return ((self.tok[input_ids] + self.pos[: input_ids.shape[1]]) * (0.5 ** 0.5),)


def load_sd15(pretrained_model_name_or_path, pretrained_vae_model_name_or_path, device, weight_dtype,
              revision=None, variant=None, lora_config=None, enable_xformers=False, gradient_checkpointing=False):
    """-> (unet, vae, noise_scheduler, text_encoder, tokenizer), the reference's tuple (:51-52, 118).

    `pretrained_model_name_or_path` is a LOCAL diffusers pipeline folder (`unet/`, `scheduler/`, `tokenizer/`,
    `text_encoder/`; there is no hub access here) or `synthetic:<SD15|SD21_BASE|TINY>` for a random-init UNet with the
    synthetic tokenizer / text encoder.  The UNet is this package's (weights read by diffusers key name);
    fp16 / bf16 `weight_dtype` selects the bf16 tensor-core mode over fp32 master weights, fp32 the fp32-exact mode.
    The VAE is returned only when `diffusers` is importable (image decode is outside the hot path): else None."""
    import json
    import os
    from .. import unet as U
    from ..scheduler import DDPMScheduler
    from . import checkpoint
    del revision, variant, lora_config, pretrained_vae_model_name_or_path
    cd = torch.float32 if weight_dtype == torch.float32 else torch.bfloat16
    name = str(pretrained_model_name_or_path)
    vae = None
    if name.startswith("synthetic:"):
        cfg = getattr(U, name.split(":")[1])
        torch.manual_seed(0)
        unet = U.UNet2DConditionModel(cfg, compute_dtype=cd)
        tokenizer = SyntheticTokenizer()
        text_encoder = SyntheticTextEncoder(cfg.cross_attention_dim).to(device)
        sched = DDPMScheduler(device=device)
    else:
        if not os.path.isdir(name):
            raise FileNotFoundError("load_sd15: %r is not a local diffusers pipeline folder (no hub access)" % name)
        unet = checkpoint.load_unet(name, compute_dtype=cd)
        skw = {}
        sc = os.path.join(name, "scheduler", "scheduler_config.json")
        if os.path.exists(sc):
            j = json.load(open(sc))
            skw = dict(num_train_timesteps=j.get("num_train_timesteps", 1000), beta_start=j.get("beta_start", 0.00085),
                       beta_end=j.get("beta_end", 0.012), prediction_type=j.get("prediction_type", "epsilon"))
        sched = DDPMScheduler(device=device, **skw)
        from transformers import AutoTokenizer, CLIPTextModel
        tokenizer = AutoTokenizer.from_pretrained(name, subfolder="tokenizer", use_fast=False)
        text_encoder = CLIPTextModel.from_pretrained(name, subfolder="text_encoder").requires_grad_(False).to(device)
        try:
            from diffusers import AutoencoderKL
            vae = AutoencoderKL.from_pretrained(name, subfolder="vae").requires_grad_(False).to(device)
        except ImportError:
            vae = None
    unet.to(device)
    if enable_xformers:
        unet.enable_xformers_memory_efficient_attention()
    if gradient_checkpointing:
        unet.enable_gradient_checkpointing()
    return unet, vae, sched, text_encoder, tokenizer
