"""DDPM schedule restated from diffusers==0.27.2 schedulers/scheduling_ddpm.py behaviour
(SURVEY.md App. A-3); called by the reference at
/root/reference/training/sid_sd_util.py:182-185,242-244,262,270.

TEST INFRASTRUCTURE (oracle).  fp32 CPU, plain torch.
"""
import math
from types import SimpleNamespace

import torch


class _StepOutput:
    def __init__(self, prev_sample, pred_original_sample):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class DDPMSchedule:
    """`scaled_linear` betas in [0.00085, 0.012], 1000 steps, epsilon prediction, no clipping."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 prediction_type="epsilon"):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                               dtype=torch.float32) ** 2
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.config = SimpleNamespace(prediction_type=prediction_type,
                                      num_train_timesteps=num_train_timesteps,
                                      clip_sample=False, variance_type="fixed_small")
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    # -- helpers -----------------------------------------------------------------
    def _coeffs(self, like, t):
        ac = self.alphas_cumprod.to(device=like.device, dtype=like.dtype)
        t = t.to(like.device)
        sa = ac[t] ** 0.5
        sb = (1 - ac[t]) ** 0.5
        sa = sa.flatten()
        sb = sb.flatten()
        while sa.dim() < like.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa, sb

    # -- diffusers protocol ------------------------------------------------------
    def add_noise(self, original_samples, noise, timesteps):
        sa, sb = self._coeffs(original_samples, timesteps)
        return sa * original_samples + sb * noise

    def get_velocity(self, sample, noise, timesteps):
        sa, sb = self._coeffs(sample, timesteps)
        return sa * noise - sb * sample

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        """Single (scalar) timestep, as the reference always calls it
        (sid_sd_util.py:185 passes init_timesteps_i[0]; :270 loops per sample)."""
        t = int(timestep)
        prev_t = t - 1  # num_inference_steps is None -> prev = t-1
        ac = self.alphas_cumprod
        alpha_prod_t = ac[t].to(sample.device)
        alpha_prod_t_prev = ac[prev_t].to(sample.device) if prev_t >= 0 else self.one.to(sample.device)
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        current_alpha_t = alpha_prod_t / alpha_prod_t_prev
        current_beta_t = 1 - current_alpha_t
        if self.config.prediction_type == "epsilon":
            x0 = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        elif self.config.prediction_type == "v_prediction":
            x0 = alpha_prod_t ** 0.5 * sample - beta_prod_t ** 0.5 * model_output
        else:
            raise ValueError(self.config.prediction_type)
        c0 = alpha_prod_t_prev ** 0.5 * current_beta_t / beta_prod_t
        ct = current_alpha_t ** 0.5 * beta_prod_t_prev / beta_prod_t
        prev = c0 * x0 + ct * sample
        if t > 0:
            # the reference discards prev_sample, but this draw advances the RNG
            z = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                            dtype=model_output.dtype)
            var = (beta_prod_t_prev / beta_prod_t * current_beta_t).clamp(min=1e-20)
            prev = prev + var ** 0.5 * z
        return _StepOutput(prev, x0)


def compute_snr(schedule, timesteps):
    """diffusers.training_utils.compute_snr: alpha^2/sigma^2 = acp/(1-acp), shape [b]."""
    ac = schedule.alphas_cumprod.to(timesteps.device)[timesteps].float()
    return ac / (1 - ac)


def timestep_embedding(timesteps, dim, max_period=10000.0):
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    arg = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)
