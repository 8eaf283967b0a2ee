"""sid_lsg_b200: the SiD-LSG distillation step (SD UNet forwards/backwards, DDPM algebra, LSG losses, fused
optimiser) on hand-written sm_100a CUDA behind a C ABI (include/sidlsg.h).  No CPU or PyTorch fallback:
importing the ops without the built library raises."""
from ._lib import lib, LIB_PATH  # noqa: F401
from .unet import UNet2DConditionModel, UNetConfig, SD15, SD21_BASE, TINY, TINY_LINEAR  # noqa: F401
from .scheduler import DDPMScheduler  # noqa: F401
from .params import FlatParams, FlatAdam  # noqa: F401
from .ddp import FlatDDP  # noqa: F401
from . import torch_utils, dnnlib  # noqa: F401
from .training.sid_sd_util import sid_sd_sampler, sid_sd_denoise, PromptBatch, load_sd15  # noqa: F401
from .training.sid_training_loop import training_loop  # noqa: F401
from .training.draws import DrawStream  # noqa: F401
from .training.step import SiDLSGStep, GraphedIteration, synth_microbatch, device_microbatch, ema_beta  # noqa: F401
from .training.prompts import PromptEncoder  # noqa: F401
from .training import checkpoint  # noqa: F401
from .training.checkpoint import load_unet, save_unet, save_network_snapshot, load_network_snapshot  # noqa: F401
from .training.checkpoint import save_training_state, load_training_state  # noqa: F401

__version__ = "0.1.0"
