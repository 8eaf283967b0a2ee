"""CPU oracle for the SiD-LSG distillation step.  TEST INFRASTRUCTURE ONLY.

Pure-PyTorch fp32 restatement of the reference's per-iteration hot path
(/root/reference/training/sid_training_loop.py:383-565 and
/root/reference/training/sid_sd_util.py:163-274) and of the un-vendored
diffusers==0.27.2 arithmetic it calls (UNet2DConditionModel, DDPMScheduler).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (sid_lsg_b200) never does.

PARITY PIN STATUS: the reference ships no tests or golden vectors and diffusers is
not installable here, so the UNet arithmetic is "parity unpinned" against upstream
(validated structurally: exact 859,520,964 / 865,910,724 parameter counts, 686
tensors, diffusers state-dict keys).  The *glue* (sampler, CFG denoise, x0
conversion) IS pinned: tests/golden/make_golden.py imports the reference's own
training/sid_sd_util.py (with stub `diffusers`/`transformers` modules backed by this
oracle) and the committed fixtures hold its outputs.
"""
from .scheduler import DDPMSchedule  # noqa: F401
from .unet import UNetConfig, UNet2DCondition, SD15, SD21_BASE, TINY  # noqa: F401
