"""Random draws of the distillation loop, in the reference's order and from the reference's seeds.

/root/reference/training/sid_training_loop.py:238-239 seeds numpy with `(seed * world + rank) % 2^31` and torch with
the first `randint(2^31)` of that numpy stream; z / noise / timesteps are then drawn ON THE DEVICE (:398-402, :413,
:479-484), the prompt-dropout mask on the CPU (:394), and diffusers' `DDPMScheduler.step` silently consumes one
`randn(model_output.shape)` per call (once per sampler sub-step, once PER SAMPLE in `predict_x0`).

`DrawStream` owns two torch generators seeded that way - one on the device the loop samples on, one on the CPU - and
exposes the draws by name.  With `compat=True` it also burns the scheduler's discarded draws, so the sequence of
numbers equals the reference's for the same seed, device type and shapes; with the CPU as sampling device both
generators are ONE object (on a CPU-only run the reference's CPU and "device" draws share the global generator), which
is how tests/golden/loop_*.pt (recorded from the reference's loop on CPU) is reproduced from the seed alone.
"""
import numpy as np
import torch


class DrawStream:
    def __init__(self, seed, rank=0, world=1, device="cuda", rng_device=None, compat=False):
        self.device = torch.device(device)
        self.rng_device = torch.device(rng_device) if rng_device is not None else self.device
        self.compat = compat
        self.np_seed = (seed * world + rank) % (1 << 31)
        self.torch_seed = int(np.random.RandomState(self.np_seed).randint(1 << 31))
        self.dev_gen = torch.Generator(device=self.rng_device).manual_seed(self.torch_seed)
        self.cpu_gen = self.dev_gen if self.rng_device.type == "cpu" else torch.Generator().manual_seed(self.torch_seed)
        self.burned = 0

    def _out(self, t):
        return t if t.device == self.device else t.to(self.device, non_blocking=True)

    def randn(self, shape, dtype=torch.float32):
        return self._out(torch.randn(tuple(shape), generator=self.dev_gen, device=self.rng_device, dtype=dtype))

    def randn_like(self, x):
        return self.randn(x.shape, torch.float32)

    def randint(self, low, high, shape):
        return self._out(torch.randint(low, high, tuple(shape), generator=self.dev_gen, device=self.rng_device,
                                       dtype=torch.long))

    def rand_cpu(self, n):
        return torch.rand(n, generator=self.cpu_gen)

    def burn(self, shape, count=1):
        """the draws diffusers' scheduler.step() makes and throws away (t > 0); no-op unless compat."""
        if not self.compat:
            return
        for _ in range(count):
            torch.randn(tuple(shape), generator=self.dev_gen, device=self.rng_device, dtype=torch.float32)
            self.burned += 1
