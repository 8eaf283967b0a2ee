"""Text conditioning front end of the distillation step.

The reference tokenises and runs the CLIP text encoder inside EVERY sampler / denoise call
(/root/reference/training/sid_sd_util.py:170-172, 221-240): per micro-batch that is five encoder passes over the same
prompts plus three over the constant '' prompt.  `PromptEncoder` produces the same tensors once:

* `encode(prompts)` -> `PromptBatch(cond, uncond)`: one tokenizer + encoder pass for the batch (exactly the reference's
  call: padding='max_length', max_length=tokenizer.model_max_length, truncation=True), and the '' embedding computed
  ONCE per (encoder, length) and expanded to the batch;
* `dropout(batch, p, generator)`: the f_psi phase's 10 % prompt dropout (:393-396) applied to embeddings instead of
  strings (replacing a prompt by '' == replacing its embedding row by the '' embedding);
* optional LRU cache keyed by the prompt string (Aesthetic6+ prompts repeat across epochs).

`sid_sd_sampler` / `sid_sd_denoise` (training/sid_sd_util.py) take the resulting PromptBatch in place of the list of
strings, so the UNet kernels never wait for redundant encoder work.  Host-side logic only; the encoder itself is whatever
`transformers.CLIPTextModel` the caller loaded (outside the hot path, SURVEY §8 "out of scope").
"""
from collections import OrderedDict

import torch

from .sid_sd_util import PromptBatch


class PromptEncoder:
    def __init__(self, tokenizer, text_encoder, device=None, cache_size=0, out_dtype=None):
        self.tokenizer, self.text_encoder = tokenizer, text_encoder
        self.device = device if device is not None else next(text_encoder.parameters()).device
        self.out_dtype = out_dtype
        self._uncond = {}                      # max_length -> [1, L, D]
        self._cache = OrderedDict() if cache_size > 0 else None
        self._cache_size = cache_size
        self.encoder_calls = 0                 # text-encoder forward passes issued (tests / logging)

    # -- the reference's two calls ------------------------------------------------------------------------------
    def _tokenize(self, prompts, max_length):
        return self.tokenizer(list(prompts), padding="max_length", max_length=max_length, truncation=True,
                              return_tensors="pt").input_ids

    @torch.no_grad()
    def _run(self, input_ids):
        self.encoder_calls += 1
        out = self.text_encoder(input_ids.to(self.device))[0]
        return out.to(self.out_dtype) if self.out_dtype is not None else out

    def uncond(self, max_length=None):
        """embedding of '' ([1, L, D]); computed once (the reference recomputes it per call and per batch row)."""
        L = max_length or self.tokenizer.model_max_length
        if L not in self._uncond:
            self._uncond[L] = self._run(self._tokenize([""], L))
        return self._uncond[L]

    def encode(self, prompts):
        prompts = list(prompts)
        L = self.tokenizer.model_max_length
        if self._cache is None:
            cond = self._run(self._tokenize(prompts, L))
        else:
            missing = [p for p in dict.fromkeys(prompts) if p not in self._cache]
            if missing:
                emb = self._run(self._tokenize(missing, L))
                for p, e in zip(missing, emb):
                    self._cache[p] = e
            for p in prompts:
                self._cache.move_to_end(p)
            cond = torch.stack([self._cache[p] for p in prompts])
            while len(self._cache) > self._cache_size:
                self._cache.popitem(last=False)
        un = self.uncond(cond.shape[1])
        return PromptBatch(cond, un.expand(len(prompts), -1, -1))

    @staticmethod
    def dropout(batch, p=0.1, generator=None):
        """f_psi phase (:393-396): each prompt becomes '' with probability p -> its embedding row becomes the ''
        embedding.  Returns (PromptBatch, bool mask of dropped rows)."""
        b = batch.cond.shape[0]
        drop = torch.rand(b, generator=generator) < p
        cond = torch.where(drop.to(batch.cond.device)[:, None, None], batch.uncond, batch.cond)
        return PromptBatch(cond, batch.uncond), drop
