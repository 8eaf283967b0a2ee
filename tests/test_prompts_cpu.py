"""PromptEncoder (training/prompts.py) against the reference's per-call tokenise + encode sequence
(sid_sd_util.py:170-172, 221-240) with a random-init CLIP text encoder: same tensors, one encoder pass per batch."""
import pytest
import torch

from sid_lsg_b200.training.prompts import PromptEncoder
from sid_lsg_b200.training.sid_sd_util import _embed

transformers = pytest.importorskip("transformers")


class ToyTokenizer:
    """whitespace tokenizer with the call signature the reference uses (padding='max_length', truncation, 'pt')."""
    model_max_length = 16

    def __call__(self, prompt, padding="max_length", max_length=None, truncation=True, return_tensors="pt"):
        L = max_length or self.model_max_length
        rows = []
        for s in prompt:
            ids = [1] + [3 + (sum(map(ord, w)) % 90) for w in s.split()][:L - 2] + [2]
            rows.append(ids + [0] * (L - len(ids)))
        return type("Enc", (), {"input_ids": torch.tensor(rows, dtype=torch.long)})()


@pytest.fixture(scope="module")
def clip():
    cfg = transformers.CLIPTextConfig(vocab_size=100, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                      num_attention_heads=2, max_position_embeddings=16)
    torch.manual_seed(0)
    return transformers.CLIPTextModel(cfg).eval()


def test_matches_reference_call_sequence(clip):
    tok = ToyTokenizer()
    prompts = ["a red fox", "", "two cats on a sofa", "a red fox"]
    ref_cond, ref_uncond = _embed(prompts, "cpu", clip, tok, True)      # what sid_sd_denoise computes per call
    enc = PromptEncoder(tok, clip)
    batch = enc.encode(prompts)
    assert torch.equal(batch.cond, ref_cond) and torch.allclose(batch.uncond, ref_uncond, atol=1e-6)
    assert enc.encoder_calls == 2                                        # prompts once, '' once
    enc.encode(prompts[:2])
    assert enc.encoder_calls == 3                                        # '' is cached


def test_cache_and_dropout(clip):
    tok = ToyTokenizer()
    enc = PromptEncoder(tok, clip, cache_size=8)
    a = enc.encode(["x y", "z", "x y"])
    calls = enc.encoder_calls
    b = enc.encode(["z", "x y"])
    assert enc.encoder_calls == calls                                    # all served from the cache
    assert torch.equal(b.cond[0], a.cond[1]) and torch.equal(b.cond[1], a.cond[0])
    g = torch.Generator().manual_seed(1)
    dropped, mask = PromptEncoder.dropout(a, p=0.5, generator=g)
    for i in range(3):
        want = a.uncond[i] if mask[i] else a.cond[i]
        assert torch.equal(dropped.cond[i], want)
