"""Weight / snapshot / training-state formats (sid_lsg_b200/training/checkpoint.py) on CPU: diffusers-keyed
safetensors round trip, config inference, state-dict snapshots, Adam state in torch.optim.Adam layout."""
import json
import os

import pytest
import torch

import oracle
import sid_lsg_b200 as S
from sid_lsg_b200.training import checkpoint as ck


def tiny():
    torch.manual_seed(3)
    return S.UNet2DConditionModel(S.TINY, compute_dtype=torch.float32)


def test_unet_safetensors_roundtrip_and_oracle_keys(tmp_path):
    net = tiny()
    folder = ck.save_unet(net, str(tmp_path / "unet"))
    assert sorted(os.listdir(folder)) == ["config.json", "diffusion_pytorch_model.safetensors"]
    sd = ck.read_unet_state_dict(str(tmp_path))            # pipeline folder containing unet/
    ref = oracle.UNet2DCondition(oracle.TINY)
    assert set(sd.keys()) == set(ref.state_dict().keys())  # the oracle is keyed like diffusers
    ref.load_state_dict(sd, strict=True)
    for k, v in net.state_dict().items():
        assert sd[k].is_contiguous() and torch.equal(sd[k], v.detach().float()), k
    again = ck.load_unet(folder, compute_dtype=torch.float32)
    assert again.cfg == S.TINY                              # heads come from config.json
    for (k, a), (_, b) in zip(again.state_dict().items(), net.state_dict().items()):
        assert torch.equal(a, b), k
    cfg = json.load(open(os.path.join(folder, "config.json")))
    assert cfg["attention_head_dim"] == [2, 2, 4, 4] and cfg["cross_attention_dim"] == 64


@pytest.mark.parametrize("cfg", [S.SD15, S.SD21_BASE, S.TINY_LINEAR])
def test_infer_config_from_shapes(cfg):
    with torch.device("meta"):
        net = S.UNet2DConditionModel(cfg)
    got = ck.infer_config(net.state_dict(), ck.diffusers_config(cfg))
    assert got == cfg
    if cfg in (S.SD15, S.SD21_BASE):                        # without a config.json the text width decides the heads
        assert ck.infer_config(net.state_dict()) == cfg


def test_network_snapshot_roundtrip(tmp_path):
    net = tiny()
    f = ck.save_network_snapshot(net, str(tmp_path / "network-snapshot-1.000000-000010.pt"))
    back = ck.load_network_snapshot(f, compute_dtype=torch.float32)
    assert back.cfg == S.TINY
    for (k, a), (_, b) in zip(back.state_dict().items(), net.state_dict().items()):
        assert torch.equal(a, b), k
    other = tiny()
    with torch.no_grad():
        for p in other.parameters():
            p.zero_()
    ck.load_network_snapshot(f, into=other)
    assert all(torch.equal(a, b) for a, b in zip(other.state_dict().values(), net.state_dict().values()))


def test_adam_state_matches_torch_layout():
    """flat second-moment bucket <-> torch.optim.Adam.state_dict(), including a channels_last conv weight."""
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(6, 4, 3, 3).contiguous(memory_format=torch.channels_last))
    b = torch.nn.Parameter(torch.randn(6))
    params = [w, b]
    offsets, total = [], 0
    for p in params:
        offsets.append(total)
        total += (p.numel() + 63) // 64 * 64
    bucket = torch.zeros(total)
    v_w, v_b = torch.rand(6, 4, 3, 3), torch.rand(6)
    torch.as_strided(bucket, w.shape, w.stride(), offsets[0]).copy_(v_w)     # physical layout of the parameter
    bucket[offsets[1]:offsets[1] + 6] = v_b
    sd = ck.adam_state_dict(params, bucket, offsets, step_count=7, lr=1e-6)
    assert torch.equal(sd["state"][0]["exp_avg_sq"], v_w) and torch.equal(sd["state"][1]["exp_avg_sq"], v_b)
    assert sd["state"][0]["exp_avg"].abs().sum() == 0 and float(sd["state"][0]["step"]) == 7
    opt = torch.optim.Adam(params, lr=1e-6, betas=(0.0, 0.999), eps=1e-8)
    opt.load_state_dict(sd)                                                   # what the reference does at :306-307
    assert torch.equal(opt.state[w]["exp_avg_sq"], v_w)
    bucket2 = torch.zeros(total)
    assert ck.load_adam_state_dict(opt.state_dict(), params, bucket2, offsets) == 7
    assert torch.equal(bucket2, bucket)
    empty = ck.adam_state_dict(params, bucket, offsets, step_count=0, lr=1e-6)
    assert empty["state"] == {} and empty["param_groups"][0]["params"] == [0, 1]
