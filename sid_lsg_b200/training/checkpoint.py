"""Weight and training-state formats either side of the distillation step.

What the reference reads and writes, and what this module does about it:

* **Pretrained UNet weights** - `UNet2DConditionModel.from_pretrained(name, subfolder="unet")`
  (/root/reference/training/sid_sd_util.py:73-79) reads `unet/diffusion_pytorch_model.safetensors` (or `.bin`) keyed by
  the diffusers state-dict names.  `read_unet_state_dict` / `load_unet` read exactly those files by name into
  `sid_lsg_b200.UNet2DConditionModel` (whose parameter names ARE the diffusers keys), `save_unet` writes them back
  (plus a diffusers `config.json`), so weights move both ways without diffusers being installed.
* **`network-snapshot-*.pkl`** - `pickle.dump({'ema': G_ema_module})` (training/sid_training_loop.py:641-652, read by
  generate_onestep.py:247-248).  A pickled diffusers module can only be unpickled where diffusers is importable, so
  snapshots here are state-dict based (`save_network_snapshot`: `{'ema': state_dict, 'config': ...}` through
  `torch.save`); `load_network_snapshot` also accepts the reference's pickle when it unpickles, via `.state_dict()`.
* **`training-state-*.pt`** - `torch.save(dict(fake_score=, G=, G_ema=, fake_score_optimizer_state=,
  g_optimizer_state=))` (:654-656), restored at :296-310.  `save_training_state` / `load_training_state` keep the same
  five keys; the networks are state dicts and the two optimiser entries are in `torch.optim.Adam.state_dict()`
  layout (`adam_state_dict`: per-parameter `step`, `exp_avg`, `exp_avg_sq` + one param group, indexed in
  `parameters()` order - which unet.py keeps equal to diffusers 0.27.2's registration order, see
  tests/test_surface_cpu.py::test_parameter_order_is_diffusers_order), so the reference's
  `fake_score_optimizer.load_state_dict(...)` accepts them and its own dumps load here (`load_adam_state_dict`, which
  checks every state tensor's shape against its parameter).

Everything is host-side I/O; tensors are moved with plain copies (the flat buckets are the only device state).
"""
import json
import os
import pickle

import torch

from ..unet import UNet2DConditionModel, UNetConfig, SD15, SD21_BASE

SNAPSHOT_FORMAT = "sid_lsg_b200/state_dict/v1"
_WEIGHT_FILES = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors",
                 "diffusion_pytorch_model.bin")


# ---- diffusers UNet weights -----------------------------------------------------------------------------------
def _resolve_weight_file(path):
    if os.path.isfile(path):
        return path
    for sub in ("", "unet"):
        for name in _WEIGHT_FILES:
            cand = os.path.join(path, sub, name)
            if os.path.isfile(cand):
                return cand
    raise FileNotFoundError("no diffusers UNet weight file (%s) under %r" % (", ".join(_WEIGHT_FILES), path))


def read_unet_state_dict(path):
    """path: a `.safetensors` / `.bin` / `.pt` file, a diffusers `unet/` folder, or a pipeline folder containing
    `unet/`.  Returns {diffusers key: CPU tensor} (dtype as stored)."""
    f = _resolve_weight_file(path)
    if f.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(f, device="cpu")
    sd = torch.load(f, map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and not any(torch.is_tensor(v) for v in sd.values()):
        sd = sd["state_dict"]
    return sd


def infer_config(state_dict, config_json=None):
    """UNetConfig from the tensor shapes (+ the head counts from a diffusers config.json when given: the number of
    heads is not recoverable from the weights; without a config, 768-wide text -> SD1.5 heads, 1024 -> SD2.1)."""
    ch = [state_dict["conv_in.weight"].shape[0]]
    i = 1
    while "down_blocks.%d.resnets.0.conv1.weight" % i in state_dict:
        ch.append(state_dict["down_blocks.%d.resnets.0.conv1.weight" % i].shape[0])
        i += 1
    layers = 0
    while "down_blocks.0.resnets.%d.conv1.weight" % layers in state_dict:
        layers += 1
    kv = state_dict["down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"]
    cross = kv.shape[1]
    linear = state_dict["down_blocks.0.attentions.0.proj_in.weight"].dim() == 2
    groups = 32
    heads = None
    sample = 64
    eps = 1e-5
    if config_json is not None:
        cfg = config_json if isinstance(config_json, dict) else json.load(open(config_json))
        hd = cfg.get("num_attention_heads") or cfg.get("attention_head_dim", 8)
        heads = tuple(hd) if isinstance(hd, (list, tuple)) else (hd,) * len(ch)
        groups = cfg.get("norm_num_groups", groups)
        sample = cfg.get("sample_size", sample)
        eps = cfg.get("norm_eps", eps)
    if heads is None:
        heads = SD21_BASE.num_heads if cross == 1024 else SD15.num_heads
        if len(heads) != len(ch):
            heads = (heads[0],) * len(ch)
    return UNetConfig(in_channels=state_dict["conv_in.weight"].shape[1], out_channels=state_dict["conv_out.weight"].shape[0],
                      block_out_channels=tuple(ch), layers_per_block=layers, cross_attention_dim=cross,
                      num_heads=tuple(heads), norm_num_groups=groups, norm_eps=eps, use_linear_projection=linear,
                      sample_size=sample)


def load_unet(path_or_state_dict, cfg=None, compute_dtype=torch.bfloat16, device=None):
    """-> UNet2DConditionModel with the checkpoint's weights (strict by name and shape).  `device` given: moved there
    and re-homed into the flat buckets (ready for the kernels); otherwise a CPU module (state-dict work only)."""
    if isinstance(path_or_state_dict, dict):
        sd, folder = path_or_state_dict, None
    else:
        f = _resolve_weight_file(path_or_state_dict)
        sd, folder = read_unet_state_dict(f), os.path.dirname(f)
    if cfg is None:
        cj = os.path.join(folder, "config.json") if folder else None
        cfg = infer_config(sd, cj if (cj and os.path.isfile(cj)) else None)
    net = UNet2DConditionModel(cfg, compute_dtype=compute_dtype)
    net.load_state_dict({k: v.to(torch.float32) for k, v in sd.items()}, strict=True)
    if device is not None:
        net.to(device)
        net.flatten_()
    return net


def diffusers_config(cfg: UNetConfig):
    """the `unet/config.json` diffusers 0.27 writes for this architecture."""
    n = len(cfg.block_out_channels)
    hd = list(cfg.num_heads) if len(set(cfg.num_heads)) > 1 else cfg.num_heads[0]
    return {
        "_class_name": "UNet2DConditionModel", "_diffusers_version": "0.27.2", "act_fn": "silu",
        "attention_head_dim": hd, "block_out_channels": list(cfg.block_out_channels), "center_input_sample": False,
        "cross_attention_dim": cfg.cross_attention_dim,
        "down_block_types": ["CrossAttnDownBlock2D"] * (n - 1) + ["DownBlock2D"], "downsample_padding": 1,
        "flip_sin_to_cos": True, "freq_shift": 0, "in_channels": cfg.in_channels,
        "layers_per_block": cfg.layers_per_block, "mid_block_scale_factor": 1, "norm_eps": cfg.norm_eps,
        "norm_num_groups": cfg.norm_num_groups, "out_channels": cfg.out_channels, "sample_size": cfg.sample_size,
        "up_block_types": ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * (n - 1),
        "use_linear_projection": cfg.use_linear_projection,
    }


def plain_state_dict(net):
    """diffusers-keyed CPU fp32 state dict with dense NCHW conv weights (what safetensors / the reference expect;
    inside the buckets 3x3 weights are channels_last)."""
    out = {}
    for k, v in net.state_dict().items():
        out[k] = v.detach().to("cpu", torch.float32).contiguous()
    return out


def save_unet(net, folder, dtype=torch.float32):
    """writes `folder/diffusion_pytorch_model.safetensors` + `folder/config.json` (a diffusers `unet/` folder)."""
    from safetensors.torch import save_file
    os.makedirs(folder, exist_ok=True)
    sd = {k: v.to(dtype).contiguous() for k, v in plain_state_dict(net).items()}
    save_file(sd, os.path.join(folder, _WEIGHT_FILES[0]), metadata={"format": "pt"})
    with open(os.path.join(folder, "config.json"), "w") as f:
        json.dump(diffusers_config(net.cfg), f, indent=2)
    return folder


# ---- network snapshots ({'ema': ...}) --------------------------------------------------------------------------
def save_network_snapshot(G_ema, fname):
    """state-dict counterpart of sid_training_loop.py:641-652."""
    data = {"format": SNAPSHOT_FORMAT, "ema": plain_state_dict(G_ema), "config": diffusers_config(G_ema.cfg)}
    torch.save(data, fname)
    return fname


def load_network_snapshot(fname, into=None, compute_dtype=torch.bfloat16, device=None, allow_pickle=False):
    """-> UNet2DConditionModel (or `into`, updated in place).  Reads this module's tensor-only snapshots with
    `weights_only=True`.  The reference's format - `pickle.dump({'ema': module})`, which is also what
    training_loop() writes - is a pickle of a module object and executes code on load: it is read only with
    `allow_pickle=True` (trusted files)."""
    try:
        data = torch.load(fname, map_location="cpu", weights_only=True)
    except (pickle.UnpicklingError, RuntimeError) as e:
        if not allow_pickle:
            raise RuntimeError("%s is not a tensor-only snapshot; pass allow_pickle=True to unpickle a trusted "
                               "{'ema': module} file (%s)" % (fname, type(e).__name__)) from e
        with open(fname, "rb") as f:
            data = pickle.load(f)
    ema = data["ema"]
    sd = ema.state_dict() if hasattr(ema, "state_dict") and not isinstance(ema, dict) else ema
    if into is None:
        cfg = infer_config(sd, data.get("config") if isinstance(data, dict) else None)
        return load_unet(dict(sd), cfg=cfg, compute_dtype=compute_dtype, device=device)
    into.load_state_dict({k: v.to(torch.float32) for k, v in sd.items()}, strict=True)
    return into


# ---- Adam state in torch.optim.Adam.state_dict() layout --------------------------------------------------------
def adam_state_dict(params, exp_avg_sq, offsets, step_count, lr, betas=(0.0, 0.999), eps=1e-8, exp_avg=None):
    """`params`: the network's parameters in bucket order; `exp_avg_sq` (`exp_avg`): flat fp32 buckets laid out like the
    master bucket (params.FlatParams); -> the dict `torch.optim.Adam(params, lr, betas, eps).state_dict()` would hold
    after `step_count` steps (beta1 = 0 keeps no first moment here: exported as zeros, which is what Adam would hold)."""
    state = {}
    if step_count > 0:
        for i, (p, off) in enumerate(zip(params, offsets)):
            n = p.numel()
            v = exp_avg_sq[off:off + n].detach().to("cpu", torch.float32)
            m = exp_avg[off:off + n].detach().to("cpu", torch.float32) if exp_avg is not None else torch.zeros(n)
            # bucket slices follow the parameter's PHYSICAL layout (channels_last conv weights): view them back
            v = torch.as_strided(v, p.shape, p.stride()).contiguous()
            m = torch.as_strided(m, p.shape, p.stride()).contiguous()
            state[i] = {"step": torch.tensor(float(step_count)), "exp_avg": m, "exp_avg_sq": v}
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "params": list(range(len(params)))}
    return {"state": state, "param_groups": [group]}


def load_adam_state_dict(sd, params, exp_avg_sq, offsets, exp_avg=None):
    """inverse of adam_state_dict: fills the flat bucket(s) from a torch Adam state dict; -> step count."""
    step = 0
    for i, (p, off) in enumerate(zip(params, offsets)):
        st = sd["state"].get(i, sd["state"].get(str(i)))
        if st is None:
            continue
        step = max(step, int(float(st["step"])))
        for src_key, bucket in (("exp_avg_sq", exp_avg_sq), ("exp_avg", exp_avg)):
            if bucket is None:
                continue
            view = torch.as_strided(bucket, p.shape, p.stride(), off)
            if tuple(st[src_key].shape) != tuple(p.shape):
                raise ValueError("optimizer state %d (%s): shape %s does not match parameter shape %s - the state dict "
                                 "was written for a different parameter order" % (i, src_key, tuple(st[src_key].shape),
                                                                               tuple(p.shape)))
            view.copy_(st[src_key].to(view.device, torch.float32))
    return step


def _flat_adam_state(net, lr, betas, eps):
    fl = net.flat
    if fl is None or fl.exp_avg_sq is None:
        return adam_state_dict(list(net.parameters()), torch.zeros(0), [0] * len(list(net.parameters())), 0, lr, betas, eps)
    return adam_state_dict(fl.params, fl.exp_avg_sq, fl.offsets, fl.step_count, lr, betas, eps, exp_avg=fl.exp_avg)


# ---- training state (resume) -----------------------------------------------------------------------------------
def save_training_state(step, fname):
    """`step`: training.step.SiDLSGStep.  Same five keys as sid_training_loop.py:654-656 (+ the image counter)."""
    data = {
        "format": SNAPSHOT_FORMAT,
        "fake_score": plain_state_dict(step.fake_score), "G": plain_state_dict(step.G),
        "G_ema": plain_state_dict(step.G_ema if step.G_ema is not None else step.G),
        "fake_score_optimizer_state": _flat_adam_state(step.fake_score, step.lr, step.betas, step.eps),
        "g_optimizer_state": _flat_adam_state(step.G, step.glr, step.betas, step.eps),
        "cur_nimg": int(step.cur_nimg),
    }
    torch.save(data, fname)
    return fname


def load_training_state(step, fname, allow_pickle=False):
    """restores networks, Adam buckets and the image counter (sid_training_loop.py:296-310)."""
    try:
        data = torch.load(fname, map_location="cpu", weights_only=True)     # this module's own dumps: tensors + containers
    except (pickle.UnpicklingError, RuntimeError):
        if not allow_pickle:
            raise
        data = torch.load(fname, map_location="cpu", weights_only=False)    # the reference's dumps hold module objects

    def sd_of(x):
        return x.state_dict() if hasattr(x, "state_dict") and not isinstance(x, dict) else x

    for key, net in (("fake_score", step.fake_score), ("G", step.G), ("G_ema", step.G_ema)):
        if net is None:
            continue
        net.load_state_dict({k: v.to(torch.float32) for k, v in sd_of(data[key]).items()}, strict=True)
    for key, net in (("fake_score_optimizer_state", step.fake_score), ("g_optimizer_state", step.G)):
        fl = net.flat
        if fl is None:
            continue
        if fl.exp_avg_sq is None:
            fl.init_adam(step.betas[0])
        fl.step_count = load_adam_state_dict(data[key], fl.params, fl.exp_avg_sq, fl.offsets, exp_avg=fl.exp_avg)
    step.cur_nimg = int(data.get("cur_nimg", step.cur_nimg))
    return step
