"""UNet2DConditionModel (SD1.5 / SD2.1-base) on the sm_100a kernels.

Drop-in for the object the reference calls as `unet(sample, timestep, encoder_hidden_states=E).sample`
(/root/reference/training/sid_sd_util.py:184,194,245,263): same call protocol, parameter names equal to the
diffusers 0.27.2 state-dict keys (SURVEY.md App. A-5, 686 tensors) so SD checkpoints and SiD-LSG snapshots load
by name, NCHW fp32 at the boundary.  Inside, activations are token-major [B, H*W, C] in the compute dtype and
every op is a kernel of libsidlsg.so (ops.py); there is no torch.nn.functional arithmetic here.
"""
import copy
import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Tuple

import torch
import torch.nn as nn

from . import ops
from .params import FlatParams


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    num_heads: Tuple[int, ...] = (8, 8, 8, 8)  # diffusers' `attention_head_dim` is really the head count
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    sample_size: int = 64

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4


SD15 = UNetConfig()
SD21_BASE = UNetConfig(cross_attention_dim=1024, num_heads=(5, 10, 20, 20), use_linear_projection=True)
TINY = UNetConfig(block_out_channels=(32, 64, 128, 128), cross_attention_dim=64, num_heads=(2, 2, 4, 4),
                  norm_num_groups=8, sample_size=16)
TINY_LINEAR = UNetConfig(block_out_channels=(32, 64, 128, 128), cross_attention_dim=48, num_heads=(1, 2, 4, 4),
                         norm_num_groups=8, use_linear_projection=True, sample_size=16)


# ---- parameter holders (names/shapes = diffusers) --------------------------------------------------------------
class _Affine(nn.Module):
    """GroupNorm / LayerNorm parameters."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class _Linear(nn.Module):
    def __init__(self, cin, cout, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        _uniform_init(self.weight, self.bias, cin)

    def forward(self, x, res=None):
        return ops.linear(x, self.weight, self.bias, res)


class _Conv(nn.Module):
    """Conv2d parameters [O, I, k, k]; 3x3 weights live channels_last (physically [O,3,3,I])."""

    def __init__(self, cin, cout, k):
        super().__init__()
        w = torch.empty(cout, cin, k, k)
        if k == 3:
            w = w.contiguous(memory_format=torch.channels_last)
        self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(torch.empty(cout))
        _uniform_init(self.weight, self.bias, cin * k * k)


def _uniform_init(w, b, fan_in):
    """torch's default Conv2d / Linear init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias."""
    bound = 1.0 / math.sqrt(fan_in)
    with torch.no_grad():
        w.uniform_(-bound, bound)
        if b is not None:
            b.uniform_(-bound, bound)


# ---- blocks ---------------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups, eps):
        super().__init__()
        self.groups, self.eps = groups, eps
        self.norm1 = _Affine(cin)
        self.conv1 = _Conv(cin, cout, 3)
        self.time_emb_proj = _Linear(temb_dim, cout)
        self.norm2 = _Affine(cout)
        self.conv2 = _Conv(cout, cout, 3)
        self.conv_shortcut = _Conv(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb_act, hw):
        """x [B, HW, Cin]; temb_act = silu(temb) fp32 [B, 1280]."""
        B, HW, cin = x.shape
        H, W = hw
        a1 = ops.group_norm(x, self.norm1.weight, self.norm1.bias, self.groups, self.eps, silu=True)
        tproj = self.time_emb_proj(temb_act)  # fp32 [B, Cout], added per pixel in conv1's epilogue
        h = ops.conv3x3(a1.view(B, H, W, cin), self.conv1.weight, self.conv1.bias, rowvec=tproj)
        cout = h.shape[-1]
        a2 = ops.group_norm(h.view(B, HW, cout), self.norm2.weight, self.norm2.bias, self.groups, self.eps, silu=True)
        if self.conv_shortcut is not None:
            sc = ops.linear(x, self.conv_shortcut.weight, self.conv_shortcut.bias)
        else:
            sc = x
        out = ops.conv3x3(a2.view(B, H, W, cout), self.conv2.weight, self.conv2.bias, res=sc.view(B, H, W, cout))
        return out.view(B, HW, cout)


class Attention(nn.Module):
    def __init__(self, dim, heads, context_dim=None):
        super().__init__()
        self.heads = heads
        kv = dim if context_dim is None else context_dim
        self.to_q = _Linear(dim, dim, bias=False)
        self.to_k = _Linear(kv, dim, bias=False)
        self.to_v = _Linear(kv, dim, bias=False)
        self.to_out = nn.ModuleList([_Linear(dim, dim), nn.Identity()])
        self._fused_cache = {}

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_fused_cache"] = {}          # views of the flat buckets are rebuilt lazily, never pickled
        return st

    def _fused(self, names):
        """FusedWeight over consecutive bias-free projections (built lazily per flat bucket; None when the module's
        parameters are not flattened back to back, e.g. before flatten_())."""
        ws = [getattr(self, n).weight for n in names]
        flat = getattr(ws[0], "_flat", None)
        key = (names, id(flat))
        if self._fused_cache.get("key") != key:
            fw = None
            if flat is not None:
                try:
                    fw = ops.FusedWeight(ws)
                except (ValueError, AttributeError):
                    fw = None
            self._fused_cache = {"key": key, "fw": fw}
        return self._fused_cache["fw"]

    def forward(self, x, context, res):
        if context is None:
            fw = self._fused(("to_q", "to_k", "to_v"))
            if fw is not None:
                # one GEMM for q|k|v (x is read once), attention on the packed thirds
                o = ops.packed_attention(ops.linear_fused(x, fw), None, self.heads)
                return self.to_out[0](o, res=res)
        else:
            fw = self._fused(("to_k", "to_v"))
            if fw is not None:
                o = ops.packed_attention(self.to_q(x), ops.linear_fused(context, fw), self.heads)
                return self.to_out[0](o, res=res)
        ctx = x if context is None else context
        q = self.to_q(x)
        k = self.to_k(ctx)
        v = self.to_v(ctx)
        o = ops.attention(q, k, v, self.heads)
        return self.to_out[0](o, res=res)  # residual add fused into the output projection


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = _Linear(dim, inner * 2)

    def forward(self, x):
        return ops.geglu(self.proj(x))


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), _Linear(dim * 4, dim)])

    def forward(self, x, res):
        return self.net[2](self.net[0](x), res=res)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, context_dim):
        super().__init__()
        self.norm1 = _Affine(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = _Affine(dim)
        self.attn2 = Attention(dim, heads, context_dim)
        self.norm3 = _Affine(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = self.attn1(ops.layer_norm(x, self.norm1.weight, self.norm1.bias), None, res=x)
        x = self.attn2(ops.layer_norm(x, self.norm2.weight, self.norm2.bias), context, res=x)
        x = self.ff(ops.layer_norm(x, self.norm3.weight, self.norm3.bias), res=x)
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, context_dim, groups, linear_proj):
        super().__init__()
        self.groups = groups
        self.norm = _Affine(dim)
        # token-major data makes Conv2d-1x1 and Linear projections the same GEMM; only the stored shape differs
        # registration order = diffusers' (norm, proj_in, transformer_blocks, proj_out): parameters() order is what
        # index-keyed optimiser state dicts are matched by (training/checkpoint.py)
        self.proj_in = _Linear(dim, dim) if linear_proj else _Conv(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, context_dim)])
        self.proj_out = _Linear(dim, dim) if linear_proj else _Conv(dim, dim, 1)

    def forward(self, x, context):
        r = x
        x = ops.group_norm(x, self.norm.weight, self.norm.bias, self.groups, 1e-6, silu=False)
        x = ops.linear(x, self.proj_in.weight, self.proj_in.bias)
        for blk in self.transformer_blocks:
            x = blk(x, context)
        return ops.linear(x, self.proj_out.weight, self.proj_out.bias, res=r)


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = _Conv(c, c, 3)

    def forward(self, x, hw):
        B, HW, C = x.shape
        y = ops.conv3x3(x.view(B, hw[0], hw[1], C), self.conv.weight, self.conv.bias, stride=2)
        return y.view(B, -1, C), (y.shape[1], y.shape[2])


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = _Conv(c, c, 3)

    def forward(self, x, hw):
        B, HW, C = x.shape
        if x.dtype == torch.bfloat16 or ops.split_active():
            # tensor-core modes: materialise the 2x copy (HBM-cheap) so the conv is a plain TMA implicit GEMM
            y = ops.conv3x3(ops.upsample2x(x.view(B, hw[0], hw[1], C)), self.conv.weight, self.conv.bias)
        else:
            y = ops.conv3x3(x.view(B, hw[0], hw[1], C), self.conv.weight, self.conv.bias, up=2)
        return y.view(B, -1, C), (y.shape[1], y.shape[2])


class DownBlock(nn.Module):
    def __init__(self, cfg, cin, cout, heads, cross, add_down):
        super().__init__()
        t, g, e = cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps
        # diffusers registers attentions before resnets in the cross-attention blocks
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cfg.cross_attention_dim, g,
                                                            cfg.use_linear_projection)
                                         for _ in range(cfg.layers_per_block)]) if cross else None
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, t, g, e)
                                      for j in range(cfg.layers_per_block)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, hw, temb_act, context):
        outs = []
        for j, res in enumerate(self.resnets):
            x = res(x, temb_act, hw)
            if self.attentions is not None:
                x = self.attentions[j](x, context)
            outs.append(x)
        if self.downsamplers is not None:
            x, hw = self.downsamplers[0](x, hw)
            outs.append(x)
        return x, hw, outs


class MidBlock(nn.Module):
    def __init__(self, cfg, c, heads):
        super().__init__()
        t, g, e = cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, cfg.cross_attention_dim, g,
                                                            cfg.use_linear_projection)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, t, g, e), ResnetBlock2D(c, c, t, g, e)])

    def forward(self, x, hw, temb_act, context):
        x = self.resnets[0](x, temb_act, hw)
        x = self.attentions[0](x, context)
        return self.resnets[1](x, temb_act, hw)


class UpBlock(nn.Module):
    def __init__(self, cfg, cin, cout, cprev, heads, cross, add_up):
        super().__init__()
        t, g, e = cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps
        n = cfg.layers_per_block + 1
        res = []
        for j in range(n):
            skip = cin if j == n - 1 else cout
            rin = cprev if j == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, t, g, e))
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cfg.cross_attention_dim, g,
                                                            cfg.use_linear_projection) for _ in range(n)]) if cross else None
        self.resnets = nn.ModuleList(res)
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, hw, skips, temb_act, context):
        for j, res in enumerate(self.resnets):
            x = ops.concat(x, skips.pop())
            x = res(x, temb_act, hw)
            if self.attentions is not None:
                x = self.attentions[j](x, context)
        if self.upsamplers is not None:
            x, hw = self.upsamplers[0](x, hw)
        return x, hw


class _TimeEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = _Linear(cin, dim)
        self.linear_2 = _Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(ops.silu(self.linear_1(x)))


class UNet2DConditionModel(nn.Module):
    """`compute_dtype` float32 = fp32-exact mode (CUDA-core FFMA GEMMs, the reference's TF32-off fp32 semantics,
    training/sid_training_loop.py:241-243); bfloat16 = tensor-core mode (fp32 master weights + bf16 shadow,
    fp32 accumulation, fp32 norm statistics / scheduler / loss)."""

    def __init__(self, cfg: UNetConfig = SD15, compute_dtype=torch.float32, device=None, tc_split=False):
        super().__init__()
        self.cfg = cfg
        self.compute_dtype = compute_dtype
        # fp32 only: run the contractions on tcgen05 as three bf16 passes (fp32-accurate; csrc/split3.cu) instead of the
        # CUDA-core kernels, so the 1e-3 parity claim is made on the kernel that is benchmarked
        self.tc_split = bool(tc_split) and compute_dtype == torch.float32
        self.config = SimpleNamespace(in_channels=cfg.in_channels, sample_size=cfg.sample_size,
                                      cross_attention_dim=cfg.cross_attention_dim)
        ch = cfg.block_out_channels
        nb = len(ch)
        self.conv_in = _Conv(cfg.in_channels, ch[0], 3)
        self.time_embedding = _TimeEmbedding(ch[0], cfg.time_embed_dim)
        downs = []
        cout = ch[0]
        for i in range(nb):
            cin, cout = cout, ch[i]
            downs.append(DownBlock(cfg, cin, cout, cfg.num_heads[i], cross=(i < nb - 1), add_down=(i < nb - 1)))
        # diffusers creates both ModuleLists before the mid block, so parameters() runs down, up, mid
        self.down_blocks = nn.ModuleList(downs)
        self.up_blocks = nn.ModuleList()
        self.mid_block = MidBlock(cfg, ch[-1], cfg.num_heads[-1])
        rev = tuple(reversed(ch))
        rheads = tuple(reversed(cfg.num_heads))
        ups = []
        cout = rev[0]
        for i in range(nb):
            cprev, cout = cout, rev[i]
            cin = rev[min(i + 1, nb - 1)]
            ups.append(UpBlock(cfg, cin, cout, cprev, rheads[i], cross=(i > 0), add_up=(i < nb - 1)))
        self.up_blocks.extend(ups)
        self.conv_norm_out = _Affine(ch[0])
        self.conv_out = _Conv(ch[0], cfg.out_channels, 3)
        half = ch[0] // 2
        # same expression as diffusers get_timestep_embedding (evaluated on the host, fp32)
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
        self.register_buffer("_freqs", freqs, persistent=False)
        self.flat = None
        self._grad_ready = None   # set by ddp.FlatDDP for one forward: callable(stage) fired from backward
        if device is not None:
            self.to(device)
            self.flatten_()

    # -- parameter storage ----------------------------------------------------------------------------------
    def flatten_(self):
        """Re-home all parameters into flat buckets (params.FlatParams); call once the module is on its GPU."""
        self.flat = FlatParams(self, shadow=(self.compute_dtype == torch.bfloat16))
        return self

    def unflatten_(self):
        """Give every parameter its own storage again (the inverse of flatten_): used before device moves and
        pickling, where views into multi-GB buckets must not travel."""
        if self.flat is None:
            return self
        with torch.no_grad():
            for p in self.flat.params:
                p.data = p.data.clone(memory_format=torch.preserve_format)
                p.grad = None
                p._shadow = None
                p._flat = None
        self.flat = None
        return self

    def _apply(self, fn, recurse=True):
        # .to(device) / .cpu() / .cuda() as the reference does on the networks it deep-copies and pickles
        # (sid_training_loop.py:284-287, 641-650; generate_onestep.py:247-248): leave the buckets first, re-flatten
        # lazily at the next forward
        if getattr(self, "flat", None) is not None:
            self.unflatten_()
        return super()._apply(fn, recurse)

    def __getstate__(self):
        """pickle / torch.save of the MODULE (`pickle.dump({'ema': G_ema})`, `torch.save(dict(G=G, ...))`,
        sid_training_loop.py:641-656): configuration + a plain CPU state dict, never the flat buckets."""
        return {"cfg": self.cfg, "compute_dtype": self.compute_dtype, "tc_split": self.tc_split, "training": self.training,
                "requires_grad": [p.requires_grad for p in self.parameters()],
                "state_dict": {k: v.detach().to("cpu", torch.float32).contiguous().clone()
                               for k, v in self.state_dict().items()}}

    def __setstate__(self, st):
        self.__init__(st["cfg"], st["compute_dtype"], tc_split=st.get("tc_split", False))
        torch.nn.Module.load_state_dict(self, st["state_dict"], strict=True)
        for p, rg in zip(self.parameters(), st["requires_grad"]):
            p.requires_grad_(rg)
        self.train(st["training"])

    def __deepcopy__(self, memo):
        """copy.deepcopy(unet) as the reference does for fake_score / G / G_ema
        (training/sid_training_loop.py:286-287, 327): fresh buckets, same values."""
        dev = next(self.parameters()).device
        new = UNet2DConditionModel(self.cfg, self.compute_dtype, tc_split=self.tc_split)
        new.to(dev)
        with torch.no_grad():
            for p_new, p in zip(new.parameters(), self.parameters()):
                p_new.copy_(p)
                p_new.requires_grad_(p.requires_grad)
        new.train(self.training)
        if self.flat is not None:
            new.flatten_()
        memo[id(self)] = new
        return new

    def load_state_dict(self, state_dict, strict=True, assign=False):
        r = super().load_state_dict(state_dict, strict=strict, assign=False)
        if self.flat is not None:
            self.flat.refresh_shadow()
        return r

    def grad_stages(self):
        """[(start, end)] element ranges of the flat gradient bucket per forward stage (see forward()): stage 0 =
        conv_in + time_embedding, 1..n = down blocks, then mid, up blocks, and the output norm + conv."""
        children = ([[self.conv_in, self.time_embedding]] + [[b] for b in self.down_blocks] + [[self.mid_block]] +
                    [[b] for b in self.up_blocks] + [[self.conv_norm_out, self.conv_out]])
        return [self.flat.ranges_of(mods) for mods in children]

    # diffusers-protocol no-ops kept so reference call sites keep working (sid_sd_util.py:111,116)
    def enable_xformers_memory_efficient_attention(self):
        return None

    def enable_gradient_checkpointing(self):
        return None

    # -- forward ----------------------------------------------------------------------------------------------
    def forward(self, sample, timestep, encoder_hidden_states=None, return_dict=True):
        with ops.tc_split(self.tc_split):
            return self._forward(sample, timestep, encoder_hidden_states, return_dict)

    def _forward(self, sample, timestep, encoder_hidden_states=None, return_dict=True):
        cfg = self.cfg
        if self.flat is None:
            if not sample.is_cuda or next(self.parameters()).device != sample.device:
                raise RuntimeError("sid_lsg_b200 UNet: move the module to the sample's CUDA device first "
                                   "(no CPU path; parameters are flattened into device buckets at the first forward)")
            self.flatten_()
        B, _, H, W = sample.shape
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        if t.dim() == 0:
            t = t[None]
        t = t.to(sample.device).expand(B)
        cd = self.compute_dtype
        temb = ops.timestep_embedding(t, self._freqs, cfg.block_out_channels[0])  # fp32 path throughout
        temb_act = ops.silu(self.time_embedding(temb))  # every ResnetBlock2D consumes silu(temb)
        context = ops.cast(encoder_hidden_states, cd)
        x = ops.nchw_to_tokens(sample, cd)
        hw = (H, W)
        x = ops.conv3x3(x.view(B, H, W, cfg.in_channels), self.conv_in.weight, self.conv_in.bias).view(B, H * W, -1)
        # Gradient-ready notifications for the data-parallel reducer (ddp.FlatDDP).  The autograd engine runs ready
        # nodes in decreasing creation order, so when the gradient of the activation ENTERING stage k is complete every
        # layer of stages k.. has launched its weight-gradient kernels: the flat gradient ranges of those stages are
        # final and their allreduce can start while the earlier stages are still in backward.
        ready, self._grad_ready = self._grad_ready, None
        stage = [0]

        def mark(t):
            stage[0] += 1
            if ready is not None and t.requires_grad:
                t.register_hook(lambda g, k=stage[0]: ready(k))
            return t

        skips = [x]
        for blk in self.down_blocks:
            x, hw, outs = blk(mark(x), hw, temb_act, context)       # stages 1..4
            skips.extend(outs)
        x = self.mid_block(mark(x), hw, temb_act, context)           # stage 5
        for blk in self.up_blocks:
            x, hw = blk(mark(x), hw, skips, temb_act, context)       # stages 6..9
        mark(x)                                                      # stage 10: conv_norm_out + conv_out
        x = ops.group_norm(x, self.conv_norm_out.weight, self.conv_norm_out.bias, cfg.norm_num_groups, cfg.norm_eps,
                           silu=True)
        x = ops.conv3x3(x.view(B, hw[0], hw[1], -1), self.conv_out.weight, self.conv_out.bias)
        out = ops.tokens_to_nchw(x.view(B, hw[0] * hw[1], cfg.out_channels), hw[0], hw[1])
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)
