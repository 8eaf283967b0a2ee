"""UNet2DConditionModel restated from diffusers==0.27.2 behaviour;
the reference calls it at /root/reference/training/sid_sd_util.py:184,194,245,263.

TEST INFRASTRUCTURE (oracle): plain torch.nn, fp32, NCHW, no fused anything.  Module / parameter
names reproduce the diffusers state-dict keys (686 tensors, 859,520,964 params for SD1.5).
"""
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .scheduler import timestep_embedding


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    # diffusers' `attention_head_dim` is really the number of heads per block
    num_heads: Tuple[int, ...] = (8, 8, 8, 8)
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    sample_size: int = 64

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4


SD15 = UNetConfig()
SD21_BASE = UNetConfig(cross_attention_dim=1024, num_heads=(5, 10, 20, 20), use_linear_projection=True)
# CPU-seconds config used by the parity fixtures (same topology, small widths)
TINY = UNetConfig(block_out_channels=(32, 64, 128, 128), cross_attention_dim=64, num_heads=(2, 2, 4, 4),
                  norm_num_groups=8, sample_size=16)
TINY_LINEAR = UNetConfig(block_out_channels=(32, 64, 128, 128), cross_attention_dim=48, num_heads=(1, 2, 4, 4),
                         norm_num_groups=8, use_linear_projection=True, sample_size=16)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, heads, context_dim=None):
        super().__init__()
        self.heads = heads
        kv = dim if context_dim is None else context_dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(kv, dim, bias=False)
        self.to_v = nn.Linear(kv, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        b, n, c = x.shape
        h = self.heads
        q = self.to_q(x).view(b, n, h, c // h).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], h, c // h).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], h, c // h).transpose(1, 2)
        s = torch.matmul(q, k.transpose(-1, -2)) * (c // h) ** -0.5
        p = torch.softmax(s, dim=-1)
        o = torch.matmul(p, v).transpose(1, 2).reshape(b, n, c)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        u, g = self.proj(x).chunk(2, dim=-1)
        return u * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, context_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, context_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), context)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, context_dim, groups, linear_proj):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        if linear_proj:
            self.proj_in = nn.Linear(dim, dim)
            self.proj_out = nn.Linear(dim, dim)
        else:
            self.proj_in = nn.Conv2d(dim, dim, 1)
            self.proj_out = nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, context_dim)])

    def forward(self, x, context):
        b, c, hh, ww = x.shape
        r = x
        x = self.norm(x)
        if self.linear_proj:
            x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
            x = self.proj_in(x)
        else:
            x = self.proj_in(x)
            x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        for blk in self.transformer_blocks:
            x = blk(x, context)
        if self.linear_proj:
            x = self.proj_out(x)
            x = x.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        else:
            x = x.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
            x = self.proj_out(x)
        return x + r


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg, cin, cout, heads, cross, add_down):
        super().__init__()
        t, g, e = cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, t, g, e)
                                      for j in range(cfg.layers_per_block)])
        if cross:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cfg.cross_attention_dim, g,
                                                                cfg.use_linear_projection)
                                             for _ in range(cfg.layers_per_block)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, context):
        outs = []
        for j, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[j](x, context)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg, c, heads):
        super().__init__()
        t, g, e = cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, t, g, e), ResnetBlock2D(c, c, t, g, e)])
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, cfg.cross_attention_dim, g,
                                                            cfg.use_linear_projection)])

    def forward(self, x, temb, context):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, context)
        return self.resnets[1](x, temb)


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
        self.resnets = nn.ModuleList(res)
        if cross:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cfg.cross_attention_dim, g,
                                                                cfg.use_linear_projection) for _ in range(n)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, context):
        for j, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[j](x, context)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class _TimeEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class UNet2DCondition(nn.Module):
    def __init__(self, cfg: UNetConfig = SD15):
        super().__init__()
        self.cfg = cfg
        self.config = SimpleNamespace(in_channels=cfg.in_channels, sample_size=cfg.sample_size,
                                      cross_attention_dim=cfg.cross_attention_dim)
        ch = cfg.block_out_channels
        nb = len(ch)
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = _TimeEmbedding(ch[0], cfg.time_embed_dim)
        downs = []
        cout = ch[0]
        for i in range(nb):
            cin, cout = cout, ch[i]
            downs.append(DownBlock(cfg, cin, cout, cfg.num_heads[i], cross=(i < nb - 1), add_down=(i < nb - 1)))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(cfg, ch[-1], cfg.num_heads[-1])
        rev = tuple(reversed(ch))
        rheads = tuple(reversed(cfg.num_heads))
        ups = []
        cout = rev[0]
        for i in range(nb):
            cprev, cout = cout, rev[i]
            cin = rev[min(i + 1, nb - 1)]
            ups.append(UpBlock(cfg, cin, cout, cprev, rheads[i], cross=(i > 0), add_up=(i < nb - 1)))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states=None, return_dict=True):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        if t.dim() == 0:
            t = t[None]
        t = t.expand(sample.shape[0])
        temb = timestep_embedding(t, self.cfg.block_out_channels[0]).to(sample.dtype)
        temb = self.time_embedding(temb)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return SimpleNamespace(sample=x)
