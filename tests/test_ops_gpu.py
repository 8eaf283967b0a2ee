"""Kernel-level parity (through the C ABI) against plain fp32 PyTorch on the CPU.

fp32 mode tolerance: rtol 2e-4 / atol 2e-5 (same arithmetic, different summation order);
bf16 mode: inputs/outputs rounded to bf16 (8-bit mantissa): rtol 3e-2 / atol 3e-2 on O(1) data.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def ops():
    from sid_lsg_b200 import ops as o
    return o


def close(a, b, rtol=2e-4, atol=2e-5, what=""):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, rtol=rtol, atol=atol), "%s max err %g (ref max %g)" % (what, err, b.abs().max().item())


def P(t, channels_last=False):
    t = t.to(DEV)
    if channels_last:
        t = t.contiguous(memory_format=torch.channels_last)
    p = torch.nn.Parameter(t)
    p.grad = torch.zeros_like(p)
    return p


@pytest.mark.parametrize("M,K,N", [(77, 48, 64), (512, 320, 640), (3, 1280, 320), (1024, 64, 136), (32, 1280, 1280),
                                   (16, 1280, 640), (100, 1280, 320), (32, 320, 1280)])
def test_linear_fwd_bwd(M, K, N):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    dy = torch.randn(M, N, generator=g)
    xr, wr, br, rr = (t.clone().requires_grad_(True) for t in (x, w, b, r))
    yr = F.linear(xr, wr, br) + rr
    yr.backward(dy)
    xd = x.to(DEV).requires_grad_(True)
    rd = r.to(DEV).requires_grad_(True)
    wp, bp = P(w), P(b)
    y = ops().linear(xd, wp, bp, rd)
    y.backward(dy.to(DEV))
    close(y, yr, what="y")
    close(xd.grad, xr.grad, what="dx")
    close(rd.grad, rr.grad, what="dres")
    close(wp.grad, wr.grad, rtol=1e-3, atol=1e-4, what="dw")
    close(bp.grad, br.grad, rtol=1e-3, atol=1e-4, what="db")


@pytest.mark.parametrize("B,H,C,N,stride,up", [(2, 16, 32, 64, 1, 1), (1, 8, 96, 32, 1, 1), (2, 16, 32, 32, 2, 1),
                                                (2, 8, 64, 64, 1, 2), (3, 16, 4, 32, 1, 1), (2, 16, 32, 4, 1, 1)])
def test_conv3x3_fwd_bwd(B, H, C, N, stride, up):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)
    b = torch.randn(N, generator=g)
    rv = torch.randn(B, N, generator=g)
    xr, wr, br, rvr = (t.clone().requires_grad_(True) for t in (x, w, b, rv))
    xin = F.interpolate(xr, scale_factor=2.0, mode="nearest") if up == 2 else xr
    yr = F.conv2d(xin, wr, br, stride=stride, padding=1) + rvr[:, :, None, None]
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).requires_grad_(True)
    rvd = rv.to(DEV).requires_grad_(True)
    wp, bp = P(w, channels_last=True), P(b)
    y = ops().conv3x3(xd, wp, bp, None, rvd, stride, up)
    y.backward(dy.permute(0, 2, 3, 1).contiguous().to(DEV))
    close(y.permute(0, 3, 1, 2), yr, what="y")
    close(xd.grad.permute(0, 3, 1, 2), xr.grad, what="dx")
    close(wp.grad, wr.grad, rtol=1e-3, atol=1e-4, what="dw")
    close(bp.grad, br.grad, rtol=1e-3, atol=1e-4, what="db")
    close(rvd.grad, rvr.grad, rtol=1e-3, atol=1e-4, what="drowvec")


@pytest.mark.parametrize("silu", [False, True])
@pytest.mark.parametrize("B,HW,C,G", [(2, 256, 32, 8), (3, 64, 320, 32), (1, 16, 1920, 32)])
def test_groupnorm(B, HW, C, G, silu):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, HW, C, generator=g) * 2 + 0.5
    ga = torch.randn(C, generator=g)
    be = torch.randn(C, generator=g)
    dy = torch.randn(B, HW, C, generator=g)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, ga, be))
    yr = F.group_norm(xr.transpose(1, 2), G, gr, br, 1e-5)
    if silu:
        yr = F.silu(yr)
    yr = yr.transpose(1, 2)
    yr.backward(dy)
    xd = x.to(DEV).requires_grad_(True)
    gp, bp = P(ga), P(be)
    y = ops().group_norm(xd, gp, bp, G, 1e-5, silu)
    y.backward(dy.to(DEV))
    close(y, yr, what="y")
    close(xd.grad, xr.grad, rtol=1e-3, atol=1e-4, what="dx")
    close(gp.grad, gr.grad, rtol=1e-3, atol=1e-3, what="dgamma")
    close(bp.grad, br.grad, rtol=1e-3, atol=1e-3, what="dbeta")


@pytest.mark.parametrize("rows,C", [(77, 32), (300, 320), (64, 1280), (10, 640)])
def test_layernorm(rows, C):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(rows, C, generator=g) * 1.5 - 0.3
    ga = torch.randn(C, generator=g)
    be = torch.randn(C, generator=g)
    dy = torch.randn(rows, C, generator=g)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, ga, be))
    yr = F.layer_norm(xr, (C,), gr, br, 1e-5)
    yr.backward(dy)
    xd = x.to(DEV).requires_grad_(True)
    gp, bp = P(ga), P(be)
    y = ops().layer_norm(xd, gp, bp, 1e-5)
    y.backward(dy.to(DEV))
    close(y, yr, what="y")
    close(xd.grad, xr.grad, rtol=1e-3, atol=1e-4, what="dx")
    close(gp.grad, gr.grad, rtol=1e-3, atol=1e-3, what="dgamma")
    close(bp.grad, br.grad, rtol=1e-3, atol=1e-3, what="dbeta")


@pytest.mark.parametrize("B,N,M,C,heads", [(2, 64, 64, 64, 2), (1, 256, 77, 128, 4), (3, 16, 77, 40, 1), (1, 1024, 1024, 80, 2)])
def test_attention(B, N, M, C, heads):
    g = torch.Generator().manual_seed(4)
    q = torch.randn(B, N, C, generator=g)
    k = torch.randn(B, M, C, generator=g)
    v = torch.randn(B, M, C, generator=g)
    do = torch.randn(B, N, C, generator=g)
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    d = C // heads

    def split(t):
        return t.view(t.shape[0], t.shape[1], heads, d).transpose(1, 2)

    s = split(qr) @ split(kr).transpose(-1, -2) * d ** -0.5
    o_ref = (torch.softmax(s, -1) @ split(vr)).transpose(1, 2).reshape(B, N, C)
    o_ref.backward(do)
    qd, kd, vd = (t.to(DEV).requires_grad_(True) for t in (q, k, v))
    o = ops().attention(qd, kd, vd, heads)
    o.backward(do.to(DEV))
    close(o, o_ref, rtol=1e-3, atol=1e-4, what="o")
    close(qd.grad, qr.grad, rtol=1e-3, atol=1e-4, what="dq")
    close(kd.grad, kr.grad, rtol=1e-3, atol=1e-4, what="dk")
    close(vd.grad, vr.grad, rtol=1e-3, atol=1e-4, what="dv")


def test_geglu_silu_concat_layout():
    g = torch.Generator().manual_seed(5)
    h = torch.randn(50, 256, generator=g)
    dy = torch.randn(50, 128, generator=g)
    hr = h.clone().requires_grad_(True)
    u, gate = hr.chunk(2, dim=-1)
    yr = u * F.gelu(gate)
    yr.backward(dy)
    hd = h.to(DEV).requires_grad_(True)
    y = ops().geglu(hd)
    y.backward(dy.to(DEV))
    close(y, yr, what="geglu")
    close(hd.grad, hr.grad, what="geglu grad")
    x = torch.randn(7, 64, generator=g)
    xr = x.clone().requires_grad_(True)
    F.silu(xr).backward(torch.ones_like(x))
    xd = x.to(DEV).requires_grad_(True)
    y = ops().silu(xd)
    y.backward(torch.ones_like(y))
    close(y, F.silu(x), what="silu")
    close(xd.grad, xr.grad, what="silu grad")
    a = torch.randn(2, 9, 32, generator=g)
    b = torch.randn(2, 9, 64, generator=g)
    ad, bd = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    c = ops().concat(ad, bd)
    close(c, torch.cat([a, b], -1), rtol=0, atol=0, what="concat")
    w = torch.randn(2, 9, 96, generator=g)
    c.backward(w.to(DEV))
    close(ad.grad, w[..., :32], rtol=0, atol=0)
    close(bd.grad, w[..., 32:], rtol=0, atol=0)
    s = torch.randn(3, 4, 8, 8, generator=g)
    sd = s.to(DEV).requires_grad_(True)
    tok = ops().nchw_to_tokens(sd, torch.float32)
    close(tok, s.permute(0, 2, 3, 1).reshape(3, 64, 4), rtol=0, atol=0)
    back = ops().tokens_to_nchw(tok, 8, 8)
    close(back, s, rtol=0, atol=0)
    back.backward(s.to(DEV))
    close(sd.grad, s, rtol=0, atol=0)


def test_timestep_embedding_matches_oracle():
    from oracle.scheduler import timestep_embedding
    t = torch.tensor([0, 1, 20, 156, 625, 979, 999])
    half = 160
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    out = ops().timestep_embedding(t.to(DEV), freqs.to(DEV), 320)
    close(out, timestep_embedding(t, 320), rtol=0, atol=2e-6, what="temb")


def test_scheduler_matches_oracle():
    from oracle import DDPMSchedule
    from oracle import step as ostep
    from sid_lsg_b200 import DDPMScheduler
    g = torch.Generator().manual_seed(6)
    so, sc = DDPMSchedule(), DDPMScheduler()
    assert torch.equal(so.alphas_cumprod, sc.alphas_cumprod)
    x0 = torch.randn(5, 4, 16, 16, generator=g)
    n = torch.randn(5, 4, 16, 16, generator=g)
    t = torch.tensor([20, 156, 625, 979, 300])
    close(sc.add_noise(x0.to(DEV), n.to(DEV), t.to(DEV)), so.add_noise(x0, n, t), rtol=1e-6, atol=1e-6, what="add_noise")
    close(sc.add_noise(None, n.to(DEV), t.to(DEV)), so.add_noise(torch.zeros_like(x0), n, t), rtol=1e-6, atol=1e-6)
    eu = torch.randn(5, 4, 16, 16, generator=g)
    ec = torch.randn(5, 4, 16, 16, generator=g)
    for kappa in (1.5, 4.5):
        eps = eu + kappa * (ec - eu)
        x0_ref = torch.stack([so.step(e, tt, xx).pred_original_sample for e, tt, xx in zip(eps, t, x0)])
        xd = x0.to(DEV).requires_grad_(True)
        eud, ecd = eu.to(DEV).requires_grad_(True), ec.to(DEV).requires_grad_(True)
        y = sc.pred_x0(eud, ecd, xd, t.to(DEV), kappa, True)
        close(y, x0_ref, rtol=1e-5, atol=1e-5, what="x0")
        w = torch.randn(y.shape, generator=g)
        y.backward(w.to(DEV))
        xr, eur, ecr = (v.clone().requires_grad_(True) for v in (x0, eu, ec))
        sa = so.alphas_cumprod[t].sqrt()[:, None, None, None]
        sb = (1 - so.alphas_cumprod[t]).sqrt()[:, None, None, None]
        ((xr - sb * (eur + kappa * (ecr - eur))) / sa).backward(w)
        close(xd.grad, xr.grad, rtol=1e-5, atol=1e-5)
        close(eud.grad, eur.grad, rtol=1e-5, atol=1e-5)
        close(ecd.grad, ecr.grad, rtol=1e-5, atol=1e-5)
    # scalar-timestep step() as the sampler calls it
    y = sc.step(eu.to(DEV), torch.tensor(625), x0.to(DEV)).pred_original_sample
    close(y, so.step(eu, 625, x0).pred_original_sample, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("alpha", [1.0, 1.2])
@pytest.mark.parametrize("nan_rows", [(), (1,), (0, 1, 2, 3)])
def test_losses_match_oracle(alpha, nan_rows):
    from oracle import step as ostep
    g = torch.Generator().manual_seed(7)
    B, shape = 4, (4, 4, 16, 16)
    xg, yr, yf, nz = (torch.randn(shape, generator=g) for _ in range(4))
    for r in nan_rows:
        yf[r, 0, 0, r] = float("nan")
    total = 8
    xr_, yrr, yfr = (t.clone().requires_grad_(True) for t in (xg, yr, yf))
    lo, n_valid = ostep.generator_loss(xr_, yrr, yfr, alpha, 100.0, total)
    if n_valid > 0:
        lo.backward()
    xd, yrd, yfd = (t.to(DEV).requires_grad_(True) for t in (xg, yr, yf))
    l, out = ops().lsg_loss(xd, yrd, yfd, alpha, 100.0 / total)
    l.backward()
    assert int(out[1].item()) == n_valid
    if n_valid > 0:
        close(l, lo, rtol=1e-4, atol=1e-4, what="lsg loss")
        for a, b in ((xd, xr_), (yrd, yrr), (yfd, yfr)):
            ref = torch.nan_to_num(b.grad, nan=0.0)  # rows dropped by the reference get exactly zero here
            for r in nan_rows:
                ref[r] = 0
            close(a.grad, ref, rtol=1e-4, atol=1e-5, what="lsg grad")
    else:
        assert float(l.item()) == 0.0 and float(xd.grad.abs().max()) == 0.0
    # fake-score loss
    er = yf.clone().requires_grad_(True)
    lo, n_valid = ostep.fake_score_loss(er, nz, 1.0, total)
    if n_valid > 0:
        lo.backward()
    ed = yf.to(DEV).requires_grad_(True)
    l, out = ops().fake_loss(ed, nz.to(DEV), 1.0 / total)
    l.backward()
    assert int(out[1].item()) == n_valid
    if n_valid > 0:
        close(l, lo, rtol=1e-4, atol=1e-4, what="fake loss")
        ref = torch.nan_to_num(er.grad, nan=0.0)
        for r in nan_rows:
            ref[r] = 0
        close(ed.grad, ref, rtol=1e-4, atol=1e-5, what="fake grad")


def test_adam_ema_matches_torch():
    from sid_lsg_b200 import FlatParams
    g = torch.Generator().manual_seed(8)
    lin = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5))
    ref = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5))
    ref.load_state_dict(lin.state_dict())
    ema_ref = [p.detach().clone() for p in ref.parameters()]
    lin.to(DEV)
    ema = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5)).to(DEV)
    ema.load_state_dict(lin.state_dict())
    fp, fe = FlatParams(lin, shadow=True), FlatParams(ema, shadow=True)
    fp.ensure_grad()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-2, betas=(0.0, 0.999), eps=1e-8)
    for it in range(3):
        grads = [torch.randn(p.shape, generator=g) for p in ref.parameters()]
        grads[0][0, 0] = float("nan")
        grads[0][0, 1] = float("inf")
        grads[1][0] = float("-inf")
        for p, pd, gr in zip(ref.parameters(), lin.parameters(), grads):
            p.grad = torch.nan_to_num(gr.clone(), nan=0, posinf=1e5, neginf=-1e5).clamp(-1, 1)
            pd.grad.copy_(gr * 2)  # world-size 2 sum; grad_scale 0.5 restores the mean
        opt.step()
        fp.adam_step(1e-2, (0.0, 0.999), 1e-8, grad_scale=0.5, clip=1.0, ema=fe, ema_beta=0.9)
        for e, p in zip(ema_ref, ref.parameters()):
            e.copy_(p.detach().lerp(e, 0.9))
    for p, pd, e, ed in zip(ref.parameters(), lin.parameters(), ema_ref, ema.parameters()):
        close(pd, p, rtol=1e-5, atol=1e-6, what="adam param")
        close(ed, e, rtol=1e-5, atol=1e-6, what="ema")
        close(pd._shadow, p.detach().bfloat16(), rtol=1e-2, atol=1e-2, what="bf16 shadow")
        # the EMA network's bf16 shadow follows its master (a bf16 forward of G_ema must see the averaged weights)
        assert torch.equal(ed._shadow, ed.detach().bfloat16()), "ema bf16 shadow is stale"


def test_no_cpu_fallback():
    x = torch.randn(4, 8)
    w = torch.nn.Parameter(torch.randn(8, 8))
    with pytest.raises(RuntimeError):
        ops().linear(x, w)


@pytest.mark.parametrize("silu", [False, True])
@pytest.mark.parametrize("B,HW,C,G", [(2, 4096, 320, 32), (2, 1024, 640, 32), (1, 4096, 960, 32), (2, 256, 1280, 32),
                                      (3, 64, 2560, 32), (2, 1024, 1920, 32), (2, 256, 32, 8), (5, 100, 64, 8)])
def test_groupnorm_bf16_sd_shapes(B, HW, C, G, silu):
    """bf16 GroupNorm (+SiLU) forward / backward at the SD shapes vs fp32 torch on the bf16-rounded input."""
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(B, HW, C, generator=g) * 2 + 0.5).bfloat16().float()
    ga = torch.randn(C, generator=g)
    be = torch.randn(C, generator=g)
    dy = torch.randn(B, HW, C, generator=g).bfloat16().float()
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, ga, be))
    yr = F.group_norm(xr.transpose(1, 2), G, gr, br, 1e-5)
    if silu:
        yr = F.silu(yr)
    yr = yr.transpose(1, 2)
    yr.backward(dy)
    xd = x.to(DEV).bfloat16().requires_grad_(True)
    gp, bp = P(ga), P(be)
    y = ops().group_norm(xd, gp, bp, G, 1e-5, silu)
    y.backward(dy.to(DEV).bfloat16())

    def relerr(a, b):
        a, b = a.detach().float().cpu(), b.detach().float().cpu()
        return float((a - b).norm() / b.norm().clamp_min(1e-20))
    assert relerr(y, yr) < 6e-3, relerr(y, yr)                  # bf16 output rounding
    assert relerr(xd.grad, xr.grad) < 6e-3, relerr(xd.grad, xr.grad)
    assert relerr(gp.grad, gr.grad) < 2e-3, relerr(gp.grad, gr.grad)   # fp32 accumulations of bf16 inputs
    assert relerr(bp.grad, br.grad) < 2e-3, relerr(bp.grad, br.grad)
