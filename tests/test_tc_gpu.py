"""tcgen05 tensor-core paths (gemm_tc.cu) through the C ABI vs fp32 PyTorch on bf16-rounded inputs.

Inputs are rounded to bf16 first, so the only differences are fp32 accumulation order and the bf16 rounding of
the outputs: tolerance 1.5e-2 relative to the output scale for bf16 outputs, 2e-3 for fp32 (wgrad) outputs.
Shapes are chosen to satisfy the tensor-core eligibility rules (channels % 64 == 0, >= 128 pixels ...).
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


def bf(t):
    return t.bfloat16().float()


def check(a, b, tol, what):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item() / scale
    rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
    assert err < tol * 4 and rel < tol, "%s: max err/scale %g, rel L2 %g" % (what, err, rel)


class BfParam:
    """fp32 master Parameter with a bf16 shadow and a zeroed fp32 grad, like params.FlatParams produces."""

    def __new__(cls, t, channels_last=False):
        t = bf(t).to(DEV)
        if channels_last:
            t = t.contiguous(memory_format=torch.channels_last)
        p = torch.nn.Parameter(t)
        p.grad = torch.zeros_like(p)
        p._shadow = p.detach().to(torch.bfloat16)  # preserves the physical layout
        return p


@pytest.mark.parametrize("M,K,N", [(256, 64, 64), (4096, 320, 320), (1000, 768, 640), (512, 1280, 2560), (300, 640, 136),
                                   (8192, 320, 960), (24576, 320, 320), (640, 320, 96), (384, 128, 80), (2048, 1280, 1280),
                                   (16384, 640, 320), (9000, 320, 2560), (8192, 384, 136)])
def test_tc_linear(M, K, N):
    g = torch.Generator().manual_seed(0)
    x = bf(torch.randn(M, K, generator=g))
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g)
    r = bf(torch.randn(M, N, generator=g))
    dy = bf(torch.randn(M, N, generator=g))
    xr, wr, br, rr = (t.clone().requires_grad_(True) for t in (x, w, b, r))
    yr = F.linear(xr, wr, br) + rr
    yr.backward(dy)
    xd = x.to(DEV).bfloat16().requires_grad_(True)
    rd = r.to(DEV).bfloat16().requires_grad_(True)
    wp = BfParam(w)
    bp = torch.nn.Parameter(b.to(DEV))
    bp.grad = torch.zeros_like(bp)
    y = ops().linear(xd, wp, bp, rd)
    y.backward(dy.to(DEV).bfloat16())
    check(y, yr, 1.5e-2, "y")
    check(xd.grad, xr.grad, 1.5e-2, "dx")
    check(wp.grad, wr.grad, 2e-3, "dw")
    check(bp.grad, br.grad, 2e-3, "db")


@pytest.mark.parametrize("B,H,C,N,stride", [(2, 16, 64, 64, 1), (1, 64, 64, 128, 1), (3, 32, 128, 64, 1),
                                            (4, 8, 320, 320, 1), (2, 16, 192, 320, 1), (1, 64, 320, 320, 1),
                                            (5, 8, 64, 192, 1), (2, 32, 64, 128, 2), (4, 16, 128, 128, 2),
                                            (1, 64, 320, 320, 2), (3, 16, 64, 64, 2), (8, 64, 64, 128, 1),
                                            (16, 8, 128, 320, 1)])
def test_tc_conv3x3(B, H, C, N, stride):
    g = torch.Generator().manual_seed(1)
    x = bf(torch.randn(B, C, H, H, generator=g))
    w = bf(torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C))
    b = torch.randn(N, generator=g)
    rv = torch.randn(B, N, generator=g)
    res = bf(torch.randn(B, N, H // stride, H // stride, generator=g))
    xr, wr, br, rvr = (t.clone().requires_grad_(True) for t in (x, w, b, rv))
    yr = F.conv2d(xr, wr, br, padding=1, stride=stride) + rvr[:, :, None, None] + res
    dy = bf(torch.randn(yr.shape, generator=g))
    yr.backward(dy)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16().requires_grad_(True)
    rvd = rv.to(DEV).requires_grad_(True)
    wp = BfParam(w, channels_last=True)
    bp = torch.nn.Parameter(b.to(DEV))
    bp.grad = torch.zeros_like(bp)
    resd = res.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
    y = ops().conv3x3(xd, wp, bp, resd, rvd, stride)
    y.backward(dy.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16())
    check(y.permute(0, 3, 1, 2), yr, 1.5e-2, "y")
    check(xd.grad.permute(0, 3, 1, 2), xr.grad, 1.5e-2, "dx")
    check(wp.grad, wr.grad, 2e-3, "dw")
    check(rvd.grad, rvr.grad, 1e-2, "drowvec")


def test_tc_wgrad_accumulates():
    """two backward passes (gradient accumulation rounds) add up in the fp32 grad bucket."""
    g = torch.Generator().manual_seed(2)
    x = bf(torch.randn(2, 64, 16, 16, generator=g))
    w = bf(torch.randn(64, 64, 3, 3, generator=g) / 24)
    wr = w.clone().requires_grad_(True)
    wp = BfParam(w, channels_last=True)
    for _ in range(2):
        dy = bf(torch.randn(2, 64, 16, 16, generator=g))
        F.conv2d(x, wr, None, padding=1).backward(dy)
        xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
        ops().conv3x3(xd, wp).backward(dy.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16())
    check(wp.grad, wr.grad, 2e-3, "dw accumulated")


@pytest.mark.parametrize("B,N,M,C,heads", [(2, 4096, 4096, 320, 8), (2, 1024, 1024, 640, 8), (3, 256, 256, 1280, 8),
                                           (2, 64, 64, 1280, 8), (2, 4096, 77, 320, 8), (1, 1024, 1024, 640, 10),
                                           (2, 256, 77, 1280, 20), (1, 200, 333, 128, 2), (2, 256, 77, 1280, 8),
                                           (3, 64, 77, 1280, 8)])
def test_tc_flash_attention(B, N, M, C, heads):
    """tcgen05 flash attention vs fp32 softmax(QK^T/sqrt(d))V on bf16-rounded inputs (bf16 P and bf16 output:
    tolerance 2e-2 of the output scale); backward = tcgen05 flash backward for d <= 80, recompute path for d = 160."""
    g = torch.Generator().manual_seed(4)
    q = bf(torch.randn(B, N, C, generator=g))
    k = bf(torch.randn(B, M, C, generator=g))
    v = bf(torch.randn(B, M, C, generator=g))
    d = C // heads

    def split(t):
        return t.view(t.shape[0], t.shape[1], heads, d).transpose(1, 2)

    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    s = split(qr) @ split(kr).transpose(-1, -2) * d ** -0.5
    o_ref = (torch.softmax(s, -1) @ split(vr)).transpose(1, 2).reshape(B, N, C)
    qd, kd, vd = (t.to(DEV).bfloat16().requires_grad_(True) for t in (q, k, v))
    o = ops().attention(qd, kd, vd, heads)
    check(o, o_ref, 2e-2, "o")
    if N * M <= 1024 * 1024 or d <= 80:
        do = bf(torch.randn(B, N, C, generator=g))
        o_ref.backward(do)
        o.backward(do.to(DEV).bfloat16())
        check(qd.grad, qr.grad, 3e-2, "dq")
        check(kd.grad, kr.grad, 3e-2, "dk")
        check(vd.grad, vr.grad, 3e-2, "dv")


@pytest.mark.parametrize("nb1,nb2,M,N,K,a_mn,b_mn,f32out", [(3, 8, 256, 256, 160, 0, 0, 1), (2, 8, 256, 77, 160, 0, 0, 1),
                                                          (2, 4, 77, 160, 256, 1, 1, 0), (2, 4, 256, 160, 77, 0, 1, 0),
                                                          (1, 5, 64, 160, 64, 1, 1, 0), (2, 3, 300, 96, 200, 0, 0, 0)])
def test_tc_batched_gemm(nb1, nb2, M, N, K, a_mn, b_mn, f32out):
    """nb1 x nb2 independent GEMMs in ONE tcgen05 launch (4-D operand maps): the score / value contractions of the
    d = 160 attention backward, incl. 77 text keys on rows padded to 80 elements.  Must not fall back to CUDA cores."""
    from sid_lsg_b200._lib import lib, F32
    import ctypes
    o = ops()
    g = torch.Generator().manual_seed(7)
    Kp, Mp, Np = (K + 7) // 8 * 8, (M + 7) // 8 * 8, (N + 7) // 8 * 8
    # physical layouts with padded leading dimensions; pad columns hold NaN to prove they are never read
    if a_mn:
        A = torch.full((nb1, nb2, K, Mp), float("nan")); A[..., :M] = bf(torch.randn(nb1, nb2, K, M, generator=g))
        a_log = A[..., :M].transpose(-1, -2)
        a_sm, a_sk, a_sb = 1, Mp, (nb2 * K * Mp, K * Mp)
    else:
        A = torch.full((nb1, nb2, M, Kp), float("nan")); A[..., :K] = bf(torch.randn(nb1, nb2, M, K, generator=g))
        a_log = A[..., :K]
        a_sm, a_sk, a_sb = Kp, 1, (nb2 * M * Kp, M * Kp)
    if b_mn:
        Bm = torch.full((nb1, nb2, K, Np), float("nan")); Bm[..., :N] = bf(torch.randn(nb1, nb2, K, N, generator=g))
        b_log = Bm[..., :N]
        b_sn, b_sk, b_sb = 1, Np, (nb2 * K * Np, K * Np)
    else:
        Bm = torch.full((nb1, nb2, N, Kp), float("nan")); Bm[..., :K] = bf(torch.randn(nb1, nb2, N, K, generator=g))
        b_log = Bm[..., :K].transpose(-1, -2)
        b_sn, b_sk, b_sb = Kp, 1, (nb2 * N * Kp, N * Kp)
    ref = a_log.float() @ b_log.float()
    ldc = Np
    C = torch.zeros((nb1, nb2, M, ldc), dtype=torch.float32 if f32out else torch.bfloat16, device=DEV)
    cnt = (ctypes.c_long * 2)()
    lib.query("counters", cnt)
    simt0 = cnt[1]
    o.gemm(A.to(DEV).bfloat16(), a_sm, a_sk, Bm.to(DEV).bfloat16(), b_sn, b_sk, C, ldc, M, N, K, a_sb=a_sb, b_sb=b_sb,
           c_sb=(nb2 * M * ldc, M * ldc), nb=(nb1, nb2), out_dtype=F32 if f32out else None)
    torch.cuda.synchronize()
    lib.query("counters", cnt)
    assert cnt[1] == simt0, "batched GEMM fell back to the CUDA-core kernel"
    check(C[..., :N], ref, 2e-3 if f32out else 1.5e-2, "C")


@pytest.mark.parametrize("B,N,M,C,heads,cross", [(2, 1024, 1024, 320, 8, False), (2, 256, 256, 1280, 8, False),
                                                 (2, 1024, 77, 320, 8, True), (2, 256, 77, 1280, 8, True),
                                                 (1, 64, 64, 1280, 8, False)])
def test_packed_attention_matches_separate(B, N, M, C, heads, cross):
    """attention on the q|k|v thirds of ONE packed projection output (row-strided operands, gradients written into
    one packed tensor) == attention on separate dense tensors: same kernels, same arithmetic -> bit-identical."""
    o_ = ops()
    g = torch.Generator().manual_seed(11)
    q = torch.randn(B, N, C, generator=g).to(DEV).bfloat16()
    k = torch.randn(B, M, C, generator=g).to(DEV).bfloat16()
    v = torch.randn(B, M, C, generator=g).to(DEV).bfloat16()
    do = torch.randn(B, N, C, generator=g).to(DEV).bfloat16()
    qs, ks, vs = (t.clone().requires_grad_(True) for t in (q, k, v))
    o_ref = o_.attention(qs, ks, vs, heads)
    o_ref.backward(do)
    if cross:
        a = q.clone().requires_grad_(True)
        b = torch.cat([k, v], dim=-1).requires_grad_(True)
        o = o_.packed_attention(a, b, heads)
        o.backward(do)
        dq, dk, dv = a.grad, b.grad[..., :C], b.grad[..., C:]
    else:
        a = torch.cat([q, k, v], dim=-1).requires_grad_(True)
        o = o_.packed_attention(a, None, heads)
        o.backward(do)
        dq, dk, dv = a.grad[..., :C], a.grad[..., C:2 * C], a.grad[..., 2 * C:]
    assert torch.equal(o, o_ref)
    # dQ partials meet through fp32 atomics / bulk reductions whose order is not fixed: compare with a tolerance
    check(dq, qs.grad, 2e-3, "dq")
    check(dk, ks.grad, 2e-3, "dk")
    check(dv, vs.grad, 2e-3, "dv")


@pytest.mark.parametrize("rows,C", [(77, 32), (300, 320), (130, 640), (64, 1280), (5000, 320)])
def test_bf16_layernorm(rows, C):
    """vectorised LayerNorm kernels (16-byte pieces per lane) in the bf16 activation mode."""
    g = torch.Generator().manual_seed(5)
    x = bf(torch.randn(rows, C, generator=g) * 1.5 - 0.3)
    ga = torch.randn(C, generator=g)
    be = torch.randn(C, generator=g)
    dy = bf(torch.randn(rows, C, generator=g))
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, ga, be))
    yr = F.layer_norm(xr, (C,), gr, br, 1e-5)
    yr.backward(dy)
    xd = x.to(DEV).bfloat16().requires_grad_(True)
    gp = torch.nn.Parameter(ga.to(DEV)); gp.grad = torch.zeros_like(gp)
    bp = torch.nn.Parameter(be.to(DEV)); bp.grad = torch.zeros_like(bp)
    y = ops().layer_norm(xd, gp, bp, 1e-5)
    y.backward(dy.to(DEV).bfloat16())
    check(y, yr, 1.5e-2, "y")
    check(xd.grad, xr.grad, 1.5e-2, "dx")
    check(gp.grad, gr.grad, 2e-3, "dgamma")
    check(bp.grad, br.grad, 2e-3, "dbeta")


def test_tma_store_switch_matches():
    """staged TMA-store epilogue == per-row store epilogue (bit-exact: same arithmetic, different way out)."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, sys; sys.path.insert(0, %r)\n"
        "from sid_lsg_b200 import ops\n"
        "g = torch.Generator().manual_seed(0)\n"
        "x = torch.randn(3000, 320, generator=g).cuda().bfloat16()\n"
        "w = (torch.randn(960, 320, generator=g) / 18).cuda()\n"
        "p = torch.nn.Parameter(w); p._shadow = w.bfloat16(); p.grad = torch.zeros_like(p)\n"
        "b = torch.randn(960, generator=g).cuda()\n"
        "r = torch.randn(3000, 960, generator=g).cuda().bfloat16()\n"
        "y = ops.linear(x, p, b, r)\n"
        "torch.cuda.synchronize(); print(float(y.float().double().sum()), float(y.float().abs().double().sum()))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("1", "0"):
        env = dict(os.environ, SIDLSG_TMA_STORE=flag)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1], outs


def test_bm2_tiles_match():
    """256-row tiles (two A tiles per B tile, SIDLSG_BM2=2 forces them) == 128-row tiles, bit for bit: the reduction
    order of every output element is the same, only the tile shape differs."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, sys, math; sys.path.insert(0, %r)\n"
        "from sid_lsg_b200 import ops\n"
        "g = torch.Generator().manual_seed(0)\n"
        "def P(t, cl=False):\n"
        "    t = t.cuda()\n"
        "    if cl: t = t.contiguous(memory_format=torch.channels_last)\n"
        "    p = torch.nn.Parameter(t); p._shadow = p.detach().bfloat16(); p.grad = torch.zeros_like(p); return p\n"
        "out = []\n"
        "for (M, K, N) in ((1000, 1280, 320), (4096, 2560, 640), (512, 320, 960), (2304, 640, 136)):\n"
        "    x = torch.randn(M, K, generator=g).cuda().bfloat16().requires_grad_(True)\n"
        "    w = P(torch.randn(N, K, generator=g) / math.sqrt(K))\n"
        "    b = torch.randn(N, generator=g).cuda()\n"
        "    r = torch.randn(M, N, generator=g).cuda().bfloat16()\n"
        "    y = ops.linear(x, w, b, r)\n"
        "    y.backward(torch.randn(M, N, generator=g).cuda().bfloat16())\n"
        "    out += [y.float().double().sum().item(), y.float().abs().double().sum().item(), x.grad.float().abs().double().sum().item()]\n"
        "for (B, H, C, N, st) in ((4, 16, 128, 320, 1), (2, 32, 64, 128, 2), (8, 8, 320, 192, 1)):\n"
        "    x = torch.randn(B, H, H, C, generator=g).cuda().bfloat16().requires_grad_(True)\n"
        "    w = P(torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C), True)\n"
        "    b = torch.randn(N, generator=g).cuda()\n"
        "    rv = torch.randn(B, N, generator=g).cuda()\n"
        "    y = ops.conv3x3(x, w, b, None, rv, st)\n"
        "    y.backward(torch.randn(y.shape, generator=g).cuda().bfloat16())\n"
        "    out += [y.float().double().sum().item(), y.float().abs().double().sum().item(), x.grad.float().abs().double().sum().item()]\n"
        "torch.cuda.synchronize(); print(' '.join(repr(v) for v in out))\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("0", "2"):
        env = dict(os.environ, SIDLSG_BM2=flag)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1], outs


# ---- fp32-accurate tensor-core mode (csrc/split3.cu): three bf16 passes of the tcgen05 kernel vs fp64 torch --------
def _relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


class F32Param:
    def __new__(cls, t, channels_last=False):
        t = t.to(DEV)
        if channels_last:
            t = t.contiguous(memory_format=torch.channels_last)
        p = torch.nn.Parameter(t)
        p.grad = torch.zeros_like(p)
        return p


@pytest.mark.parametrize("M,K,N", [(4096, 320, 320), (1000, 768, 640), (512, 1280, 2560), (8192, 320, 960)])
def test_split3_linear_matches_fp64(M, K, N):
    """fwd, dgrad, wgrad of a Linear in fp32 through gemm_split3: relative error ~1e-5 (bf16 alone: ~4e-3), and the
    tensor-core kernel really ran (sidlsg_counters)."""
    import ctypes
    from sid_lsg_b200._lib import lib
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    dy = torch.randn(M, N, generator=g)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = F.linear(xr, wr, b.double()) + r.double()
    yr.backward(dy.double())
    cnt0 = (ctypes.c_long * 2)()
    lib.query("counters", cnt0)
    xd = x.to(DEV).requires_grad_(True)
    wd = F32Param(w)
    with ops().tc_split(True):
        yd = ops().linear(xd, wd, b.to(DEV), r.to(DEV))
        yd.backward(dy.to(DEV))
    torch.cuda.synchronize()
    cnt1 = (ctypes.c_long * 2)()
    lib.query("counters", cnt1)
    assert cnt1[0] - cnt0[0] == 9 and cnt1[1] == cnt0[1], (list(cnt0), list(cnt1))   # 3 GEMMs x 3 passes, no CUDA-core GEMM
    assert _relerr(yd, yr) < 3e-5, _relerr(yd, yr)
    assert _relerr(xd.grad, xr.grad) < 3e-5, _relerr(xd.grad, xr.grad)
    assert _relerr(wd.grad, wr.grad) < 3e-5, _relerr(wd.grad, wr.grad)


@pytest.mark.parametrize("B,H,C,N,stride", [(2, 32, 64, 128, 1), (1, 64, 320, 320, 1), (2, 16, 128, 128, 2)])
def test_split3_conv3x3_matches_fp64(B, H, C, N, stride):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)
    b = torch.randn(N, generator=g)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = F.conv2d(xr, wr, b.double(), stride=stride, padding=1)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy.double())
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).requires_grad_(True)
    wd = F32Param(w, channels_last=True)
    with ops().tc_split(True):
        yd = ops().conv3x3(xd, wd, b.to(DEV), stride=stride)
        yd.backward(dy.permute(0, 2, 3, 1).contiguous().to(DEV))
    assert ops().lib.query("last_path") in (0, 1)
    assert _relerr(yd.permute(0, 3, 1, 2), yr) < 3e-5, _relerr(yd.permute(0, 3, 1, 2), yr)
    assert _relerr(xd.grad.permute(0, 3, 1, 2), xr.grad) < 3e-5, _relerr(xd.grad.permute(0, 3, 1, 2), xr.grad)
    assert _relerr(wd.grad, wr.grad) < 3e-5, _relerr(wd.grad, wr.grad)


@pytest.mark.parametrize("B,H,C,N,with_rv", [(2, 64, 4, 320, True), (3, 16, 4, 64, False), (2, 64, 320, 4, False),
                                             (1, 32, 64, 4, False)])
def test_narrow_conv_as_tensor_core_gemm(B, H, C, N, with_rv):
    """conv_in (4 -> 320) / conv_out (320 -> 4) in bf16: im2col of the 4-channel tensor + tcgen05 GEMMs (ops.py
    `_narrow_kind`), forward, data, weight, bias and timestep-row gradients vs fp32 torch; no CUDA-core GEMM runs."""
    import ctypes
    from sid_lsg_b200._lib import lib
    g = torch.Generator().manual_seed(4)
    x = bf(torch.randn(B, C, H, H, generator=g))
    w = bf(torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C))
    b = torch.randn(N, generator=g)
    rv = torch.randn(B, N, generator=g)
    xr, wr, br, rvr = (t.clone().requires_grad_(True) for t in (x, w, b, rv))
    yr = F.conv2d(xr, wr, br, padding=1)
    if with_rv:
        yr = yr + rvr[:, :, None, None]
    dy = bf(torch.randn(yr.shape, generator=g))
    yr.backward(dy)
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16().requires_grad_(True)
    rvd = rv.to(DEV).requires_grad_(True)
    wp = BfParam(w, channels_last=True)
    bp = torch.nn.Parameter(b.to(DEV))
    bp.grad = torch.zeros_like(bp)
    c0 = (ctypes.c_long * 2)()
    lib.query("counters", c0)
    y = ops().conv3x3(xd, wp, bp, None, rvd if with_rv else None)
    y.backward(dy.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16())
    torch.cuda.synchronize()
    c1 = (ctypes.c_long * 2)()
    lib.query("counters", c1)
    assert c1[0] - c0[0] == 3 and c1[1] == c0[1], (list(c0), list(c1))      # fwd, dgrad, wgrad: three tcgen05 GEMMs
    check(y.permute(0, 3, 1, 2), yr, 1.5e-2, "y")
    check(xd.grad.permute(0, 3, 1, 2), xr.grad, 1.5e-2, "dx")
    check(wp.grad, wr.grad, 3e-3, "dw")
    check(bp.grad, br.grad, 3e-3, "dbias")
    if with_rv:
        check(rvd.grad, rvr.grad, 1e-2, "drowvec")
