"""CPU-side checks of the drop-in boundary: the header parses, the shared library loads and exports every
declared symbol (no compute calls here), and the product refuses to run without CUDA."""
import ctypes
import os

import pytest
import torch


def test_header_declares_expected_entry_points():
    from sid_lsg_b200._lib import parse_header
    decl = parse_header()
    for name in ("sidlsg_gemm", "sidlsg_conv3x3", "sidlsg_conv3x3_wgrad", "sidlsg_groupnorm_fwd", "sidlsg_groupnorm_bwd",
                 "sidlsg_layernorm_fwd", "sidlsg_layernorm_bwd", "sidlsg_softmax_fwd", "sidlsg_softmax_bwd",
                 "sidlsg_add_noise", "sidlsg_cfg_x0_fwd", "sidlsg_cfg_x0_bwd", "sidlsg_fake_loss", "sidlsg_lsg_loss",
                 "sidlsg_adam_step", "sidlsg_ema_update", "sidlsg_last_error", "sidlsg_version"):
        assert name in decl, name
    assert len(decl["sidlsg_gemm"][1]) == 31


def test_library_exports_every_declared_symbol():
    from sid_lsg_b200._lib import parse_header, LIB_PATH
    if not os.path.exists(LIB_PATH):
        from sid_lsg_b200.build import build
        build()
    dll = ctypes.CDLL(LIB_PATH)
    for name in parse_header():
        assert hasattr(dll, name), "libsidlsg.so does not export %s" % name
    dll.sidlsg_version.restype = ctypes.c_int
    assert dll.sidlsg_version() >= 100
    dll.sidlsg_last_error.restype = ctypes.c_char_p
    assert isinstance(dll.sidlsg_last_error(), bytes)


def test_argument_errors_are_reported_without_a_gpu():
    """Argument validation happens before any launch: status < 0 and a message, RuntimeError in Python."""
    from sid_lsg_b200._lib import lib
    with pytest.raises(RuntimeError, match="adam_step"):
        lib.call("adam_step", None, None, None, None, None, None, None, 6, 1e-3, 0.0, 0.999, 1e-8, 1, 1.0, 0.0, 0.0, 0.0, None, None)
    with pytest.raises(RuntimeError, match="CHW"):
        lib.call("lsg_loss", None, None, None, None, None, None, None, 1, 7, 1.0, 1.0, None)


def test_product_refuses_cpu_tensors():
    from sid_lsg_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(torch.randn(2, 8), torch.nn.Parameter(torch.randn(4, 8)))
    import sid_lsg_b200 as S
    m = S.UNet2DConditionModel(S.TINY)
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 4, 16, 16), torch.tensor([1]), encoder_hidden_states=torch.randn(1, 77, 64))


def test_unet_structure_matches_oracle_and_diffusers_counts():
    import oracle
    import sid_lsg_b200 as S
    for ocfg, cfg, n in ((oracle.SD15, S.SD15, 859_520_964), (oracle.SD21_BASE, S.SD21_BASE, 865_910_724)):
        with torch.device("meta"):
            m = S.UNet2DConditionModel(cfg)
            o = oracle.UNet2DCondition(ocfg)
        assert sum(p.numel() for p in m.parameters()) == n
        so, sm = o.state_dict(), m.state_dict()
        assert sorted(so.keys()) == sorted(sm.keys())   # same names; the product ENUMERATES in diffusers' order (test_surface_cpu)
        assert all(so[k].shape == sm[k].shape for k in so)


def test_ema_beta_matches_oracle():
    from oracle.step import ema_beta as ref
    from sid_lsg_b200 import ema_beta
    for bs, nimg, hl in ((512, 0, 50), (512, 512, 50), (256, 10_000_000, 50), (32, 4096, 0.5)):
        assert ema_beta(bs, nimg, hl) == ref(bs, nimg, hl)
