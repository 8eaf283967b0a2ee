"""The oracle (oracle/) against fixtures produced by the REFERENCE's own code
(tests/golden/make_golden.py ran /root/reference/training/sid_sd_util.py and
/root/reference/training/sid_training_loop.py unmodified on CPU)."""
import copy

import pytest
import torch

from oracle import DDPMSchedule, UNet2DCondition, TINY, SD15, SD21_BASE
from oracle import step as ostep
from golden_util import load, parse_loop, embedding_table


def fresh_unet():
    torch.manual_seed(0)
    return UNet2DCondition(TINY)


def test_param_counts_and_keys():
    for cfg, n in ((SD15, 859_520_964), (SD21_BASE, 865_910_724)):
        with torch.device("meta"):
            m = UNet2DCondition(cfg)
        assert sum(p.numel() for p in m.parameters()) == n
        assert len(list(m.parameters())) == 686
    keys = set(UNet2DCondition(TINY).state_dict().keys())
    for k in ("time_embedding.linear_1.weight", "down_blocks.0.resnets.0.time_emb_proj.bias",
              "down_blocks.2.attentions.1.transformer_blocks.0.attn2.to_out.0.bias",
              "down_blocks.1.downsamplers.0.conv.weight", "mid_block.attentions.0.proj_in.weight",
              "up_blocks.0.upsamplers.0.conv.bias", "up_blocks.3.attentions.2.transformer_blocks.0.ff.net.2.weight",
              "up_blocks.2.resnets.0.conv_shortcut.weight", "conv_norm_out.weight", "conv_out.bias"):
        assert k in keys, k
    assert "down_blocks.3.attentions.0.norm.weight" not in keys
    assert "up_blocks.3.upsamplers.0.conv.weight" not in keys


def test_alpha_table_spot_values():
    s = DDPMSchedule()
    for t, v in ((20, 0.98131430), (156, 0.81929100), (312, 0.57029963), (468, 0.32071662), (625, 0.13776892),
                 (979, 0.00591277)):
        assert abs(float(s.alphas_cumprod[t]) - v) < 2e-7, (t, float(s.alphas_cumprod[t]))
    assert abs(float(s.alphas_cumprod[625]) ** 0.5 - 0.37117237) < 1e-6
    assert abs(float(1 - s.alphas_cumprod[625]) ** 0.5 - 0.92856401) < 1e-6


def test_glue_matches_reference():
    fx = load("glue.pt")
    unet = fresh_unet().eval().requires_grad_(False)
    chk = torch.tensor([sum(float(p.double().sum()) for p in unet.parameters()),
                        sum(float(p.double().abs().sum()) for p in unet.parameters())], dtype=torch.float64)
    assert torch.allclose(chk, fx["unet_checksum"], rtol=1e-9), "seeded init drifted: regenerate fixtures"
    sched = DDPMSchedule()
    table = embedding_table(TINY.cross_attention_dim)
    cond = table[fx["ctx_ids"]]
    uncond = table[torch.zeros(3, dtype=torch.long)]
    init_t = 625 * torch.ones((3,), dtype=torch.long)
    with torch.no_grad():
        x1 = ostep.sampler(unet, sched, fx["z"], cond, init_t)
        assert torch.allclose(x1, fx["sampler_1step"], rtol=1e-5, atol=1e-6)
        # analytic identity at num_steps=1
        eps = unet(0.92856401 * fx["z"], init_t, encoder_hidden_states=cond).sample
        assert torch.allclose(x1, (0.92856401 * fx["z"] - 0.92856401 * eps) / 0.37117237, rtol=1e-4, atol=1e-5)
        for kappa in (1, 1.5, 4.5):
            for px0, nm in ((True, "x0"), (False, "eps")):
                y = ostep.denoise(unet, sched, fx["sampler_1step"], fx["noise"], cond, uncond, fx["t"],
                                  predict_x0=px0, guidance_scale=kappa)
                assert torch.allclose(y, fx[f"denoise_{nm}_k{kappa}"], rtol=1e-5, atol=1e-5), (kappa, nm)
        x4 = ostep.sampler(unet, sched, fx["z"], cond, init_t, num_steps=4, sub_noise=list(fx["sampler_4step_sub_noise"]))
        assert torch.allclose(x4, fx["sampler_4step"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["loop_1step.pt", "loop_2step_alpha12.pt"])
def test_iteration_matches_reference_training_loop(name):
    fx = load(name)
    iters = parse_loop(fx, TINY.cross_attention_dim)
    true_score = fresh_unet().eval().requires_grad_(False)
    fake = copy.deepcopy(true_score).train().requires_grad_(True)
    G = copy.deepcopy(true_score).train().requires_grad_(True)
    G_ema = copy.deepcopy(G).eval().requires_grad_(False)
    opt_f = ostep.make_optimizer(fake.parameters(), lr=fx["lr"])
    opt_g = ostep.make_optimizer(G.parameters(), lr=fx["lr"])
    sched = DDPMSchedule()
    ref_losses = [(n, v) for n, v in fx["losses"] if n.endswith("Loss/loss")]
    ref_f = [v for n, v in ref_losses if n.startswith("fake")][::2]
    ref_g = [v for n, v in ref_losses if n.startswith("G_")][::2]
    cur_nimg = 0
    for it, (mb_f, mb_g) in enumerate(iters):
        lf, lg = ostep.iteration(G, G_ema, fake, true_score, sched, opt_f, opt_g, mb_f, mb_g, kappa=fx["kappa"],
                                 alpha=fx["alpha"], batch_size=fx["batch"], cur_nimg=cur_nimg,
                                 ema_halflife_kimg=fx["ema_halflife_kimg"], num_steps=fx["num_steps"],
                                 loss_scaling=fx["loss_scaling"], loss_scaling_G=fx["loss_scaling_G"])
        cur_nimg += fx["batch"]
        assert abs(lf - ref_f[it]) <= 2e-4 * abs(ref_f[it]), (it, lf, ref_f[it])
        assert abs(lg - ref_g[it]) <= 2e-3 * abs(ref_g[it]) + 1e-3, (it, lg, ref_g[it])
    for key, net in (("G", G), ("fake_score", fake), ("G_ema", G_ema)):
        sd = net.state_dict()
        init = true_score.state_dict()
        for k, ref in fx[key].items():
            # compare the *update* (lr 1e-3 Adam steps), not just the init
            assert torch.allclose(sd[k] - init[k], ref - init[k], atol=2e-4), (key, k, (sd[k] - ref).abs().max())
        moved = max(float((fx[key][k] - init[k]).abs().max()) for k in fx[key])
        assert moved > 5e-4, "fixture should contain a visible update"
