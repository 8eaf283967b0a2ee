"""N ranks on shards == 1 rank on the concatenated batch, on the PRODUCT path (SiDLSGStep through ddp.FlatDDP over
NCCL): SURVEY.md §4 tier (v), the reference's DDP semantics (/root/reference/training/sid_training_loop.py:316-323,
`loss / batch_gpu_total` at :445, :530).  fp32-exact mode, TINY UNet; needs >= 2 GPUs (gpurun --gpus 2), skipped on a
single-GPU box."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(S, oracle, dev, overlap=True):
    torch.manual_seed(0)
    o = oracle.UNet2DCondition(oracle.TINY)
    nets = []
    for _ in range(4):
        m = S.UNet2DConditionModel(S.TINY, compute_dtype=torch.float32)
        m.load_state_dict(o.state_dict())
        nets.append(m.to(dev).flatten_())
    return S.SiDLSGStep(nets[0], nets[1], nets[2], nets[3], S.DDPMScheduler(device=dev), lr=1e-4, glr=1e-4,
                        cfg_train_fake=1.5, cfg_eval_fake=1.5, cfg_eval_real=1.5, overlap_allreduce=overlap)


def _micro(S, rank, round_idx, phase, dev):
    from sid_lsg_b200.training.step import synth_microbatch
    return synth_microbatch(2, S.TINY, 9000 + 100 * phase + 10 * round_idx + rank, dev, dropout=(phase == 0))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import oracle
    import sid_lsg_b200 as S
    from sid_lsg_b200.torch_utils import distributed as dist
    dist.init()                                               # NCCL
    dev = torch.device("cuda", rank)
    res = {}
    for overlap in (True, False):
        st = _build(S, oracle, dev, overlap)
        rounds = 2
        # 2 accumulation rounds per rank: no_sync on the first, overlapped reduction on the second
        lf = st.fake_score_phase([_micro(S, rank, r, 0, dev) for r in range(rounds)], batch_gpu_total=2 * rounds)
        lg = st.generator_phase([_micro(S, rank, r, 1, dev) for r in range(rounds)], batch_gpu_total=2 * rounds,
                                batch_size=2 * rounds * world)
        torch.cuda.synchronize()
        res[overlap] = dict(f=st.fake_score.flat.master.cpu(), g=st.G.flat.master.cpu(), e=st.G_ema.flat.master.cpu(),
                            reduced=(st.fake_score_ddp.reduced_elems, st.G_ddp.reduced_elems, st.G.flat.numel))
    if rank == 0:
        torch.save(res, os.path.join(out_dir, "r0.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_equal_one_rank_on_concatenated_batch(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = 29100 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "r0.pt"))
    import oracle
    import sid_lsg_b200 as S
    dev = torch.device("cuda", 0)
    st = _build(S, oracle, dev)
    init = st.G.flat.master.cpu().clone()
    rounds = 2
    mf = [_micro(S, rk, r, 0, dev) for r in range(rounds) for rk in range(world)]
    mg = [_micro(S, rk, r, 1, dev) for r in range(rounds) for rk in range(world)]
    st.fake_score_phase(mf, batch_gpu_total=2 * rounds * world)
    st.generator_phase(mg, batch_gpu_total=2 * rounds * world, batch_size=2 * rounds * world)
    torch.cuda.synchronize()
    want = dict(f=st.fake_score.flat.master.cpu(), g=st.G.flat.master.cpu(), e=st.G_ema.flat.master.cpu())
    for overlap in (True, False):
        r = got[overlap]
        assert r["reduced"][0] == r["reduced"][2] and r["reduced"][1] == r["reduced"][2], r["reduced"]
        for k in ("f", "g", "e"):
            # compare the UPDATE (Adam's first step is ~lr * sign(g)); summation order differs between 2+2 and 4 samples
            du, dr = r[k] - init, want[k] - init
            rel = float((du - dr).norm() / dr.norm().clamp_min(1e-30))
            assert rel < 2e-2, (overlap, k, rel)
            assert float((du - dr).abs().max()) < 2.1e-4, (overlap, k)
    # overlapped and non-overlapped reductions add the same numbers; the runs differ only through the order-dependent
    # fp32 atomics of the split-K weight-gradient kernels
    for k in ("f", "g", "e"):
        du, dr = got[True][k] - init, got[False][k] - init
        assert float((du - dr).norm() / dr.norm().clamp_min(1e-30)) < 2e-2, k
