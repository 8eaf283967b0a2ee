"""World-size-2 (gloo, CPU) checks of the data-parallel host logic: the flat-bucket gradient reduction used by
SiDLSGStep (one all_reduce(SUM) per network, mean folded into the optimiser's grad_scale) reproduces the
single-process gradient of the concatenated batch - the reference's DDP semantics
(/root/reference/training/sid_training_loop.py:316-323, :445 `/ batch_gpu_total`), and bench.py's reference arm
only works on rank 0."""
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oracle import step as ostep
    from sid_lsg_b200.params import FlatParams
    from sid_lsg_b200.training.step import synth_microbatch, _world
    torch.manual_seed(0)
    net = oracle.UNet2DCondition(oracle.TINY)
    sched = oracle.DDPMSchedule()
    B = 4
    shard = B // world
    # every rank draws ITS OWN micro-batch (seed depends on rank), like bench.py
    m = synth_microbatch(shard, oracle.TINY, 1000 + rank, None)
    eps = ostep.denoise(net, sched, m["z"], m["noise"], m["cond"], m["uncond"], m["t"], predict_x0=False, guidance_scale=1.5)
    loss, _ = ostep.fake_score_loss(eps, m["noise"], 1.0, shard)          # / batch_gpu_total, as the reference
    loss.backward()

    class Bucket:  # the part of FlatParams the reduction touches
        pass
    bk = Bucket()
    bk.grad = torch.cat([p.grad.flatten() for p in net.parameters()])
    FlatParams.allreduce_grad(bk)
    assert _world() == world
    mean = bk.grad * (1.0 / _world())                                     # grad_scale of adam_step
    if rank == 0:
        torch.save(dict(mean=mean, z=m["z"]), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucket_reduce_equals_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    sys.path.insert(0, ROOT)
    import oracle
    from oracle import step as ostep
    from sid_lsg_b200.training.step import synth_microbatch
    torch.manual_seed(0)
    net = oracle.UNet2DCondition(oracle.TINY)
    sched = oracle.DDPMSchedule()
    ms = [synth_microbatch(2, oracle.TINY, 1000 + r, None) for r in range(2)]
    assert not torch.equal(ms[0]["z"], ms[1]["z"]) and torch.equal(ms[0]["z"], got["z"])
    cat = {k: torch.cat([m[k] for m in ms]) for k in ("z", "noise", "cond", "uncond", "t")}
    eps = ostep.denoise(net, sched, cat["z"], cat["noise"], cat["cond"], cat["uncond"], cat["t"], predict_x0=False,
                        guidance_scale=1.5)
    loss, _ = ostep.fake_score_loss(eps, cat["noise"], 1.0, 4)
    loss.backward()
    ref = torch.cat([p.grad.flatten() for p in net.parameters()])
    rel = ((got["mean"] - ref).norm() / ref.norm()).item()   # fp32 summation order differs between 2+2 and 4 samples
    assert rel < 1e-5, rel


def test_bench_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_reference_arm_prints_contract_line():
    """rank 0 of the reference arm on a tiny model (keeps the CPU suite short); checks the JSON contract keys."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "TINY",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0
