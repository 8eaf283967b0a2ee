#!/usr/bin/env python
"""bench.py - SiD-LSG train images/sec on B200 (BASELINE.json metric).

A step = one full SiD-LSG iteration (fake-score update + generator update + EMA,
/root/reference/training/sid_training_loop.py:383-567) on `--batch` images per GPU; images/s = global batch /
step time.  Workload at N=1 = BASELINE.json configs[1]: SD1.5 (random init), kappa 1.5, batch 32, 64x64x4 latents,
bf16 compute (fp32 master weights / Adam / losses), synthetic prompt embeddings.  N>1: weak scaling, data
parallel, one NCCL allreduce of each network's flat gradient bucket per iteration.

  python bench.py [--gpus N] [--steps K] [--warmup W]          (torchrun launches the N>1 ranks)
  python bench.py --impl reference ...                          the CPU oracle restatement on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what every field means.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

F_UNET = 0.8033e12          # SD1.5 UNet forward FLOPs per sample
STEP_FLOPS_PER_IMAGE = 18 * F_UNET   # 8 forward + 10 backward-equivalent sample passes


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per iteration")
    ap.add_argument("--batch-gpu", type=int, default=32, help="micro-batch (gradient accumulation rounds = batch / batch-gpu)")
    ap.add_argument("--kappa", type=float, default=1.5)
    ap.add_argument("--num-steps", type=int, default=1, help="generator sub-steps (config 5 uses 4)")
    ap.add_argument("--model", default="SD15", choices=["SD15", "SD21_BASE", "TINY"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (ncu launch-list captures only)")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0)
    ap.add_argument("--shapes", default=None, help="write the per-shape kernel table of the roofline step to this file")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the iteration as a CUDA graph in the timed regions "
                                                         "(training.step.GraphedIteration), 0: eager launches")
    return ap.parse_args()


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:  # noqa: BLE001
            return None
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (NVML, every 200 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = get(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_oracle_iteration_time(args, budget_s, steps, warmup, parity=None):
    """The reference's algorithm (oracle restatement: pure torch.nn fp32, TF32 off, per-sample x0 loop, torch Adam)
    on this host's cores: one SiD-LSG iteration at batch 1 per step.  Returns (images/s, ms/step, steps run, cores).
    `parity` (dict, optional) receives the first (warm-up) iteration's inputs, initial weights and results, which the
    caller replays on the GPU path to MEASURE the benchmarked mode's deviation from the fp32 reference arithmetic."""
    import oracle
    from oracle import step as ostep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = getattr(oracle.unet, args.model)
    torch.manual_seed(0)
    true_score = oracle.UNet2DCondition(cfg).eval().requires_grad_(False)
    fake = oracle.UNet2DCondition(cfg)
    G = oracle.UNet2DCondition(cfg)
    with torch.no_grad():
        for m in (fake, G):
            for p, q in zip(m.parameters(), true_score.parameters()):
                p.copy_(q)
    sched = oracle.DDPMSchedule()
    opt_f, opt_g = ostep.make_optimizer(fake.parameters()), ostep.make_optimizer(G.parameters())
    times = []
    t_begin = time.perf_counter()
    i = 0
    done_warm = 0
    while True:
        mb_f = [ostep.synth_microbatch(1, cfg, 1000 + i, dropout=True, num_steps=args.num_steps)]
        mb_g = [ostep.synth_microbatch(1, cfg, 2000 + i, num_steps=args.num_steps)]
        if parity is not None and i == 0:
            parity["init"] = {k: v.clone() for k, v in true_score.state_dict().items()}
            parity["mb_f"], parity["mb_g"] = mb_f, mb_g
            with torch.no_grad():
                parity["latents"] = ostep.sampler(true_score, sched, mb_g[0]["z"], mb_g[0]["cond"], torch.full((1,), 625),
                                                  num_steps=args.num_steps, sub_noise=mb_g[0].get("sub_noise"))
        t0 = time.perf_counter()
        lf, lg = ostep.iteration(G, None, fake, true_score, sched, opt_f, opt_g, mb_f, mb_g, kappa=args.kappa,
                                 batch_size=1, num_steps=args.num_steps)
        dt = time.perf_counter() - t0
        if parity is not None and i == 0:
            parity["loss_fake"], parity["loss_G"] = float(lf), float(lg)
        i += 1
        if done_warm < min(warmup, 1):
            done_warm += 1
        else:
            times.append(dt)
        elapsed = time.perf_counter() - t_begin
        if len(times) >= steps or (times and elapsed + dt > budget_s):
            break
    ms = 1e3 * sum(times) / len(times)
    return 1e3 / ms, ms, len(times), cores


def run_reference(args, rank):
    if rank != 0:
        return
    v, ms, n, cores = cpu_oracle_iteration_time(args, args.cpu_budget_s, args.steps, args.warmup)
    sample = ("each step = 1 full SiD-LSG iteration (f_psi + G_theta updates, Adam) at BATCH 1 (not the workload's batch "
              "%d: a bounded sample, images/s = 1 / step time), %s fp32, 64x64x4 latents; %d of the requested %d steps fit "
              "the %.0f s CPU budget" % (args.batch, args.model, n, args.steps, args.cpu_budget_s))
    line = {"impl": "reference", "metric": "SiD-LSG train images/sec", "value": v, "unit": "images/s",
            "n_gpus": args.gpus, "steps": n, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": dict(workload_config(args, 1), sample_batch=1),
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "SiD-LSG iteration (f_psi update + G_theta update + EMA), %s random init, kappa=%.1f, alpha=1, "
                        "%d-step generator, 64x64x4 latents, 77x%d synthetic prompt embeddings"
                        % (args.model, args.kappa, args.num_steps, 768 if args.model != "SD21_BASE" else 1024),
            "batch_per_gpu": args.batch, "micro_batch": args.batch_gpu, "global_batch": args.batch * world,
            "parallelism": "dp%d" % world,
            "l2": "working set per step (3 x 3.4 GB weights + activations) >> 126 MB L2: no flush needed"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    import sid_lsg_b200 as S
    from sid_lsg_b200._lib import KernelTimer
    from sid_lsg_b200.training.step import synth_microbatch, to_device

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        import datetime
        # a short watchdog: a collective mismatch must fail fast instead of holding the GPUs for 10 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    cfg = getattr(S, args.model)
    cd = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    with torch.device(dev):
        base = S.UNet2DConditionModel(cfg, compute_dtype=cd)
    base.flatten_()
    if world > 1:
        dist.broadcast(base.flat.master, src=0)
        base.flat.refresh_shadow()
    import copy
    true_score = base
    fake, G = copy.deepcopy(base), copy.deepcopy(base)
    G_ema = copy.deepcopy(base)
    st = S.SiDLSGStep(true_score, fake, G, G_ema, S.DDPMScheduler(device=dev), cfg_train_fake=args.kappa,
                      cfg_eval_fake=args.kappa, cfg_eval_real=args.kappa, num_steps=args.num_steps)
    rounds = max(1, args.batch // args.batch_gpu)
    mbsz = args.batch // rounds
    total_steps = args.warmup + args.steps

    def host_inputs(step_idx):
        seed = ((step_idx * world + rank) * 1000) % (2 ** 31)
        f = [synth_microbatch(mbsz, cfg, seed + r, None, dropout=True, num_steps=args.num_steps, pinned=True) for r in range(rounds)]
        g = [synth_microbatch(mbsz, cfg, seed + 500 + r, None, num_steps=args.num_steps, pinned=True) for r in range(rounds)]
        return f, g

    def nbytes(mbs):
        n = 0
        for m in mbs:
            for v in m.values():
                n += sum(x.numel() * x.element_size() for x in v) if isinstance(v, list) else v.numel() * v.element_size()
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident run: inputs uploaded before the timed region --------------------------------------
    host = [host_inputs(i) for i in range(total_steps)]
    resident = [([to_device(m, dev) for m in f], [to_device(m, dev) for m in g]) for f, g in host]
    losses = []

    def step_resident(i, off=0):
        f, g = resident[off + i]
        losses.append(st.iteration(f, g, batch_size=args.batch * world))

    for i in range(args.warmup):
        step_resident(i)

    # ---- roofline pass: one eager step with per-call CUDA events on the kernels' own stream (before the graph capture:
    # events per call cannot live inside a captured graph, and the graph's memory pool should not sit beside the eager one)
    peaks = load_peaks()
    timer = None
    if not args.no_roofline:
        # every rank runs the extra step (it contains the gradient allreduces); only rank 0 instruments it
        timer = KernelTimer() if rank == 0 else None
        S.lib.timer = timer
        step_resident(0, args.warmup)
        torch.cuda.synchronize()
        S.lib.timer = None

    graphed = None
    if args.graph:
        f0, g0 = resident[args.warmup]
        graphed = S.GraphedIteration(st, f0, g0, batch_size=args.batch * world)

        def step_resident(i, off=0):  # noqa: F811
            f, g = resident[off + i]
            losses.append(graphed(f, g))

    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = S.lib.launches
    ms_total = timed(lambda i: step_resident(i, args.warmup), args.steps)
    launches = S.lib.launches - n0
    if graphed is not None:
        launches = graphed.calls_per_replay * args.steps    # every replay re-issues the kernels of the captured C-ABI calls
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_step = ms_total / args.steps
    value = args.batch * world / (ms_step / 1e3)
    last_f, last_g = losses[-1]
    finite = bool(torch.isfinite(last_f[0]).item() and torch.isfinite(last_g[0]).item())

    # ---- end-to-end run: pinned host inputs copied in, losses read back, every step ---------------------------
    loss_host = torch.empty((4,), dtype=torch.float32).pin_memory()

    def step_e2e(i):
        f, g = host[args.warmup + i]
        fd = [to_device(m, dev) for m in f]
        gd = [to_device(m, dev) for m in g]
        lf, lg = graphed(fd, gd) if graphed is not None else st.iteration(fd, gd, batch_size=args.batch * world)
        loss_host[:2].copy_(lf, non_blocking=True)
        loss_host[2:].copy_(lg, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the loop reads the losses every iteration (:452, :535)

    if args.no_e2e:
        ms_e2e = float("nan")
    else:
        step_e2e(0)
        ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = args.batch * world / (ms_e2e / 1e3)
    h2d = nbytes(host[args.warmup][0]) + nbytes(host[args.warmup][1])

    line = {"metric": "SiD-LSG train images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, world), "clocks": sampler.result(),
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 16},
            "gpu_launches": launches, "launch_mode": ("cuda graph replay of the whole iteration (kernels per replay = "
                                                      "gpu_launches / steps)" if graphed is not None else "eager"),
            "losses_finite": finite,
            "loss_fake": float(last_f[0].item()), "loss_G": float(last_g[0].item()),
            "step_tflops_algorithmic": STEP_FLOPS_PER_IMAGE * args.batch / (ms_step / 1e3) / 1e12 if args.model == "SD15" else None,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}

    if rank == 0 and not args.no_roofline:
        summ = timer.summary()
        kern = {}
        for key, d in summ.items():
            per = d["ms"] / max(d["launches"], 1)
            if d["unit"] == "flop":
                ach = d["work"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else 0.0
                kern[key] = {"launches": d["launches"], "ms": round(d["ms"], 3), "avg_launch_ms": round(per, 4),
                             "achieved": round(ach, 2), "unit": "TFLOP/s", "frac": round(ach / peaks["tflops_sustained"], 4),
                             "share_of_step": round(d["ms"] / ms_step, 4)}
            else:
                ach = d["work"] / (d["ms"] / 1e3) / 1e9 if d["ms"] > 0 else 0.0
                kern[key] = {"launches": d["launches"], "ms": round(d["ms"], 3), "avg_launch_ms": round(per, 4),
                             "achieved": round(ach, 1), "unit": "GB/s", "frac": round(ach / peaks["hbm_gbs"], 4),
                             "share_of_step": round(d["ms"] / ms_step, 4)}
        tc_keys = [k for k in ("gemm", "conv3x3", "conv3x3_wgrad") if k in summ]
        work = sum(summ[k]["work"] for k in tc_keys)
        ms = sum(summ[k]["ms"] for k in tc_keys)
        n = sum(summ[k]["launches"] for k in tc_keys)
        ach = work / (ms / 1e3) / 1e12 if ms > 0 else 0.0
        line["roofline"] = {"kernel": "gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM conv3x3: linear+conv fwd, dgrad, wgrad)",
                            "bound": "tensor", "achieved": round(ach, 2), "peak": peaks["tflops_sustained"],
                            "unit": "TFLOP/s", "frac": round(ach / peaks["tflops_sustained"], 4),
                            "traffic": (load_traffic() or {}).get("dram_bytes"), "traffic_detail": load_traffic(),
                            "peak_source": peaks["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                            "launches": n, "avg_launch_ms": round(ms / max(n, 1), 4),
                            "algorithmic_flop_per_launch": work / max(n, 1), "share_of_step": round(ms / ms_step, 4),
                            "measured_on": "1 instrumented eager step right before the timed region (per-call CUDA events on the launch stream)",
                            "limiter": "L2->SM operand delivery, 42-47 B/clk/SM in ncu (l1tex__m_xbar2l1tex_read_bytes; profiles/r01_ncu_conv_*.txt): 256-row tiles cut bytes per FLOP 1.4-1.5x; the FLOP roofline is reported because the contract offers hbm|tensor"}
        line["kernels"] = kern
        if args.shapes:
            with open(args.shapes, "w") as f:
                for d in timer.by_shape():
                    rate = d["work"] / max(d["ms"], 1e-9) / (1e9 if d["unit"] == "flop" else 1e6)
                    f.write("%-16s tc=%d %-44s n=%-5d ms=%9.3f avg_us=%8.1f %8.1f %s\n" % (
                        d["name"], d["tc"], d["shape"], d["launches"], d["ms"], 1e3 * d["ms"] / d["launches"], rate,
                        "TFLOP/s" if d["unit"] == "flop" else "GB/s"))

    # ---- CPU baseline (rank 0, N=1 only): the oracle on the host cores, bounded sample ------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del resident
        if graphed is not None:              # the captured iteration owns ~100 GB of activations in its private pool
            graphed.graph.reset()
            graphed.static_f = graphed.static_g = graphed.out = None
        losses.clear()
        torch.cuda.empty_cache()
        del st, true_score, fake, G, G_ema, base
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        parity = {}
        try:
            v, ms, n, cores = cpu_oracle_iteration_time(args, min(args.cpu_budget_s, 150.0), 2, 1, parity=parity)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": "%d timed SiD-LSG iteration(s) at batch 1 (%s fp32 oracle, 64x64x4 latents) after 1 warm-up; %.1f s per iteration"
                                              % (n, args.model, ms / 1e3)}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %r" % (e,)}
        # ---- measured deviation of the benchmarked arithmetic from the fp32 reference arithmetic --------------------
        # the oracle's warm-up iteration (weights, inputs, losses, generated latents) replayed through the CUDA path in
        # the benchmarked dtype: relative errors, printed rather than assumed
        if "loss_G" in parity:
            try:
                nets = []
                for _ in range(3):
                    m = S.UNet2DConditionModel(cfg, compute_dtype=cd)
                    m.load_state_dict(parity["init"])
                    nets.append(m.to(dev).flatten_())
                pst = S.SiDLSGStep(nets[0], nets[1], nets[2], None, S.DDPMScheduler(device=dev), cfg_train_fake=args.kappa,
                                   cfg_eval_fake=args.kappa, cfg_eval_real=args.kappa, num_steps=args.num_steps,
                                   ema_halflife_kimg=0)
                mg = to_device(parity["mb_g"][0], dev)
                with torch.no_grad():
                    img = S.sid_sd_sampler(pst.true_score, mg["z"], S.PromptBatch(mg["cond"], mg["uncond"]),
                                           torch.full((1,), 625, device=dev), pst.sched, num_steps=args.num_steps,
                                           sub_noise=mg.get("sub_noise"))
                lf, lg = pst.iteration([to_device(m, dev) for m in parity["mb_f"]],
                                       [to_device(m, dev) for m in parity["mb_g"]], batch_size=1)
                ref = parity["latents"]
                line["parity_rel_err"] = {
                    "mode": args.dtype, "against": "oracle fp32 iteration at batch 1, same weights and inputs",
                    "latents": float((img.cpu() - ref).norm() / ref.norm()),
                    "loss_fake": abs(float(lf[0].item()) - parity["loss_fake"]) / abs(parity["loss_fake"]),
                    "loss_G": abs(float(lg[0].item()) - parity["loss_G"]) / max(abs(parity["loss_G"]), 1e-30)}
            except Exception as e:  # noqa: BLE001
                line["parity_rel_err"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down order matters: a CUDA graph that captured NCCL collectives must be gone before the communicator is,
        # and destroying the communicator after graph-captured collectives was seen to hang (N = 2, NCCL 2.28.9): the
        # line is out, every rank has passed the barrier, so the processes simply exit.
        if graphed is not None:
            graphed.graph.reset()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
