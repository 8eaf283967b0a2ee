"""Helpers to read the committed fixtures written by tests/golden/make_golden.py."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_PROMPTS = 40
GRID_TE_CALLS = 14  # fakes_init export before the loop: 28 prompts / batch_gpu 2


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def embedding_table(d):
    g = torch.Generator().manual_seed(777)
    return torch.randn([N_PROMPTS + 1, 77, d], generator=g)


def parse_loop(fix, d):
    """-> list of iterations, each (mb_f, mb_g); a micro-batch is the dict oracle.step expects
    (z, noise, t, cond, uncond[, sub_noise]) built from the draws the reference made."""
    table = embedding_table(d)
    rounds = fix["batch"] // fix["batch_gpu"]
    draws = [x for x in fix["draws"] if not (x[1] == "sid_training_loop.py" and x[2] == 267)]
    te = fix["te_calls"][GRID_TE_CALLS:]
    iters = []
    di = 0
    ti = 0
    n_iter = len(te) // (rounds * 8)
    for _ in range(n_iter):
        phases = []
        for phase, n_te in (("f", 3), ("g", 5)):
            mbs = []
            for _r in range(rounds):
                m = {"sub_noise": []}
                zline, nline, tline = (398, 399, 413) if phase == "f" else (479, 480, 484)
                while True:
                    name, src, line, val = draws[di]
                    if src == "sid_sd_util.py":
                        m["sub_noise"].append(val)
                    elif line == 394:
                        pass
                    elif line == zline:
                        if "z" in m:
                            break
                        m["z"] = val
                    elif line == nline:
                        m["noise"] = val
                    elif line == tline:
                        m["t"] = val
                    else:
                        break
                    di += 1
                    if di == len(draws):
                        break
                    # a micro-batch is complete once z, noise, t and all sub-noise draws are in
                    if all(k in m for k in ("z", "noise", "t")) and len(m["sub_noise"]) == fix["num_steps"] - 1:
                        break
                ids = te[ti]
                m["cond"] = table[ids]
                m["uncond"] = table[te[ti + 2]]
                m["cond_ids"] = ids
                ti += n_te
                mbs.append(m)
            phases.append(mbs)
        iters.append(tuple(phases))
    assert di == len(draws) and ti == len(te), (di, len(draws), ti, len(te))
    return iters
