"""Prompt-only datasets for the distillation loop (`dataset_prompt_text_kwargs.class_name`).

`PromptTextDataset` mirrors the item protocol of the reference's prompt dataset
(/root/reference/training/aesthetics_dataset.py:12-49): `len()` prompts, item = `(dummy image array, prompt string)`,
attributes `name` and `resolution` (used by the snapshot grid, sid_training_loop.py:39-51).  `SyntheticPrompts` is the
offline stand-in (no prompt file needed) used by bench.py / tests."""
import os

import numpy as np
import torch


class PromptTextDataset(torch.utils.data.Dataset):
    FILES = ("aesthetics_6_plus.txt", "aesthetics_625_plus.txt", "aesthetics_65_plus.txt")

    def __init__(self, path, resolution=512, prompt_only=True, **_unused):
        if not prompt_only:
            raise AssertionError("prompt_only must be True: the distillation loop is data-free")
        self.name, self.resolution = "aesthetics", resolution
        fname = path
        if os.path.isdir(path):
            fname = next((os.path.join(path, f) for f in self.FILES if os.path.exists(os.path.join(path, f))), None)
            if fname is None:
                raise FileNotFoundError("no prompt file (%s) under %r" % (", ".join(self.FILES), path))
        with open(fname, "rt") as f:
            self.prompt_list = [row.rstrip("\n") for row in f]

    def __len__(self):
        return len(self.prompt_list)

    def __getitem__(self, idx):
        return np.zeros((1, 4, 4), dtype=np.float32), self.prompt_list[idx]


class SyntheticPrompts(torch.utils.data.Dataset):
    def __init__(self, num_prompts=4096, resolution=512, **_unused):
        self.name, self.resolution = "synthetic", resolution
        self.prompt_list = ["synthetic prompt %06d" % i for i in range(num_prompts)]

    def __len__(self):
        return len(self.prompt_list)

    def __getitem__(self, idx):
        return np.zeros((1, 4, 4), dtype=np.float32), self.prompt_list[idx]
