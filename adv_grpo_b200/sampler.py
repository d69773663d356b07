"""`DistributedKRepeatSampler` and `TextPromptDataset` of the training scripts
(`scripts/train_sd3_fast_pickscore.py:50-66,87-129`): every iteration draws m = world*batch/k distinct
prompts with a generator seeded `seed + epoch` (identical on every rank), repeats each k times,
shuffles, and hands rank r its slice.  Pure index arithmetic on the host (ints, negligible time)."""
import os

import torch


class TextPromptDataset:
    def __init__(self, dataset, split="train"):
        path = os.path.join(dataset, f"{split}.txt")
        with open(path, "r") as f:
            self.prompts = [line.strip() for line in f.readlines()]

    def __len__(self):
        return len(self.prompts)

    def __getitem__(self, idx):
        return {"prompt": self.prompts[idx], "metadata": {}}

    @staticmethod
    def collate_fn(examples):
        return [e["prompt"] for e in examples], [e["metadata"] for e in examples]


class DistributedKRepeatSampler:
    def __init__(self, dataset, batch_size, k, num_replicas, rank, seed=0):
        self.dataset, self.batch_size, self.k = dataset, batch_size, k
        self.num_replicas, self.rank, self.seed = num_replicas, rank, seed
        self.total_samples = num_replicas * batch_size
        if self.total_samples % k != 0:
            raise AssertionError(f"k can not divide n*b, k{k}-num_replicas{num_replicas}-batch_size{batch_size}")
        self.m = self.total_samples // k
        self.epoch = 0

    def indices_for_epoch(self, epoch):
        g = torch.Generator()
        g.manual_seed(self.seed + epoch)
        picked = torch.randperm(len(self.dataset), generator=g)[: self.m].tolist()
        repeated = [i for i in picked for _ in range(self.k)]
        order = torch.randperm(len(repeated), generator=g).tolist()
        shuffled = [repeated[i] for i in order]
        return [shuffled[r * self.batch_size:(r + 1) * self.batch_size] for r in range(self.num_replicas)]

    def __iter__(self):
        while True:
            yield self.indices_for_epoch(self.epoch)[self.rank]

    def set_epoch(self, epoch):
        self.epoch = epoch
