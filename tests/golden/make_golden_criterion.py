"""Golden vectors G13 for the branches of `CLIPCriterion.calc_loss` the presets do not use
(`adv_grpo/pick_score_training.py:117-203`): `in_batch_negatives=True`, per-example labels with ties, and
`is_distributed=True` (features gathered over 2 gloo ranks), produced by EXECUTING the reference class verbatim.
Build container only:   python tests/golden/make_golden_criterion.py   -> golden_criterion.json / g13_tensors.pt"""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def reference_classes():
    psrc = open(f"{REF}/adv_grpo/pick_score_training.py").read()
    start = psrc.index("@dataclass\nclass CLIPCriterionConfig")
    end = psrc.index("# ====== 数据准备 ======")
    ns = {}
    exec("import torch\nfrom dataclasses import dataclass\nfrom torch.nn.modules.loss import _Loss\n" + psrc[start:end], ns)
    return ns["CLIPCriterion"], ns["CLIPCriterionConfig"]


def inputs():
    g = torch.Generator().manual_seed(13)
    B, D = 6, 32
    t, i0, i1 = (torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1) for _ in range(3))
    label_0 = torch.tensor([1.0, 0.0, 0.5, 1.0, 0.5, 0.0])          # preferred image 0 / image 1 / tie
    label_1 = 1.0 - label_0
    return t, i0, i1, label_0, label_1


def _rank(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.distributed.nn  # noqa: F401  (the reference calls torch.distributed.nn.all_gather)
    Crit, Cfg = reference_classes()
    t, i0, i1, l0, l1 = inputs()
    n = t.shape[0] // world
    sl = slice(rank * n, (rank + 1) * n)
    res = {}
    for ibn in (False, True):
        cfg = Cfg()
        cfg.is_distributed, cfg.in_batch_negatives = True, ibn
        tt = t[sl].clone().requires_grad_(True)
        loss = Crit(cfg).calc_loss(tt, i0[sl], i1[sl], torch.tensor(100.0), l0[sl], l1[sl], torch.ones(n))
        loss.backward()
        res[f"loss_ibn{int(ibn)}"] = loss.item()
        res[f"grad_t_sum_ibn{int(ibn)}"] = tt.grad.double().abs().sum().item()
    out[rank] = res
    dist.destroy_process_group()


def main():
    Crit, Cfg = reference_classes()
    t, i0, i1, l0, l1 = inputs()
    gold = {}
    for ibn in (False, True):
        cfg = Cfg()
        cfg.in_batch_negatives = ibn
        tt = t.clone().requires_grad_(True)
        loss = Crit(cfg).calc_loss(tt, i0, i1, torch.tensor(100.0), l0, l1, torch.ones(t.shape[0]))
        loss.backward()
        gold[f"local_loss_ibn{int(ibn)}"] = loss.item()
        gold[f"local_grad_t_abs_sum_ibn{int(ibn)}"] = tt.grad.double().abs().sum().item()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_rank, args=(2, 29533, out), nprocs=2, join=True)
        gold["distributed_world2"] = {str(k): dict(v) for k, v in out.items()}
    torch.save({"t": t, "i0": i0, "i1": i1, "label_0": l0, "label_1": l1}, os.path.join(OUT, "g13_tensors.pt"))
    with open(os.path.join(OUT, "golden_criterion.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print(json.dumps(gold, indent=1))


if __name__ == "__main__":
    sys.exit(main())
