"""Oracle: PickScore discriminator loss (`adv_grpo/pick_score_training.py:117-203`,
`in_batch_negatives=False`, not distributed).  Per prompt i the two logits are
s*<t_i, real_i> and s*<t_i, fake_i>; loss = label_0*CE(.,0) + label_1*CE(.,1)
(+ log 0.5 on ties), mean over prompts.
Test infrastructure only (see oracle/__init__.py)."""
import torch
import torch.nn.functional as F


def clip_pair_loss(text_features, image_0_features, image_1_features, logit_scale, label_0, label_1):
    """CLIPCriterion.calc_loss, in_batch_negatives=False branch (pick_score_training.py:117-203)."""
    all_img = torch.cat([image_0_features, image_1_features], dim=0)
    text_logits = logit_scale * text_features @ all_img.T                   # :139
    t0, t1 = text_logits.chunk(2, dim=-1)                                    # :162
    idx = torch.arange(t0.shape[0])
    pair = torch.stack([t0[idx, idx], t1[idx, idx]], dim=-1)                 # :165-167
    lab0 = torch.zeros(pair.shape[0], dtype=torch.long)
    l0 = F.cross_entropy(pair, lab0, reduction="none")                       # :170
    l1 = F.cross_entropy(pair, lab0 + 1, reduction="none")                   # :171
    loss = label_0 * l0 + label_1 * l1                                       # :174
    is_tie = (torch.as_tensor(label_0) == torch.as_tensor(label_1)).float()
    loss = loss + is_tie * torch.log(torch.tensor(0.5))                      # :177-179
    return loss.mean()                                                       # :193
