"""Drop-in for the criterion half of `adv_grpo/pick_score_training.py` (`CLIPCriterionConfig`,
`CLIPCriterion`), the PickScore discriminator loss of `train_pickscore`
(`scripts/train_sd3_fast_pickscore.py:151-183`).  The offline finetuning script in the same reference
file is out of scope."""
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch.nn.modules.loss import _Loss


@dataclass
class CLIPCriterionConfig:
    _target_: str = "trainer.criterions.clip_criterion.CLIPCriterion"
    is_distributed: bool = False
    label_0_column_name: str = "label_0"
    label_1_column_name: str = "label_1"
    input_ids_column_name: str = "input_ids"
    pixels_0_column_name: str = "pixels_0"
    pixels_1_column_name: str = "pixels_1"
    num_examples_per_prompt_column_name: str = "num_examples_per_prompt"
    in_batch_negatives: bool = False


class CLIPCriterion(_Loss):
    def __init__(self, cfg: CLIPCriterionConfig):
        super().__init__()
        self.cfg = cfg

    @staticmethod
    def get_features(model, input_ids, pixels_0_values, pixels_1_values):
        pixels = torch.cat([pixels_0_values, pixels_1_values], dim=0)
        text = model.get_text_features(input_ids=input_ids)
        img = model.get_image_features(pixel_values=pixels)
        img = img / img.norm(dim=-1, keepdim=True)
        text = text / text.norm(dim=-1, keepdim=True)
        i0, i1 = img.chunk(2, dim=0)
        return i0, i1, text

    def calc_loss(self, text_features, image_0_features, image_1_features, logit_scale, label_0, label_1,
                  num_examples_per_prompt, *args, **kwargs):
        if self.cfg.in_batch_negatives or self.cfg.is_distributed:
            raise NotImplementedError("the training scripts use in_batch_negatives=False, is_distributed=False")
        # per prompt: softmax over {real_i, fake_i} of s * <t_i, .>  (row-wise dots; no [B,2B] matmul)
        l0 = logit_scale * (text_features * image_0_features).sum(-1)
        l1 = logit_scale * (text_features * image_1_features).sum(-1)
        pair = torch.stack([l0, l1], dim=-1)
        zeros = torch.zeros(pair.shape[0], dtype=torch.long, device=pair.device)
        loss = label_0 * F.cross_entropy(pair, zeros, reduction="none") + \
            label_1 * F.cross_entropy(pair, zeros + 1, reduction="none")
        is_tie = (torch.as_tensor(label_0) == torch.as_tensor(label_1)).float()
        loss = loss + is_tie * torch.log(torch.tensor(0.5, device=pair.device))
        return loss.mean()

    def forward(self, model, batch):
        c = self.cfg
        i0, i1, t = self.get_features(model, batch[c.input_ids_column_name], batch[c.pixels_0_column_name],
                                      batch[c.pixels_1_column_name])
        return self.calc_loss(t, i0, i1, model.logit_scale.exp(), batch[c.label_0_column_name],
                              batch[c.label_1_column_name], batch[c.num_examples_per_prompt_column_name])
