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

    @staticmethod
    def gather_features(features):
        """Autograd-aware all-gather over the ranks (pick_score_training.py:108-111)."""
        import torch.distributed.nn
        return torch.cat(torch.distributed.nn.all_gather(features), dim=0)

    def calc_loss(self, text_features, image_0_features, image_1_features, logit_scale, label_0, label_1,
                  num_examples_per_prompt, *args, **kwargs):
        device = image_0_features.device
        if self.cfg.is_distributed:                                   # pick_score_training.py:135-141
            image_0_features = self.gather_features(image_0_features)
            image_1_features = self.gather_features(image_1_features)
            text_features = self.gather_features(text_features)
            label_0 = self.gather_features(label_0)
            label_1 = self.gather_features(label_1)
        label_0, label_1 = torch.as_tensor(label_0, device=device), torch.as_tensor(label_1, device=device)
        zeros = torch.zeros(text_features.shape[0], dtype=torch.long, device=device)
        # text loss of the pair {image_0_i, image_1_i}: softmax over s * <t_i, .> (row-wise dots; the reference takes
        # the diagonals of a [B, 2B] matmul, :172-181)
        l0 = logit_scale * (text_features * image_0_features).sum(-1)
        l1 = logit_scale * (text_features * image_1_features).sum(-1)
        if self.cfg.in_batch_negatives:                               # :150-170: every other example is a negative
            all_images = torch.cat([image_0_features, image_1_features], dim=0)
            logits_per_image = logit_scale * all_images @ text_features.T
            image_0_logits, image_1_logits = logits_per_image.chunk(2, dim=0)
            text_logits = logit_scale * text_features @ all_images.T
            labels = torch.arange(all_images.shape[0], device=device, dtype=torch.long)
            image_0_labels, image_1_labels = labels.chunk(2, dim=0)
            text_labels = torch.arange(text_features.shape[0], device=device, dtype=torch.long)
            image_loss = label_0 * F.cross_entropy(image_0_logits, text_labels, reduction="none") + \
                label_1 * F.cross_entropy(image_1_logits, text_labels, reduction="none")
            text_0_loss = F.cross_entropy(text_logits, image_0_labels, reduction="none")
            text_1_loss = F.cross_entropy(text_logits, image_1_labels, reduction="none")
        else:
            pair = torch.stack([l0, l1], dim=-1)
            text_0_loss = F.cross_entropy(pair, zeros, reduction="none")
            text_1_loss = F.cross_entropy(pair, zeros + 1, reduction="none")
        text_loss = label_0 * text_0_loss + label_1 * text_1_loss
        is_tie = (label_0 == label_1).float()                         # a tie's ideal loss is log 2: shift it to 0 (:186-190)
        text_loss = text_loss + is_tie * torch.log(torch.tensor(0.5, device=device))
        loss = (image_loss + text_loss) / 2 if self.cfg.in_batch_negatives else text_loss
        return loss.mean()

    def forward(self, model, batch):
        c = self.cfg
        i0, i1, t = self.get_features(model, batch[c.input_ids_column_name], batch[c.pixels_0_column_name],
                                      batch[c.pixels_1_column_name])
        return self.calc_loss(t, i0, i1, model.logit_scale.exp(), batch[c.label_0_column_name],
                              batch[c.label_1_column_name], batch[c.num_examples_per_prompt_column_name])
