"""`encode_prompt` with the reference's name, signature and return convention
(`adv_grpo/diffusers_patch/train_dreambooth_lora_sd3.py:13-144`, imported by both training scripts,
`scripts/train_sd3_fast_pickscore.py:22`): CLIP-L + CLIP-G (`hidden_states[-2]`, pooled text_embeds) and T5
(`encoder(ids)[0]`) -> `(prompt_embeds [B*n, 77 + max_sequence_length, 4096], pooled [B, 768 + 1280])`.
The encoders are any objects with the transformers call convention; `adv_grpo_b200.text_encoders` provides them
on the B200 kernels.  Tokenizers are the caller's (no tokenizer files ship with this repo): pass them as in the
reference, or pass `text_input_ids_list` and `tokenizers=[None, None, None]`."""
import torch


def _ids(tokenizer, prompt, max_length, text_input_ids, **kw):
    if tokenizer is not None:
        return tokenizer(prompt, padding="max_length", max_length=max_length, truncation=True, return_tensors="pt",
                         **kw).input_ids
    if text_input_ids is None:
        raise ValueError("text_input_ids must be provided when the tokenizer is not specified")
    return text_input_ids


def _encode_prompt_with_t5(text_encoder, tokenizer, max_sequence_length, prompt=None, num_images_per_prompt=1,
                           device=None, text_input_ids=None):
    prompt = [prompt] if isinstance(prompt, str) else prompt
    batch_size = len(prompt)
    ids = _ids(tokenizer, prompt, max_sequence_length, text_input_ids, add_special_tokens=True)
    prompt_embeds = text_encoder(ids.to(device))[0]
    prompt_embeds = prompt_embeds.to(dtype=text_encoder.dtype, device=device)
    _, seq_len, _ = prompt_embeds.shape
    prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1)
    return prompt_embeds.view(batch_size * num_images_per_prompt, seq_len, -1)


def _encode_prompt_with_clip(text_encoder, tokenizer, prompt, device=None, text_input_ids=None, num_images_per_prompt=1):
    prompt = [prompt] if isinstance(prompt, str) else prompt
    batch_size = len(prompt)
    ids = _ids(tokenizer, prompt, 77, text_input_ids)
    out = text_encoder(ids.to(device), output_hidden_states=True)
    pooled_prompt_embeds = out[0]
    prompt_embeds = out.hidden_states[-2].to(dtype=text_encoder.dtype, device=device)
    _, seq_len, _ = prompt_embeds.shape
    prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1)
    return prompt_embeds.view(batch_size * num_images_per_prompt, seq_len, -1), pooled_prompt_embeds


def encode_prompt(text_encoders, tokenizers, prompt, max_sequence_length, device=None, num_images_per_prompt=1,
                  text_input_ids_list=None):
    prompt = [prompt] if isinstance(prompt, str) else prompt
    clip_embeds, clip_pooled = [], []
    for i, (tokenizer, text_encoder) in enumerate(zip(tokenizers[:2], text_encoders[:2])):
        e, pooled = _encode_prompt_with_clip(
            text_encoder, tokenizer, prompt, device=device if device is not None else text_encoder.device,
            num_images_per_prompt=num_images_per_prompt,
            text_input_ids=text_input_ids_list[i] if text_input_ids_list else None)
        clip_embeds.append(e)
        clip_pooled.append(pooled)
    clip_prompt_embeds = torch.cat(clip_embeds, dim=-1)
    pooled_prompt_embeds = torch.cat(clip_pooled, dim=-1)
    t5_prompt_embed = _encode_prompt_with_t5(
        text_encoders[-1], tokenizers[-1], max_sequence_length, prompt=prompt, num_images_per_prompt=num_images_per_prompt,
        text_input_ids=text_input_ids_list[-1] if text_input_ids_list else None,
        device=device if device is not None else text_encoders[-1].device)
    clip_prompt_embeds = torch.nn.functional.pad(clip_prompt_embeds,
                                                 (0, t5_prompt_embed.shape[-1] - clip_prompt_embeds.shape[-1]))
    return torch.cat([clip_prompt_embeds, t5_prompt_embed], dim=-2), pooled_prompt_embeds


def compute_text_embeddings(prompt, text_encoders, tokenizers, max_sequence_length, device):
    """`scripts/train_sd3_fast_pickscore.py:186-193`."""
    with torch.no_grad():
        prompt_embeds, pooled_prompt_embeds = encode_prompt(text_encoders, tokenizers, prompt, max_sequence_length)
        return prompt_embeds.to(device), pooled_prompt_embeds.to(device)
