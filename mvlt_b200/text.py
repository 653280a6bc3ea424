"""Host side of the reference's text pipeline: offline tokenizer, caption -> id rows, ITM pair sampling.

Reference: /root/reference/mcloader/fashion_gen.py
  :46,444,638  ``BertTokenizer.from_pretrained('bert-base-uncased')`` (needs the hub) -> ``load_tokenizer(vocab_file)``
               on the vocabulary the reference vendors (``preweights/bert-base-uncased-vocab.txt``);
  :321-381     ``text_process``: tokenize, keep at most T-2 word pieces, ``[CLS] pieces [SEP] [PAD]...``, segment ids,
               attention mask -> ``encode_captions``; the random 15 % masking inside it (:334, :383-409) runs on the device
               (``mvlt_b200.masking.mask_tokens_batch``, bit-exact for a given ``random.seed``);
  :121-146     image-text-matching pair sampling in ``__getitem__`` -> ``itm_text_index``.
These are integer / string host utilities (no floating-point compute); the model never sees anything but their id rows.
"""
from __future__ import annotations

import random
from typing import Iterable, Sequence, Tuple

import torch


def load_tokenizer(vocab_file: str, do_lower_case: bool = True):
    """``BertTokenizer`` over a local vocabulary file (bert-base-uncased: 30522 entries, [PAD] 0, [UNK] 100, [CLS] 101,
    [SEP] 102, [MASK] 103). No network / hub access."""
    import inspect
    from transformers import BertTokenizer
    if "vocab" in inspect.signature(BertTokenizer.__init__).parameters:      # transformers >= 5: ``vocab`` (dict or path)
        with open(vocab_file, encoding="utf-8") as f:
            vocab = {line.rstrip("\n"): i for i, line in enumerate(f)}
        return BertTokenizer(vocab=vocab, do_lower_case=do_lower_case)
    return BertTokenizer(vocab_file=vocab_file, do_lower_case=do_lower_case)  # transformers 4.x (the reference pins 4.10.2)


def encode_captions(tokenizer, captions: Iterable[str], max_token_length: int = 128):
    """fashion_gen.py:324-358 without the masking: returns ``ori_input_ids``, ``attention_mask``, ``segment_ids`` as int64
    [B, T] host tensors (T = ``max_token_length``). Feed ``ori_input_ids`` to ``masking.mask_tokens_batch`` for the MLM
    corruption and labels."""
    T = int(max_token_length)
    ids_rows, att_rows = [], []
    cls_id, sep_id, pad_id = tokenizer.cls_token_id, tokenizer.sep_token_id, tokenizer.pad_token_id
    for cap in captions:
        pieces = tokenizer.tokenize(cap)
        if len(pieces) > T - 2:                      # :327-329 drop the tail, keep room for [CLS] / [SEP]
            pieces = pieces[:T - 2]
        ids = [cls_id] + tokenizer.convert_tokens_to_ids(pieces) + [sep_id]
        n = len(ids)
        ids_rows.append(ids + [pad_id] * (T - n))    # :346-347 right padding
        att_rows.append([1] * n + [0] * (T - n))     # :359
    ori = torch.tensor(ids_rows, dtype=torch.long)
    att = torch.tensor(att_rows, dtype=torch.long)
    seg = torch.zeros_like(ori)                      # cls / sequence / pad segment ids are all 0 (:322, :341, :349)
    return ori, att, seg


def itm_text_index(rng: random.Random, index: int, size: int) -> Tuple[int, int]:
    """fashion_gen.py:121-146: with probability 1/2 (``random() <= 0.5``) the sample keeps its own caption (label 1);
    otherwise the caption of sample ``index + randint(50, size // 2)`` (wrapped) is used (label 0). ``rng`` is a
    ``random.Random`` (the reference uses the module-level instance), so a seeded stream reproduces the reference's
    draws exactly. Returns (text_index, itm_label)."""
    if rng.random() <= 0.5:
        return int(index), 1
    j = int(index) + rng.randint(50, size // 2)
    if j > size - 1:
        j -= size
    return j, 0


def itm_pairs(seed: int, indices: Sequence[int], size: int):
    """Vector form over a batch of dataset indices with one seeded stream: (text_indices, itm_labels [B, 1] int64)."""
    rng = random.Random(seed)
    pairs = [itm_text_index(rng, i, size) for i in indices]
    return [p[0] for p in pairs], torch.tensor([[p[1]] for p in pairs], dtype=torch.long)
