"""ORACLE (test infrastructure only, never imported by the product package): CPU restatement of the reference's BERT
word-piece masking on token ids.

Follows /root/reference/mcloader/fashion_gen.py:383-409 (``random_masking_features``) as it is called from
``text_process`` (:334) on the word pieces between [CLS] and [SEP]: the reference mutates a list of token STRINGS and
looks labels up in ``tokenizer.vocab``; with the vendored vocabulary (preweights/bert-base-uncased-vocab.txt, ids in
file order) ``random.choice(list(vocab.items()))[0]`` is the token whose id equals the drawn index, so the same loop on
ids is equivalent. Pinned by tests/golden/token_mask_golden.npz (recorded from the reference method itself).
"""
import random

import numpy as np


def mask_tokens(seed, ori_ids, mask_rate=0.15, sep_id=102, mask_id=103, vocab_size=30522):
    """ori_ids: 1-D int sequence ``[CLS] pieces... [SEP] [PAD]...``. Returns (input_ids, labels) as int64 arrays."""
    rng = random.Random()
    rng.seed(int(seed))
    ids = np.array(ori_ids, dtype=np.int64).copy()
    labels = np.full(ids.shape, -1, dtype=np.int64)
    for i in range(1, len(ids)):
        tok = int(ori_ids[i])
        if tok == sep_id:
            break
        prob = rng.random()                                   # fashion_gen.py:388
        if prob < mask_rate:                                  # :391
            prob /= mask_rate                                 # :392
            if prob < 0.8:                                    # :394-395
                ids[i] = mask_id
            elif prob < 0.9:                                  # :397-398: random.choice(list(vocab.items()))[0]
                ids[i] = rng.choice(range(vocab_size))
            labels[i] = tok                                   # :401-402
    return ids, labels
