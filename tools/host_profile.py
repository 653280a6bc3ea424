"""cProfile of the host side (Python + ctypes) of one pre-training step: where does the enqueue time go?"""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mvlt_b200  # noqa: E402
from mvlt_b200 import masking  # noqa: E402
from mvlt_b200.optim import AdamW, param_groups_no_decay  # noqa: E402
from mvlt_b200.synthetic import make_batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
lt = {"itm": 1, "mlm": 1, "t2i": 1, "cls": 0}
m = mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.1,
                           drop_block_rate=None, token_hidden_size=768, num_text_tokens=128, loss_type=lt,
                           pretrained_pth="").to(dev).train()
opt = AdamW(param_groups_no_decay(m, 0.01), lr=1e-4)
b = make_batch(128, 0)
n = int((b["mlm_labels"] != -1).sum())
b = {k: v.to(dev) for k, v in b.items()}
seeds = torch.arange(128, device=dev)


def step(i):
    img = b["images"]
    x = masking.apply_grid_mask(img, masking.grid_mask_batch(seeds + i)) if i % 2 else img
    total, _ = m(x, b["input_ids"], mlm_labels=b["mlm_labels"], itm_labels=b["itm_labels"], target_images=img, mlm_count=n)
    total.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for i in range(3):
    step(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(4):
    step(i)
    torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
