"""Tuning / evidence run of the planted-positive retrieval protocol (tests/test_engine_gpu.py pins it):
fit on the GPU, then rank K queries x 101 candidates with the sm_100a kernels and with the fp32 CPU oracle on the same weights."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--lr", type=float, default=5e-4)
    ap.add_argument("--queries", type=int, default=8)
    ap.add_argument("--oracle", type=int, default=1)
    ap.add_argument("--classes", type=int, default=8)
    a = ap.parse_args()
    import mvlt_b200
    from mvlt_b200 import retrieval
    from mvlt_b200.synthetic import planted_query
    torch.manual_seed(0)
    lt = {"itm": 1, "mlm": 0, "t2i": 0, "cls": 0}
    m = mvlt_b200.create_model("pvlt_tiny", pretrained=True, num_classes=1000, drop_rate=0.0, drop_path_rate=0.0,
                               drop_block_rate=None, token_hidden_size=768, num_text_tokens=128, loss_type=dict(lt),
                               pretrained_pth="").cuda()
    m.text_embeddings.dropout.p = 0.0
    t = time.time()
    retrieval.fit_planted_itm(m, a.steps, a.batch, a.lr, n_classes=a.classes, log=print)
    torch.cuda.synchronize()
    print("fit time", time.time() - t)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    out = []
    for q in range(a.queries):
        for mode in ("tir", "itr"):
            img, ids = planted_query(q, mode=mode, n_classes=a.classes)
            ranks, logits = retrieval.rank_queries(m, img.cuda(), ids.cuda(), 101)
            lg = logits[0].cpu()
            row = {"q": q, "mode": mode, "rank_gpu": int(ranks[0]), "gap_gpu": float((lg[0, 1] - lg[0, 0]) - (lg[1:, 1] - lg[1:, 0]).max())}
            if a.oracle:
                from oracle import pvlt_oracle as O
                with torch.no_grad():
                    ref = torch.cat([O.forward(sd, img[i:i + 26], ids[i:i + 26], lt, training=False)["itm_logits"].view(-1, 2)
                                     for i in range(0, 101, 26)])
                row["rank_oracle"] = O.retrieval_rank(ref)
                row["max_logit_err"] = float((lg - ref).abs().max())
                order_g = torch.sort(torch.softmax(lg, -1)[:, 1], descending=True)[1]
                order_o = torch.sort(torch.softmax(ref, -1)[:, 1], descending=True)[1]
                row["top10_same"] = bool((order_g[:10] == order_o[:10]).all())
            print(row, flush=True)
            out.append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/planted_retrieval.json", "w"), indent=1)


if __name__ == "__main__":
    main()
