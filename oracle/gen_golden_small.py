"""TEST INFRASTRUCTURE ONLY — full-size golden fixture from the UNMODIFIED reference (SURVEY §8c sets (ii) and (iii)).

Run in the build container (needs /root/reference):  ``python oracle/gen_golden_small.py``  (~2 minutes, ~12 GB RAM)
Writes tests/golden/small_shape.pt: the omnibiote-small shape (8L / 1024 / 8h, vocab 65536, bf16) with weights from
oracle/golden_recipe.py, run through /root/reference/training/model.py on CPU:
  (iii) B = 2, T = 1024 packed documents with the reference's own create_attention_mask: ln_f embeddings, every 64th
        logit column, the MLM loss of train_encoder.py:301-305 (n_accum = 2), and per parameter the gradient norm
        and 4096 sampled gradient elements;
  (ii)  B = 1, t in {6, 137, 1023}, no mask (the evals' call pattern): ln_f embeddings and every 64th logit column.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  (imports the reference's model.py)
import golden_recipe as rec  # noqa: E402
import omnibiota_oracle as orc  # noqa: E402

L, C, H, V, T = 8, 1024, 8, 65536, 1024
STRIDE = 64


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    model = gg.build_reference(L, C, H, V, T, torch.bfloat16, seed=0)
    rec.load_recipe_weights(model, C)
    wm = float(model.lm_head.width_mult())
    rng = np.random.RandomState(2024)
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    ids = torch.from_numpy(bench.synth_ids(2, T, rng))
    lm, masked = orc.mlm_mask(ids, np.random.RandomState(11))
    mask3 = gg.create_ref_mask(torch.ones((2, T, T), dtype=torch.bfloat16) * -1e9, ids, False)
    out = {"cfg": dict(n_layer=L, n_embd=C, n_head=H, vocab_size=V, block_size=T), "width_mult": wm, "stride": STRIDE,
           "ids": ids, "mlm_mask": lm, "ids_masked": masked}
    # the interval form of the mask (2 x int32 [B,T]) instead of the dense 2 x 1024 x 1024 tensor; the dense mask is
    # rebuilt by the test and must reproduce these intervals exactly
    lo, hi = orc.mask_intervals(mask3)
    out["mask_lo"], out["mask_hi"] = lo, hi
    model.train()
    logits = model.forward(masked, attn_mask=mask3.unsqueeze(1).expand(-1, H, -1, -1))
    loss = torch.nn.functional.cross_entropy(logits.view(-1, V), ids.view(-1), reduction="none") / 2
    loss *= lm.view(-1).float()
    loss = loss.sum() / lm.view(-1).sum()
    loss.backward()
    out["loss"] = loss.detach().clone()
    out["logits_sub"] = logits.detach()[..., ::STRIDE].clone()
    out["grads"] = {}
    for n, p in model.named_parameters():
        g = p.grad.detach().float().reshape(-1)
        out["grads"][n] = {"norm": float(g.double().norm()), "sample": p.grad.detach().reshape(-1)[rec.sample_indices(n, g.numel())].clone()}
    model.eval()
    with torch.no_grad():
        out["emb"] = model(masked, attn_mask=mask3.unsqueeze(1).expand(-1, H, -1, -1), return_embeddings=True).clone()
        out["odd"] = {}
        for t in (6, 137, 1023):
            x = ids[:1, :t].clone()
            out["odd"][t] = {"ids": x, "emb": model(x, return_embeddings=True).clone(),
                             "logits_sub": model(x)[..., ::STRIDE].clone(),
                             "encode_mean": model.encode(x, method="mean").clone(),
                             "encode_max": model.encode(x, method="max").clone()}
    path = os.path.join(gg.OUT, "small_shape.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path) // 1024, "KiB", "loss", float(out["loss"]))


if __name__ == "__main__":
    main()
