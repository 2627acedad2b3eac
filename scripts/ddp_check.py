#!/usr/bin/env python
"""Run under torchrun with N >= 2 GPUs: the data-parallel gradient path (flat buckets, in-place wgrad accumulation,
bucketed NCCL all-reduce overlapped with the last backward, 1/world folding) must reproduce, up to bf16 summation
order, the gradient a single process gets by looping over all ranks' micro-batches (what DDP's mean does in the
reference, training/train_encoder.py:185,284-311). Prints DDP_OK on rank 0."""
from __future__ import annotations

import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))  # test infrastructure: test_shapes_gpu imports the oracle

from test_shapes_gpu import make_model  # noqa: E402
from omnibiote_b200 import functional as Fn  # noqa: E402
from omnibiote_b200 import ops  # noqa: E402
from omnibiote_b200.parallel import FlatGradBuckets, model_buckets  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    vocab, T, mbs, H = 1024, 256, 2, 2
    model = make_model(2, 256, H, vocab, T, seed=rank).train()     # different seeds ...
    for p in model.parameters():
        dist.broadcast(p.data, 0)                                   # ... made identical like DDP's constructor
    # (copy.deepcopy drops Parameter.__dict__, hence mup's infshape — same as in the reference)
    ref_model = make_model(2, 256, H, vocab, T, seed=rank).train()
    ref_model.load_state_dict(model.state_dict())
    g = torch.Generator().manual_seed(123)
    ids_all = torch.randint(20, vocab, (world, mbs, T), generator=g)
    ids_all[:, :, T // 2] = 3
    lm_all = torch.rand(world, mbs, T, generator=g) < 0.15
    ids_all, lm_all = ids_all.to(dev), lm_all.to(dev)

    # data-parallel path
    buckets = FlatGradBuckets(model_buckets(model), None, torch.cuda.Stream())
    Fn.set_grad_sink(buckets)
    with Fn.direct_grad_accumulation(True):
        y, m = ids_all[rank], lm_all[rank]
        lo, hi = ops.doc_mask_intervals(y, 3, False)
        loss, _ = model.mlm_loss(y.masked_fill(m, 2), y, m, attn_mask=ops.MaskSpec(None, mbs, H, T, lo, hi), n_accum=1)
        buckets.arm()
        loss.backward()
    buckets.finish()
    Fn.set_grad_sink(None)
    torch.cuda.synchronize()

    # single-process reference: all ranks' micro-batches, plain autograd accumulation, mean over ranks
    for r in range(world):
        y, m = ids_all[r], lm_all[r]
        lo, hi = ops.doc_mask_intervals(y, 3, False)
        loss, _ = ref_model.mlm_loss(y.masked_fill(m, 2), y, m, attn_mask=ops.MaskSpec(None, mbs, H, T, lo, hi), n_accum=1)
        loss.backward()
    worst = 0.0
    for (n, p), (_, q) in zip(model.named_parameters(), ref_model.named_parameters()):
        a = p.grad.float() / world
        b = q.grad.float() / world
        worst = max(worst, float((a - b).norm() / (b.norm() + 1e-30)))
    ok = worst < 1.5e-2

    # ---- the drop-in module wrapped in stock torch DistributedDataParallel, exactly as train_encoder.py:185 does
    # (no no_sync(): DDP's reducer hooks fire on the custom autograd Functions' parameter gradients and average over
    # ranks); the reference's own loss expression on the dense logits (train_encoder.py:301-305)
    ddp_model = make_model(2, 256, H, vocab, T, seed=rank).train()
    ddp_model.load_state_dict(model.state_dict())
    ddp = torch.nn.parallel.DistributedDataParallel(ddp_model, device_ids=[local])
    y, m = ids_all[rank], lm_all[rank]
    lo, hi = ops.doc_mask_intervals(y, 3, False)
    mask4 = ops.mask_from_intervals(lo, hi).unsqueeze(1).expand(-1, H, -1, -1)
    logits = ddp(y.masked_fill(m, 2), attn_mask=mask4)
    ce = torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)), y.view(-1), reduction="none")
    ce = ce * m.view(-1).float()
    (ce.sum() / m.view(-1).sum()).backward()
    worst_ddp = 0.0
    for (n, p), (_, q) in zip(ddp_model.named_parameters(), ref_model.named_parameters()):
        a = p.grad.float()                      # DDP already averaged over ranks
        b = q.grad.float() / world
        worst_ddp = max(worst_ddp, float((a - b).norm() / (b.norm() + 1e-30)))
    ok = ok and worst_ddp < 2.5e-2

    # ---- MLMTrainer across ranks: identical parameters after a step on every rank, and the step bookkeeping
    # ([loss_sum, n_masked, n_tokens] in ONE async NCCL all-reduce instead of the reference's two Gloo gathers)
    from omnibiote_b200.train import MLMTrainer
    tr_model = make_model(2, 256, H, vocab, T, seed=rank).train()
    tr_model.load_state_dict(model.state_dict())
    tr = MLMTrainer(tr_model, global_batch=world * 2 * mbs, mini_batch_size=mbs, ctx_len=T, lr=1e-3, token_budget=1e9)
    gi = torch.Generator().manual_seed(1000 + rank)
    my_ids = torch.randint(20, vocab, (2 * mbs, T), generator=gi)
    my_ids[:, T // 2] = 3
    my_ids[0, T - 10 * (rank + 1):] = 1                     # some PAD, a different amount on every rank
    my_ids = my_ids.to(dev)
    torch.manual_seed(77 + rank)
    local_loss = tr.step(my_ids)
    stats = tr.read_stats()
    gathered = [torch.zeros(2, device=dev) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([local_loss.float().reshape(1),
                                         (my_ids != 1).sum().float().reshape(1)]))
    want_loss = float(sum(g[0] for g in gathered)) / world
    want_tokens = int(sum(g[1] for g in gathered))
    ok = ok and abs(stats["loss"] - want_loss) <= 1e-5 * abs(want_loss) and stats["n_tokens"] == want_tokens
    ok = ok and stats["tokens_seen"] == want_tokens and stats["n_masked"] > 0
    sums = torch.stack([p.detach().float().sum() for p in tr_model.parameters()])
    lo_, hi_ = sums.clone(), sums.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    ok = ok and bool(torch.equal(lo_, hi_))                 # every rank applied the same update
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"worst relative gradient difference: flat buckets {worst:.2e}, torch DDP wrapper {worst_ddp:.2e}; "
              f"step stats {stats} (want loss {want_loss:.6f}, tokens {want_tokens})")
        print("DDP_OK" if float(flag) == 1.0 else "DDP_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
