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
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"worst relative gradient difference {worst:.2e}")
        print("DDP_OK" if float(flag) == 1.0 else "DDP_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
