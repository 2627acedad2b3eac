"""The data-parallel MLM pre-training step of the reference (training/train_encoder.py:270-318) on the B200 kernels.

One ``MLMTrainer.step(input_ids)`` = one optimizer step over this rank's share of the global batch:
MLM masking (:273-279) -> for each micro-batch: document attention mask (:290-292), forward + loss (:296-305),
backward (:308) with bf16 gradient accumulation -> gradient all-reduce over ranks (DDP, :185) -> clip_grad_norm_(1.0)
(:316) -> MuAdamW step (:317) -> LinearLR step (:318).

Differences from the reference's schedule that do not change the arithmetic contract:
  * no host synchronisation inside the step (the reference calls ``loss.item()`` per micro-batch and builds masks
    with host-synchronising loops); the summed loss stays on the device until the caller reads it;
  * gradients are all-reduced once per optimizer step (sum, then 1/world folded into the optimizer) instead of
    after every micro-batch; the result differs from the reference's only by bf16 summation order;
  * the MLM Bernoulli mask comes from a device Philox stream instead of host numpy.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import _lib, ops
from . import functional as Fn
from .optim import MuAdamW, FusedAdamW
from .parallel import FlatGradBuckets, model_buckets

EOS_TOKEN, MASK_TOKEN, PAD_TOKEN = 3, 2, 1  # training/loader.py:4-6, training/train_encoder.py:20


def mlm_mask(ids: torch.Tensor, prob: float = 0.15):
    """Device version of train_encoder.py:273-279. Returns (masked_ids int64, mask uint8)."""
    ids = ids.contiguous()
    masked = torch.empty_like(ids)
    mask = torch.empty(ids.shape, dtype=torch.uint8, device=ids.device)
    seed, off = ops.philox_args(ids.device, 4)
    rc = _lib.load().obt_mlm_mask(ids.data_ptr(), masked.data_ptr(), mask.data_ptr(), ids.numel(), float(prob), seed, off,
                                  PAD_TOKEN, EOS_TOKEN, MASK_TOKEN, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "obt_mlm_mask")
    return masked, mask


class MLMTrainer:
    def __init__(self, model, *, global_batch: int, mini_batch_size: int, ctx_len: int, lr: float = 1e-2,
                 weight_decay: float = 1e-2, betas=(0.9, 0.999), eps: float = 1e-8, token_budget: float = 250e9,
                 use_padding: bool = False, mask_prob: float = 0.15, max_grad_norm: float = 1.0, force_lr: bool = False,
                 process_group=None, masked_rows_head: bool = False):
        self.model = model
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        assert global_batch % self.world == 0, "Batch size must be divisible by the number of processes."
        self.global_batch = global_batch
        self.batch_size = global_batch // self.world           # per-rank batch (train_encoder.py:118)
        assert self.batch_size % mini_batch_size == 0
        self.mini_batch_size = mini_batch_size
        self.n_accum = self.batch_size // mini_batch_size       # train_encoder.py:284
        self.ctx_len = ctx_len
        self.use_padding = use_padding
        self.mask_prob = mask_prob
        self.max_grad_norm = max_grad_norm
        self.n_head = model.transformer.h[0].attn.n_head
        # optional masked-rows-only head (same loss and gradients; see functional.HeadLossMaskedRowsFunction): capacity
        # = mean + 8 sigma of the Binomial number of masked rows; an overflow is counted on the device and raises at
        # the next check_head_overflow()
        self.head_cap = Fn.masked_rows_capacity(mini_batch_size * ctx_len, mask_prob) if masked_rows_head else 0
        self.head_overflow = torch.zeros(1, dtype=torch.int32, device=next(model.parameters()).device)

        # train_encoder.py:194-201
        total_iters = max(1, int(token_budget / (self.world * self.batch_size * ctx_len)))
        scaled_lr = lr * math.sqrt(global_batch) / 32
        make = FusedAdamW if force_lr else MuAdamW
        self.optimizer = make(model.parameters(), lr=scaled_lr, weight_decay=weight_decay, betas=betas, eps=eps)
        self.scheduler = torch.optim.lr_scheduler.LinearLR(self.optimizer, start_factor=1.0, end_factor=0.0,
                                                           total_iters=total_iters)
        comm_stream = torch.cuda.Stream() if (self.world > 1 and next(model.parameters()).is_cuda) else None
        self.buckets = FlatGradBuckets(model_buckets(model), process_group, comm_stream)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=next(model.parameters()).device)
        self.trained_tokens = 0

    def step(self, input_ids: torch.Tensor) -> torch.Tensor:
        """input_ids: this rank's (batch_size, ctx_len) int64 token ids on the device.
        Returns the device scalar sum of the micro-batch losses (the reference's ``cum_loss``)."""
        assert input_ids.shape == (self.batch_size, self.ctx_len), input_ids.shape
        mbs, T, H = self.mini_batch_size, self.ctx_len, self.n_head
        masked_ids, mask = mlm_mask(input_ids, self.mask_prob)
        self.loss_sum.zero_()
        Fn.set_grad_sink(self.buckets)
        try:
            with Fn.direct_grad_accumulation(True):
                for j in range(self.n_accum):
                    x = masked_ids[j * mbs:(j + 1) * mbs]
                    y = input_ids[j * mbs:(j + 1) * mbs]
                    m = mask[j * mbs:(j + 1) * mbs]
                    lo, hi = ops.doc_mask_intervals(y, EOS_TOKEN, self.use_padding)
                    spec = ops.MaskSpec(None, mbs, H, T, lo, hi)
                    loss, scalars = self.model.mlm_loss(x, y, m, attn_mask=spec, n_accum=self.n_accum,
                                                        masked_rows_cap=self.head_cap)
                    if self.head_cap:
                        self.head_overflow += self.model.head_rows_meta[1:2]
                    if j == self.n_accum - 1:
                        self.buckets.arm()  # overlap the all-reduce with the last micro-batch's backward
                    loss.backward()
                    self.loss_sum += scalars[0:1]
            self.buckets.finish()
        finally:
            Fn.set_grad_sink(None)
        self.optimizer.clip_and_step(self.max_grad_norm, grad_scale=1.0 / self.world, zero_grad=True)
        self.scheduler.step()
        self.trained_tokens += self.global_batch * T
        return self.loss_sum

    def check_head_overflow(self) -> None:
        """Host check (synchronises): raises if any micro-batch had more masked rows than the head capacity."""
        if self.head_cap and int(self.head_overflow.item()) != 0:
            raise RuntimeError("omnibiote_b200: masked-rows head capacity exceeded; use masked_rows_head=False")
