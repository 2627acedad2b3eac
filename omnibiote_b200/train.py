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
  * the MLM Bernoulli mask comes from a device Philox stream instead of host numpy;
  * step bookkeeping (:311,334-336,350-356: ``loss.item()`` per micro-batch, two Gloo ``all_gather_object`` of the
    loss and of the non-PAD token count per step) is ONE asynchronous NCCL all-reduce of the device vector
    ``[loss_sum, n_masked, n_tokens]`` per step; nothing is read back unless the caller asks (``read_stats``).
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist

from . import _lib, ops
from . import functional as Fn
from .optim import MuAdamW, FusedAdamW
from .parallel import FlatGradBuckets, model_buckets

EOS_TOKEN, MASK_TOKEN, PAD_TOKEN = 3, 2, 1  # training/loader.py:4-6, training/train_encoder.py:20


def mlm_mask(ids: torch.Tensor, prob: float = 0.15, counters: torch.Tensor | None = None):
    """Device version of train_encoder.py:273-279. Returns (masked_ids int64, mask uint8).
    counters (optional fp32[2] device tensor): += {masked positions, non-PAD tokens} (train_encoder.py:350)."""
    if not ids.is_cuda or ids.dtype != torch.int64:
        raise RuntimeError("omnibiote_b200: mlm_mask needs int64 CUDA token ids (there is no CPU path)")
    ids = ids.contiguous()
    masked = torch.empty_like(ids)
    mask = torch.empty(ids.shape, dtype=torch.uint8, device=ids.device)
    seed, off = ops.philox_args(ids.device, 4)
    if counters is not None and (counters.dtype != torch.float32 or counters.numel() < 2 or not counters.is_contiguous()):
        raise RuntimeError("omnibiote_b200: mlm_mask counters must be a contiguous fp32 tensor with 2 elements")
    ops.LAUNCHES += 1
    rc = _lib.load().obt_mlm_mask(ids.data_ptr(), masked.data_ptr(), mask.data_ptr(), ids.numel(), float(prob), seed, off,
                                  PAD_TOKEN, EOS_TOKEN, MASK_TOKEN, 0 if counters is None else counters.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
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
        self.head_overflow = torch.zeros(1, dtype=torch.int32, device=next(model.parameters()).device)  # cumulative
        self._step_overflow = torch.zeros(1, dtype=torch.int32, device=next(model.parameters()).device)

        # train_encoder.py:194-201
        total_iters = max(1, int(token_budget / (self.world * self.batch_size * ctx_len)))
        scaled_lr = lr * math.sqrt(global_batch) / 32
        make = FusedAdamW if force_lr else MuAdamW
        self.optimizer = make(model.parameters(), lr=scaled_lr, weight_decay=weight_decay, betas=betas, eps=eps)
        self.scheduler = torch.optim.lr_scheduler.LinearLR(self.optimizer, start_factor=1.0, end_factor=0.0,
                                                           total_iters=total_iters)
        dev = next(model.parameters()).device
        self.process_group = process_group
        self.comm_stream = torch.cuda.Stream() if (self.world > 1 and dev.type == "cuda") else None
        self.buckets = FlatGradBuckets(model_buckets(model), self._gradient_group(process_group, dev), self.comm_stream)
        # per-step bookkeeping on the device: [loss_sum, n_masked, n_tokens] of this rank, two rotating buffers (the
        # all-reduce of step k runs behind the optimizer of step k and is only waited for at step k+1's gradient sync)
        self._stats_buf = [torch.zeros(3, dtype=torch.float32, device=dev) for _ in range(2)]
        self._stats_work = None
        self.stats = self._stats_buf[0]          # after step(): summed over ranks once the async all-reduce lands
        self.loss_sum = self.stats[0:1]
        self.tokens_seen = torch.zeros(1, dtype=torch.float64, device=dev)  # cumulative non-PAD tokens, all ranks
        self.n_steps = 0
        self.trained_tokens = 0  # host-side count of token POSITIONS (global_batch * ctx_len per step)

    def _gradient_group(self, process_group, dev):
        """Communicator of the gradient all-reduce. The compute kernels are persistent and fill every SM, so a NCCL kernel
        enqueued during the backward only gets SMs at a kernel boundary, where it competes with the next compute kernel;
        on a default-priority stream it mostly lost and the all-reduce ran AFTER the backward (exposed wait ~ the whole
        all-reduce: profiles/r02m_*). A dedicated NCCL communicator whose internal stream is high priority wins those
        boundaries. OBT_NCCL_HIGH_PRIORITY=0 keeps the caller's group."""
        if (self.world <= 1 or dev.type != "cuda" or process_group is not None
                or os.environ.get("OBT_NCCL_HIGH_PRIORITY", "1") == "0" or dist.get_backend() != "nccl"):
            return process_group
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        return dist.new_group(backend="nccl", pg_options=opts)

    def step(self, input_ids: torch.Tensor) -> torch.Tensor:
        """input_ids: this rank's (batch_size, ctx_len) int64 token ids on the device.
        Returns the device scalar sum of the micro-batch losses (the reference's ``cum_loss``)."""
        assert input_ids.shape == (self.batch_size, self.ctx_len), input_ids.shape
        mbs, T, H = self.mini_batch_size, self.ctx_len, self.n_head
        stats = self._stats_buf[self.n_steps & 1]
        stats.zero_()
        masked_ids, mask = mlm_mask(input_ids, self.mask_prob, counters=stats[1:3])
        self.stats, self.loss_sum = stats, stats[0:1]
        if self.head_cap:
            self._step_overflow.zero_()
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
                        self._step_overflow += self.model.head_rows_meta[1:2]
                    if j == self.n_accum - 1:
                        self.buckets.arm()  # overlap the all-reduce with the last micro-batch's backward
                    loss.backward()
                    stats[0:1] += scalars[0:1]
            self.buckets.finish()
        finally:
            Fn.set_grad_sink(None)
        # clip_grad_norm_(1.0) + optimizer.step() (:316-317) in one pass; with the masked-rows head a capacity overflow
        # (incomplete gradients) makes the kernel skip the update instead of applying a corrupt step
        self.optimizer.step(max_norm=self.max_grad_norm, grad_scale=1.0 / self.world, zero_grad=True,
                            skip_flag=self._step_overflow if self.head_cap else None)
        if self.head_cap:
            self.head_overflow += self._step_overflow
        self.scheduler.step()
        local_loss = stats[0:1].clone() if self.world > 1 else stats[0:1]
        self._reduce_stats(stats)
        self.n_steps += 1
        self.trained_tokens += self.global_batch * T
        return local_loss

    def _reduce_stats(self, stats: torch.Tensor) -> None:
        """[loss_sum, n_masked, n_tokens] summed over ranks: one async NCCL all-reduce per step on the side stream
        (replaces the two Gloo all_gather_object calls of train_encoder.py:334-336,350-356), then the cumulative
        non-PAD token counter (:369) — all on the device."""
        if self.world > 1:
            if self.comm_stream is not None:
                self.comm_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.comm_stream):
                    self._stats_work = dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.process_group,
                                                       async_op=True)
                    self._stats_work.wait()  # stream-level wait only: the host does not block
                    self.tokens_seen += stats[2:3].double()
            else:
                self._stats_work = dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
                self._stats_work.wait()
                self.tokens_seen += stats[2:3].double()
        else:
            self.tokens_seen += stats[2:3].double()

    def read_stats(self) -> dict:
        """Host read (synchronises; call at logging cadence): the latest step's loss averaged over ranks — the
        reference's ``np.mean(all_cum_loss)`` — and token counts summed over ranks."""
        if self._stats_work is not None:
            self._stats_work.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        loss_sum, n_masked, n_tokens = self.stats.tolist()
        return {"loss": loss_sum / self.world, "n_masked": int(n_masked), "n_tokens": int(n_tokens),
                "tokens_seen": int(self.tokens_seen.item()), "lr": [g["lr"] for g in self.optimizer.param_groups]}

    def check_head_overflow(self) -> None:
        """Host check (synchronises): raises if any micro-batch had more masked rows than the head capacity (the
        optimizer step of such a batch was skipped on the device)."""
        if self.head_cap and int(self.head_overflow.item()) != 0:
            raise RuntimeError("omnibiote_b200: masked-rows head capacity exceeded; use masked_rows_head=False")

    def check_token_ids(self) -> None:
        """Host check (synchronises): raises if any embedding lookup saw an id outside the vocabulary since the last
        check (the reference's nn.Embedding device-asserts; the gather kernel flags and reads row 0 instead)."""
        ops.check_embedding_ids(next(self.model.parameters()).device)

    # ---- checkpoint / resume (train_encoder.py:174-178,209-223,412-423) -------------------------------------------
    def state_dict(self) -> dict:
        return {"model": self.model.state_dict(), "optimizer": self.optimizer.state_dict(),
                "scheduler": self.scheduler.state_dict(), "trained_tokens": self.trained_tokens,
                "n_steps": self.n_steps, "tokens_seen": float(self.tokens_seen.item())}

    def load_state_dict(self, state: dict) -> None:
        self.model.load_state_dict(state["model"])
        self.optimizer.load_state_dict(state["optimizer"])   # gradients stay views of the flat bucket buffer
        self.scheduler.load_state_dict(state["scheduler"])
        self.trained_tokens = int(state.get("trained_tokens", 0))
        self.n_steps = int(state.get("n_steps", 0))
        self.tokens_seen.fill_(float(state.get("tokens_seen", 0.0)))
