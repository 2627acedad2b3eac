"""Fused gradient clipping + (mu)AdamW on the B200 kernels.

``FusedAdamW`` is a ``torch.optim.Optimizer`` whose ``step`` is one multi-tensor kernel (plus two tiny ones for
the global grad norm) instead of the ~10 foreach passes of ``torch.optim.AdamW``; state is kept in the parameter
dtype (bf16) under the same keys (``exp_avg``, ``exp_avg_sq``, ``step``) as the reference's optimizer pickles
(training/train_encoder.py:199,317,412-423). ``MuAdamW`` applies mup's parameter-group scaling first
(training/train_encoder.py:199). Learning-rate schedulers (``LinearLR``, train_encoder.py:201) work unchanged
because they only rewrite ``param_groups[i]["lr"]``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .mup import mu_param_groups

_META_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("lr", "<f4"), ("wd", "<f4")])


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._init_runtime_state()

    def _init_runtime_state(self):
        self._plan_key = None
        self._plan = None
        self._global_step = 0
        self.last_grad_norm = None  # device tensor [norm, clip_coef] of the latest clip_and_step

    def __setstate__(self, state):
        # torch.save(optimizer) / torch.load (the reference's checkpoint and resume format, train_encoder.py:209,413)
        # keeps only defaults / state / param_groups: the launch plan and its device buffers are rebuilt lazily
        super().__setstate__(state)
        self._init_runtime_state()
        steps = [int(st["step"]) for st in self.state.values() if "step" in st]
        self._global_step = max(steps) if steps else 0

    # -- block plan: which chunk of which tensor each CUDA block processes (depends on shapes only) --------------
    def _build_plan(self, params):
        lib = _lib.load()
        assert lib.obt_opt_meta_bytes() == _META_DTYPE.itemsize
        chunk = lib.obt_opt_chunk_elems()
        blk_tensor, blk_off = [], []
        for i, p in enumerate(params):
            n = p.numel()
            offs = np.arange(0, n, chunk, dtype=np.int64)
            blk_off.append(offs)
            blk_tensor.append(np.full(len(offs), i, dtype=np.int32))
        dev = params[0].device
        plan = {
            "blk_tensor": torch.from_numpy(np.concatenate(blk_tensor)).to(dev),
            "blk_off": torch.from_numpy(np.concatenate(blk_off)).to(dev),
        }
        plan["n_blocks"] = int(plan["blk_tensor"].numel())
        plan["partial"] = torch.empty(plan["n_blocks"], dtype=torch.float32, device=dev)
        plan["norm"] = torch.zeros(2, dtype=torch.float32, device=dev)
        plan["metas_host"] = torch.empty(len(params) * _META_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        plan["metas_dev"] = torch.empty(len(params) * _META_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        return plan

    def _collect(self):
        params, lrs, wds = [], [], []
        betas = eps = None
        for group in self.param_groups:
            if betas is None:
                betas, eps = group["betas"], group["eps"]
            elif betas != group["betas"] or eps != group["eps"]:
                raise RuntimeError("FusedAdamW: all param groups must share betas and eps")
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.bfloat16 or p.grad.dtype != torch.bfloat16:
                    raise RuntimeError("FusedAdamW: parameters and gradients must be bf16 CUDA tensors "
                                       "(there is no CPU / fp32 path)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters and gradients must be contiguous")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                params.append(p)
                lrs.append(float(group["lr"]))
                wds.append(float(group["weight_decay"]))
        return params, lrs, wds, betas, eps

    @torch.no_grad()
    def clip_and_step(self, max_norm: float | None = None, grad_scale: float = 1.0, zero_grad: bool = False,
                      skip_flag: torch.Tensor | None = None):
        """clip_grad_norm_(params, max_norm) + AdamW step in one pass over the gradients.

        grad_scale multiplies every gradient first (e.g. 1/world_size after a sum all-reduce);
        zero_grad=True also clears the gradient buffers in place (keeps their addresses stable);
        skip_flag: optional int32 device scalar, non-zero = leave parameters and moments untouched this step (the
        gradients are known to be incomplete); decided on the device, no host synchronisation."""
        params, lrs, wds, betas, eps = self._collect()
        if not params:
            return None
        key = tuple((id(p), p.numel()) for p in params)
        if key != self._plan_key:
            self._plan = self._build_plan(params)
            self._plan_key = key
        plan = self._plan
        meta = np.zeros(len(params), dtype=_META_DTYPE)
        for i, p in enumerate(params):
            st = self.state[p]
            meta[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                       p.numel(), lrs[i], wds[i])
        if plan.get("copy_evt") is not None:
            plan["copy_evt"].synchronize()  # the pinned staging buffer is reused every step
        plan["metas_host"].numpy()[:] = meta.view(np.uint8)
        plan["metas_dev"].copy_(plan["metas_host"], non_blocking=True)
        plan["copy_evt"] = torch.cuda.Event()
        plan["copy_evt"].record()
        lib = _lib.load()
        stream = torch.cuda.current_stream().cuda_stream
        clip_ptr = 0
        if max_norm is not None:
            rc = lib.obt_grad_norm(plan["metas_dev"].data_ptr(), plan["blk_tensor"].data_ptr(),
                                   plan["blk_off"].data_ptr(), plan["n_blocks"], float(grad_scale), float(max_norm),
                                   plan["partial"].data_ptr(), plan["norm"].data_ptr(), stream)
            _lib.check(rc, "obt_grad_norm")
            clip_ptr = plan["norm"].data_ptr()
            self.last_grad_norm = plan["norm"]
        self._global_step += 1
        for p in params:
            self.state[p]["step"] += 1
        step = int(self.state[params[0]]["step"].item())
        if skip_flag is not None and (not skip_flag.is_cuda or skip_flag.dtype != torch.int32):
            raise RuntimeError("FusedAdamW: skip_flag must be an int32 CUDA tensor")
        rc = lib.obt_adamw_step(plan["metas_dev"].data_ptr(), plan["blk_tensor"].data_ptr(), plan["blk_off"].data_ptr(),
                                plan["n_blocks"], clip_ptr, 0 if skip_flag is None else skip_flag.data_ptr(),
                                float(grad_scale), 1.0, float(betas[0]), float(betas[1]), float(eps), step,
                                int(zero_grad), stream)
        _lib.check(rc, "obt_adamw_step")
        return self.last_grad_norm

    @torch.no_grad()
    def step(self, closure=None, *, max_norm: float | None = None, grad_scale: float = 1.0, zero_grad: bool = False,
             skip_flag: torch.Tensor | None = None):
        """``optimizer.step()`` of the reference (train_encoder.py:317); the keyword arguments fold the preceding
        ``clip_grad_norm_`` (:316) and the DDP mean into the same pass (see clip_and_step). Going through ``step`` keeps
        torch's LR-scheduler bookkeeping (it wraps this method) intact."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.clip_and_step(max_norm, grad_scale=grad_scale, zero_grad=zero_grad, skip_flag=skip_flag)
        return loss


def MuAdamW(params, impl=FusedAdamW, decoupled_wd=False, **kwargs):
    """mup.MuAdamW (training/train_encoder.py:199) on the fused kernel: matrix-like parameters get
    lr / width_mult and weight_decay * width_mult; groups are ordered [matrix-like..., vector-like]."""
    groups = mu_param_groups(params, lr=kwargs["lr"], weight_decay=kwargs.get("weight_decay", 0.0),
                             decoupled_wd=decoupled_wd)
    return impl(groups, **kwargs)
