#!/usr/bin/env python
"""Benchmark of the OmniBioTA MLM pre-training step (BASELINE.json metric: MLM train tokens/s; encode sequences/s;
% of BF16 tensor-core peak).

    python bench.py --gpus N --steps K --warmup W            # this implementation (B200 kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores
    python bench.py --config large --gpus 8 ...              # BASELINE config 4 (32L / 2048 / 16h) instead of small

A step = one optimizer step of training/train_encoder.py:270-318 on the omnibiote-small shape (8L/1024d/8h,
ctx 1024, vocab 65536, bf16, mini_batch_size 32, dropout 0.1 = the reference default): the global batch of 1024
sequences (BASELINE config 3) is split over the N ranks (strong scaling in the batch, as the reference does:
train_encoder.py:115-118) and accumulated in micro-batches of 32. Synthetic packed mixed nucleotide/peptide token
ids, random-init weights.  Prints ONE JSON line (rank 0).

Keys beyond the base contract (all measured in this process, after the headline numbers):
  roofline              per-GEMM CUDA events in a SEPARATE instrumented step (the timed `value` loop carries none)
  encode                BASELINE configs 1 / 5: encode() sequences/s at ctx 1024 and 4096 (all / max / mean), device
                        resident and end to end (pinned ids H2D, result D2H), + the CPU fp32 encode("mean") B=2 leg
  dropin_module_loop    the nn.Module driven by the reference's own loop (dense bias, full logits, ATen CE, torch clip)
  gpu_eager_reference   the reference arithmetic (oracle restatement: torch eager, cuBLAS + SDPA, bf16) timed on the
                        same GPU for the same step schedule: the GPU bar the kernels are measured against
  cpu_baseline          the same on the host cores (bounded sample)
  masked_rows_head      optional head restricted to the masked rows (never used for value / roofline)
  large_config          (N = 8 only) a short run of BASELINE config 4 appended to the default line
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SMALL = dict(n_layer=8, n_embd=1024, n_head=8, block_size=1024, vocab_size=65536)
LARGE = dict(n_layer=32, n_embd=2048, n_head=16, block_size=1024, vocab_size=65536)   # BASELINE config 4
CONFIGS = {"small": SMALL, "large": LARGE}
N_NONEMB_SMALL = 167_789_568


def n_nonembedding(cfg):
    """get_num_params() of training/model.py:213-223 for a config: 12 L C^2 + (2L + 1) C + C V."""
    L, C, V = cfg["n_layer"], cfg["n_embd"], cfg["vocab_size"]
    return 12 * L * C * C + (2 * L + 1) * C + C * V


def flops_per_token(n_layer, n_embd, ctx, n_nonemb):
    """The reference's own estimate (train_encoder.py:360): 6 N + 12 L C T."""
    return 6 * n_nonemb + 12 * n_layer * n_embd * ctx


def encode_flops_per_token(cfg, T):
    """forward without the head: 2 * 12 L C^2 + 4 L C T."""
    return 24 * cfg["n_layer"] * cfg["n_embd"] ** 2 + 4 * cfg["n_layer"] * cfg["n_embd"] * T


def synth_ids(batch, T, rng, vocab=65536, padded=False):
    """Packed documents [tag][body][EOS]... per row, single modality per row (80 % nucleotide / 20 % peptide rows),
    lognormal document lengths (median 200, sigma 1.0, clipped to [8, 4T]) — SURVEY §8d. Synthetic, as stated."""
    ids = np.full((batch, T), 1, dtype=np.int64)
    for b in range(batch):
        tag = 4 if rng.random_sample() < 0.8 else 18
        pos = 0
        while pos < T:
            n = int(np.clip(rng.lognormal(np.log(200.0), 1.0), 8, 4 * T))
            body = rng.randint(20, 65533, size=n)
            doc = np.concatenate([[tag], body, [3]])
            if padded and pos + len(doc) > T:
                break
            doc = doc[: T - pos]
            ids[b, pos:pos + len(doc)] = doc
            pos += len(doc)
    return ids


def workload_config(cfg_name, global_batch, mbs, dropout, world):
    """`config` of the JSON line: identical for the B200 arm and the reference arm of the same invocation."""
    c = CONFIGS[cfg_name]
    return {"workload": f"omnibiote-{cfg_name} MLM pretraining step (mask + fwd + bwd + clip + muP AdamW), "
                        f"{c['n_layer']}L/{c['n_embd']}d/{c['n_head']}h ctx {c['block_size']} vocab {c['vocab_size']}",
            "global_batch": global_batch, "mini_batch_size": mbs, "grad_accum_per_rank": global_batch // world // mbs,
            "seq_len": c["block_size"], "dropout": dropout, "parallelism": f"dp{world}",
            "l2": "inputs + activations per micro-batch (>4 GiB logits) exceed the 126 MB L2"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# -------------------------------------------------------------------------------------------------------------------
# Reference legs: the reference's own arithmetic (oracle/omnibiota_oracle.py, a restatement that is bit-exact against
# the unmodified reference on CPU: tests/test_oracle.py; /root/reference itself does not exist on the GPU box).
# These legs only MEASURE the checker as a baseline; the product path never touches it.
# -------------------------------------------------------------------------------------------------------------------
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import omnibiota_oracle as orc
    return orc


def _oracle_params(cfg, T, device, dtype, seed=0, complex_rope=False):
    """random-init reference state_dict (torch default init ranges + the muP readout rescale)"""
    orc = _oracle()
    torch.manual_seed(seed)
    C, V, L, H = cfg["n_embd"], cfg["vocab_size"], cfg["n_layer"], cfg["n_head"]
    wm = C / 24
    mk = lambda t: t.to(dtype).to(device)
    lin = lambda o, i: mk((torch.rand(o, i) * 2 - 1) / np.sqrt(i))  # nn.Linear default init range
    p = {"transformer.wte.weight": mk(torch.randn(V, C))}
    for l in range(L):
        pre = f"transformer.h.{l}."
        p[pre + "ln_1.weight"] = mk(torch.ones(C))
        f = orc.precompute_freqs_cis(C // H, T)
        p[pre + "attn.freqs_cis"] = f.to(device) if complex_rope else mk(f.real)
        p[pre + "attn.c_attn.weight"] = lin(3 * C, C)
        p[pre + "attn.c_proj.weight"] = lin(C, C)
        p[pre + "ln_2.weight"] = mk(torch.ones(C))
        p[pre + "mlp.c_fc.weight"] = lin(4 * C, C)
        p[pre + "mlp.c_proj.weight"] = lin(C, 4 * C)
    p["transformer.ln_f.weight"] = mk(torch.ones(C))
    p["lm_head.weight"] = mk(((torch.rand(V, C) * 2 - 1) / np.sqrt(C)) * np.sqrt(wm))
    return p, wm


class ReferenceStep:
    """The reference's step schedule (train_encoder.py:270-318) on the oracle restatement, on any torch device:
    micro_batch() = MLM mask + attention mask + forward + loss + backward of `batch` sequences (gradients accumulate
    like autograd's `+=`), optimizer() = clip_grad_norm_(1.0) + AdamW over the muP groups."""

    def __init__(self, cfg, device, batch, seed=0):
        self.orc = _oracle()
        self.cfg, self.device, self.batch = cfg, device, batch
        self.T = cfg["block_size"]
        self.p, self.wm = _oracle_params(cfg, self.T, device, torch.bfloat16, seed)
        self.names = [k for k in self.p if "freqs" not in k]
        for k in self.names:
            self.p[k].requires_grad_(True)
        self.m = {k: torch.zeros_like(self.p[k]) for k in self.names}
        self.v = {k: torch.zeros_like(self.p[k]) for k in self.names}
        self.rng = np.random.RandomState(seed)
        self.ids = torch.from_numpy(synth_ids(batch, self.T, self.rng)).to(device)
        self.lr = 1e-2 * np.sqrt(1024) / 32
        self.t = 0

    def micro_batch(self, n_accum):
        orc, cfg = self.orc, self.cfg
        L, H, T = cfg["n_layer"], cfg["n_head"], self.T
        lm, masked = orc.mlm_mask(self.ids.cpu(), self.rng)          # host numpy RNG, as the reference
        lm, masked = lm.to(self.device), masked.to(self.device)
        mask = orc.create_attention_mask(torch.ones(self.batch, T, T, dtype=torch.bfloat16, device=self.device) * -1e9,
                                         self.ids, padding=False)
        mask = mask.unsqueeze(1).expand(-1, H, -1, -1)
        logits = orc.forward(self.p, L, H, masked, mask, readout_width_mult=self.wm)
        loss = orc.mlm_loss(logits, self.ids, lm, n_accum)
        loss.backward()
        return float(loss.detach())                                   # the reference's loss.item() per micro-batch

    def optimizer(self):
        orc = self.orc
        self.t += 1
        grads = [self.p[k].grad for k in self.names]
        _, coef = orc.clip_grad_norm(grads, 1.0)
        with torch.no_grad():
            for k, g in zip(self.names, grads):
                g = (g * coef).to(torch.bfloat16)
                lr_k, wd_k = orc.mu_lr_wd(k, self.p[k].shape, self.lr, 1e-2, self.cfg["n_embd"])
                np_, self.m[k], self.v[k] = orc.adamw_step(self.p[k].detach(), g, self.m[k], self.v[k], self.t, lr_k, wd_k)
                self.p[k].data.copy_(np_)
                self.p[k].grad = None


def cpu_reference_step(steps, warmup, cfg=SMALL, global_batch=1024, mbs=32, sample_batch=2):
    """Like-for-like CPU arm: the SAME step (global batch 1024 in micro-batches of 32, one optimizer step) on the host
    cores, measured on a bounded sample — each timed `step` is ONE forward + backward of `sample_batch` sequences (a
    1/16 slice of a micro-batch), the optimizer step is timed once — and scaled to the schedule:
    t_step = (global_batch / sample_batch) * mean(t_sample) + t_optimizer. Returns (tokens/s, t_step seconds, detail)."""
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ReferenceStep(cfg, torch.device("cpu"), sample_batch)
    n_accum = global_batch // mbs
    for _ in range(warmup):
        ref.micro_batch(n_accum)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref.micro_batch(n_accum)
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    ref.optimizer()
    t_opt = time.perf_counter() - t0
    t_sample = float(np.mean(ts))
    t_step = (global_batch / sample_batch) * t_sample + t_opt
    detail = {"t_sample_s": t_sample, "t_optimizer_s": t_opt, "sample_batch": sample_batch,
              "samples_per_step": global_batch // sample_batch}
    return global_batch * cfg["block_size"] / t_step, t_step, detail


def cpu_encode_baseline(cfg=SMALL, B=2, T=1024):
    """BASELINE config 1: fp32 encode(method="mean"), batch 2, ctx 1024, CPU, true (complex) rotary; 1 warm-up +
    best of 3."""
    orc = _oracle()
    torch.set_num_threads(os.cpu_count() or 1)
    p, _ = _oracle_params(cfg, T, torch.device("cpu"), torch.float32, seed=0, complex_rope=True)
    ids = torch.from_numpy(synth_ids(B, T, np.random.RandomState(7), padded=True))
    best = float("inf")
    with torch.no_grad():
        for it in range(4):
            t0 = time.perf_counter()
            orc.encode(p, cfg["n_layer"], cfg["n_head"], ids, "mean")
            dt = time.perf_counter() - t0
            if it >= 1:
                best = min(best, dt)
    return {"value": B / best, "unit": "sequences/s", "method": "mean", "batch": B, "ctx_len": T, "dtype": "f32",
            "cores": os.cpu_count(), "kind": "port", "ms_per_batch": best * 1e3}


def gpu_eager_reference(device, cfg=SMALL, global_batch=1024, mbs=32, n_micro=3):
    """The reference arithmetic in torch eager (bf16; cuBLAS GEMMs, SDPA attention, ATen LN / CE / AdamW) on the SAME
    GPU and the same step schedule: fwd + bwd of `n_micro` micro-batches of `mbs` sequences with the reference's host
    mask loops and per-micro-batch loss.item(), one optimizer step; t_step = n_accum * mean(t_mb) + t_opt."""
    out = None
    for try_mbs in (mbs, mbs // 2, mbs // 4):
        try:
            ref = ReferenceStep(cfg, device, try_mbs)
            n_accum = global_batch // try_mbs
            ref.micro_batch(n_accum)                                   # warm-up (cuBLAS / SDPA heuristics, allocator)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n_micro):
                ref.micro_batch(n_accum)
            torch.cuda.synchronize()
            t_mb = (time.perf_counter() - t0) / n_micro
            t0 = time.perf_counter()
            ref.optimizer()
            torch.cuda.synchronize()
            t_opt = time.perf_counter() - t0
            t_step = n_accum * t_mb + t_opt
            out = {"value": global_batch * cfg["block_size"] / t_step, "unit": "tokens/s", "mini_batch_size": try_mbs,
                   "ms_per_micro_batch": t_mb * 1e3, "ms_optimizer": t_opt * 1e3, "ms_per_step": t_step * 1e3,
                   "kind": "port", "what": "oracle restatement of training/model.py + train_encoder.py:270-318 in torch "
                                           "eager bf16 on this GPU (cuBLAS, SDPA, ATen), wall clock with synchronize"}
            del ref
            break
        except torch.OutOfMemoryError:
            out = {"error": f"out of memory at mini_batch_size {try_mbs}"}
        finally:
            torch.cuda.empty_cache()
    return out


def dropin_module_loop(device, cfg=SMALL, global_batch=1024, mbs=32, dropout=0.1, n_micro=3):
    """The path the reference's OWN loop (train_encoder.py:270-318) takes through the drop-in nn.Module, without
    MLMTrainer and without interval masks: a dense additive (b, n_head, t, t) bias handed to `model(idx, attn_mask)`
    (read by every layer's attention kernel), the full (b, t, vocab) logits returned to the caller (4 GiB per
    micro-batch), ATen `F.cross_entropy` + the mask / sum / div arithmetic in torch, `loss.backward()`, a
    `loss.item()` per micro-batch, then torch's `clip_grad_norm_` and `MuAdamW.step()` / `zero_grad()`. The MLM draw and
    the dense mask are built on the device (the reference's host loops are the caller's code, not the module's).
    t_step = n_accum * mean(t_mb) + t_opt, wall clock with synchronize."""
    import torch.nn.functional as F
    from omnibiote_b200 import ops
    from omnibiote_b200.optim import MuAdamW
    from omnibiote_b200.train import mlm_mask
    T, H, V = cfg["block_size"], cfg["n_head"], cfg["vocab_size"]
    model = build_model(device, dropout, cfg)
    model.train()
    opt = MuAdamW(model.parameters(), lr=1e-2 * np.sqrt(global_batch) / 32, weight_decay=1e-2)
    ids = torch.from_numpy(synth_ids(mbs, T, np.random.RandomState(5))).to(device)
    n_accum = global_batch // mbs

    def micro_batch():
        masked, loss_mask = mlm_mask(ids, 0.15)
        lo, hi = ops.doc_mask_intervals(ids, 3, False)
        bias = ops.mask_from_intervals(lo, hi).unsqueeze(1).expand(-1, H, -1, -1)        # (b, n_head, t, t) bf16
        logits = model(masked, attn_mask=bias)
        ce = F.cross_entropy(logits.view(-1, V), ids.view(-1), reduction="none") / n_accum
        ce = ce * loss_mask.view(-1).to(ce.dtype)
        loss = ce.sum() / loss_mask.sum()
        loss.backward()
        return loss.item()

    def optimizer():
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)

    micro_batch(); optimizer()                                      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_micro):
        last = micro_batch()
    torch.cuda.synchronize()
    t_mb = (time.perf_counter() - t0) / n_micro
    t0 = time.perf_counter()
    optimizer()
    torch.cuda.synchronize()
    t_opt = time.perf_counter() - t0
    t_step = n_accum * t_mb + t_opt
    del model, opt
    torch.cuda.empty_cache()
    return {"value": global_batch * T / t_step, "unit": "tokens/s", "ms_per_micro_batch": t_mb * 1e3,
            "ms_optimizer": t_opt * 1e3, "ms_per_step": t_step * 1e3, "loss": last,
            "what": "omnibiote_b200.OmniBioTA driven like train_encoder.py:270-318: dense (b,h,t,t) bias per layer, full "
                    "logits + ATen cross_entropy, loss.item() per micro-batch, torch clip_grad_norm_ + MuAdamW.step()"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    tps, t_step, detail = cpu_reference_step(steps, warmup, cfg, args.global_batch, args.mini_batch_size)
    cores = os.cpu_count() or 1
    sample = (f"{steps} timed + {warmup} warm-up samples, each ONE fwd+bwd of {detail['sample_batch']} x "
              f"{cfg['block_size']} tokens ({detail['t_sample_s'] * 1e3:.0f} ms), + one clip+AdamW step "
              f"({detail['t_optimizer_s'] * 1e3:.0f} ms); step time = {detail['samples_per_step']} x sample + optimizer "
              "(oracle port of training/model.py + train_encoder.py:273-318, torch CPU bf16, all host threads)")
    out = {
        "impl": "reference", "metric": "mlm_train_tokens_per_s", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args.config, args.global_batch, args.mini_batch_size, args.dropout, max(1, args.gpus)),
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# -------------------------------------------------------------------------------------------------------------------
def build_model(device, dropout, cfg_dict=None, on_device=False):
    from omnibiote_b200.model import OmniBioTA, OmniBioTAConfig
    from omnibiote_b200.mup import set_base_shapes
    import copy
    import contextlib
    import io
    import warnings
    cfg = OmniBioTAConfig()
    for k, v in (cfg_dict or SMALL).items():
        setattr(cfg, k, v)
    cfg.dropout = dropout
    cfg.flash = True
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c2 = copy.copy(cfg); c2.n_embd, c2.n_head = 24, 3
        c3 = copy.copy(cfg); c3.n_embd, c3.n_head = 48, 12
        base, delta = OmniBioTA(c2), OmniBioTA(c3)
        if on_device:  # random init straight in HBM (the 1.9 B-parameter config takes minutes on the host)
            with torch.device(device):
                m = OmniBioTA(cfg)
        else:
            m = OmniBioTA(cfg)
        set_base_shapes(m, base, delta=delta)  # train_encoder.py:158-166
        m.to(torch.bfloat16).to(device)
    return m


def read_gemm_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant GEMM (c_fc forward shape,
    32768x4096x1024) from the committed `ncu --set full` capture; None when the file is absent."""
    path = os.path.join(ROOT, "profiles", "r02_gemm_dram_bytes.csv")
    try:
        rd = wr = None
        with open(path) as f:
            for row in csv.DictReader(l for l in f if not l.startswith("#")):
                name, val = row.get("metric"), float(row.get("bytes_per_launch", "nan"))
                if name == "dram__bytes_read.sum":
                    rd = val
                elif name == "dram__bytes_write.sum":
                    wr = val
        return (rd + wr) if rd is not None and wr is not None else None
    except Exception:
        return None


def time_steps(trainer, ids, steps, barrier):
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        loss = trainer.step(ids)
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1), loss


def encode_bench(device, reps=40):
    """BASELINE config 5 (+ the GPU side of config 1): eval(), encode() without attn_mask on variable-length PADDED
    batches; device-resident sequences/s and end-to-end (pinned ids H2D + result D2H inside the timed region)."""
    out = []
    cfg = dict(SMALL)
    cfg["block_size"] = 4096
    model = build_model(device, 0.0, cfg).eval()
    for T, B in ((1024, 32), (4096, 8)):
        host_ids = torch.from_numpy(synth_ids(B, T, np.random.RandomState(7), padded=True)).pin_memory()
        ids = host_ids.to(device)
        ftok = encode_flops_per_token(SMALL, T)
        for method in ("all", "max", "mean"):
            with torch.no_grad():
                for _ in range(10):  # also settles the power state: 8-rep timings drifted 15 % from first to last method
                    res = model.encode(ids, method)
                host_out = torch.empty(res.shape, dtype=res.dtype).pin_memory()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    res = model.encode(ids, method)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                e0.record()
                for _ in range(reps):
                    d = host_ids.to(device, non_blocking=True)
                    res = model.encode(d, method)
                    host_out.copy_(res, non_blocking=True)
                e1.record()
                torch.cuda.synchronize()
                ms_e2e = e0.elapsed_time(e1) / reps
            out.append({"method": method, "ctx_len": T, "batch": B, "value": B / (ms / 1e3), "unit": "sequences/s",
                        "e2e": B / (ms_e2e / 1e3), "h2d_bytes_per_batch": int(host_ids.numel() * 8),
                        "d2h_bytes_per_batch": int(host_out.numel() * 2), "ms_per_batch": ms,
                        "model_tflops": ftok * B * T / (ms / 1e3) / 1e12,
                        "pad_fraction": float((host_ids == 1).float().mean())})
    del model
    torch.cuda.empty_cache()
    return out


def run_large_extra(device, world, rank, barrier, steps=2):
    """BASELINE config 4 (32L / 2048 / 16h, ~1.9 B parameters) for a few steps: tokens/s, model-FLOP fraction and the
    part of the gradient all-reduce (3.76 GB bf16) that is NOT hidden behind the backward."""
    import torch.distributed as dist
    from omnibiote_b200.train import MLMTrainer
    cfg, mbs, gb = LARGE, 32, 1024
    T = cfg["block_size"]
    model = build_model(device, 0.1, cfg, on_device=True).train()
    if world > 1:
        for p_ in model.parameters():
            dist.broadcast(p_.data, 0)
    trainer = MLMTrainer(model, global_batch=gb, mini_batch_size=mbs, ctx_len=T, lr=1e-2, weight_decay=1e-2,
                         token_budget=20e9)
    ids = torch.from_numpy(synth_ids(trainer.batch_size, T, np.random.RandomState(4321 + rank))).to(device)
    trainer.step(ids)
    ms, _ = time_steps(trainer, ids, steps, barrier)
    t = torch.tensor([ms, trainer.buckets.exposed_ms()], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, exposed = float(t[0]), float(t[1])
    ftok = flops_per_token(cfg["n_layer"], cfg["n_embd"], T, n_nonembedding(cfg))
    value = gb * T * steps / (ms / 1e3)
    res = {"config": workload_config("large", gb, mbs, 0.1, world), "value": value, "unit": "tokens/s", "steps": steps,
           "ms_per_step": ms / steps, "params": sum(p.numel() for p in model.parameters()),
           "flops_per_token": ftok, "model_flops_frac_of_2.25PF": value * ftok / world / 2.25e15,
           "allreduce_bytes_per_step": trainer.buckets.flat.numel() * 2, "n_buckets": len(trainer.buckets.bucket_sizes),
           "exposed_allreduce_ms_last_step": exposed,
           "hbm_gb_allocated_peak": torch.cuda.max_memory_allocated(device) / 1e9}
    del trainer, model
    torch.cuda.empty_cache()
    return res


def run_b200(args):
    import torch.distributed as dist
    from omnibiote_b200 import ops
    from omnibiote_b200.train import MLMTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    cfg = CONFIGS[args.config]
    mbs = args.mini_batch_size

    torch.manual_seed(0)
    model = build_model(device, args.dropout, cfg, on_device=args.config == "large").train()
    if world > 1:  # DDP's initial parameter broadcast (train_encoder.py:185)
        for p_ in model.parameters():
            dist.broadcast(p_.data, 0)
    T = cfg["block_size"]
    trainer = MLMTrainer(model, global_batch=args.global_batch, mini_batch_size=mbs, ctx_len=T, lr=1e-2,
                         weight_decay=1e-2, token_budget=20e9)
    per_rank = trainer.batch_size
    rng = np.random.RandomState(1234 + rank)
    host_ids = torch.from_numpy(synth_ids(per_rank, T, rng)).pin_memory()
    dev_ids = host_ids.to(device, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also allocates every workspace and optimizer state)
    for _ in range(max(args.warmup, 3)):
        trainer.step(dev_ids)
    barrier()

    # ---- device-resident timing: `value` (no instrumentation inside the timed loop)
    ops.LAUNCHES = 0
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, loss = time_steps(trainer, dev_ids, args.steps, barrier)
    launches = ops.LAUNCHES
    clock_info = clocks.stop() if rank == 0 else None
    last_loss = float(loss) / trainer.n_accum
    exposed_ar = trainer.buckets.exposed_ms()

    # ---- end-to-end timing through the public API with host buffers: `e2e`
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_loss = torch.empty(1, dtype=torch.float32).pin_memory()
    e0.record()
    for _ in range(args.steps):
        ids = host_ids.to(device, non_blocking=True)       # H2D of this step's inputs from pinned memory
        l = trainer.step(ids)
        host_loss.copy_(l, non_blocking=False)              # D2H read of the step's loss
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    stats = trainer.read_stats()

    # ---- roofline pass: ONE extra step with a CUDA-event pair around every GEMM launch (kept out of `value`)
    ops.PROFILE_GEMM = []
    ms_prof, _ = time_steps(trainer, dev_ids, 1, barrier)
    gemm_prof, ops.PROFILE_GEMM = ops.PROFILE_GEMM, None
    g_flops = sum(r[0] for r in gemm_prof)
    g_ms = sum(r[1].elapsed_time(r[2]) for r in gemm_prof)
    by_shape = {}
    for f, a, b, label in gemm_prof:
        e = by_shape.setdefault(label, [0, 0.0, 0.0])
        e[0] += 1
        e[1] += a.elapsed_time(b)
        e[2] += f
    gemm_by_shape = {k: {"launches": v[0], "ms": round(v[1], 3), "tflops": round(v[2] / (v[1] / 1e3) / 1e12, 1)}
                     for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][1])}

    # ---- optional second device-resident measurement: the masked-rows-only head (same loss and gradients, the head
    # GEMMs / CE run on the ~15 % of rows inside the MLM mask). Reported under its own key, never as `value`: the
    # roofline fraction and the headline count the head dense over all positions, as the reference executes it.
    ms_mr, mr_steps, head_cap = float("nan"), max(1, min(args.steps, 5)), 0
    if not args.skip_masked_rows_head:
        from omnibiote_b200 import functional as Fn
        trainer.head_cap = Fn.masked_rows_capacity(mbs * T, trainer.mask_prob)
        trainer.step(dev_ids)
        ms_mr, _ = time_steps(trainer, dev_ids, mr_steps, barrier)
        trainer.check_head_overflow()
        head_cap, trainer.head_cap = trainer.head_cap, 0
    trainer.check_token_ids()

    times = torch.tensor([ms, ms_e2e, ms_mr, exposed_ar], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_mr, exposed_ar = (float(x) for x in times)

    # ---- extra legs (after every headline number has been taken); they need the HBM the trainer holds
    del trainer, model
    torch.cuda.empty_cache()
    extras = {}
    if args.config == "small" and not args.skip_extras:
        if world == 1:
            try:
                extras["encode"] = {"gpu": encode_bench(device)}
            except Exception as e:  # never lose the headline line to an auxiliary leg
                extras["encode"] = {"error": repr(e)[:300]}
            try:
                extras["gpu_eager_reference"] = gpu_eager_reference(device, cfg, args.global_batch, mbs)
            except Exception as e:
                extras["gpu_eager_reference"] = {"error": repr(e)[:300]}
            try:
                extras["dropin_module_loop"] = dropin_module_loop(device, cfg, args.global_batch, mbs, args.dropout)
            except Exception as e:
                extras["dropin_module_loop"] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
        if world == 8 and not args.skip_large:
            try:
                res = run_large_extra(device, world, rank, barrier)
                extras["large_config"] = res
            except Exception as e:
                extras["large_config"] = {"error": repr(e)[:300]}

    if rank == 0:
        tokens_per_step = args.global_batch * T
        value = tokens_per_step * args.steps / (ms / 1e3)
        e2e = tokens_per_step * args.steps / (ms_e2e / 1e3)
        ftok = flops_per_token(cfg["n_layer"], cfg["n_embd"], T, n_nonembedding(cfg))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
        # dominant kernel: the tcgen05 GEMM. achieved = algorithmic FLOPs of all GEMM launches / their summed duration
        roofline = {
            "kernel": "gemm_bf16_kernel (tcgen05/TMEM/TMA)", "bound": "tensor",
            "achieved": g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else None, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (g_flops / (g_ms / 1e3) / 1e12 / peak_tf) if g_ms > 0 else None,
            # DRAM bytes of ONE launch of the c_fc forward shape (32768x4096x1024) from the committed ncu capture
            # (profiles/r02_gemm_dram_bytes.csv) against 344.0e6 algorithmic bytes (A + B + D)
            "traffic": read_gemm_dram_traffic(),
            "traffic_shape": "32768x4096x1024 (c_fc forward), algorithmic 344.0e6 B",
            "peak_source": peak_src, "launches": len(gemm_prof),
            "measured_in": "one extra instrumented step after the timed region (CUDA events around every GEMM launch)",
            "share_of_step": g_ms / ms_prof if ms_prof > 0 else None,
            # every distinct GEMM of the step (M x N x K, operand layouts N = K-major / T = MN-major, epilogue id):
            # launches, summed CUDA-event time and TFLOP/s inside the step
            "by_shape": gemm_by_shape,
            "step_model_flops_frac_of_2.25PF": value * ftok / world / 2.25e15,
            "step_model_flops_frac_of_measured_sustained": value * ftok / world / (peak_tf * 1e12),
        }
        cpu_leg = None
        if world == 1 and not args.skip_cpu_baseline:
            cpu_tps, cpu_t_step, detail = cpu_reference_step(8, 2, cfg, args.global_batch, mbs)
            cpu_leg = {"value": cpu_tps, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                       "sample": f"8 timed + 2 warm-up samples of ONE fwd+bwd of {detail['sample_batch']} x {T} tokens "
                                 f"({detail['t_sample_s'] * 1e3:.0f} ms each) + one clip+AdamW step "
                                 f"({detail['t_optimizer_s'] * 1e3:.0f} ms), scaled to the same 1024-sequence step "
                                 "(oracle port, torch CPU bf16, all host threads)"}
            if "encode" in extras and args.config == "small":
                try:
                    extras["encode"]["cpu_baseline"] = cpu_encode_baseline()
                except Exception as e:
                    extras["encode"]["cpu_baseline"] = {"error": repr(e)[:300]}
        out = {
            "metric": "mlm_train_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.config, args.global_batch, mbs, args.dropout, world),
            "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": int(host_ids.numel() * 8) * world,
                    "d2h_bytes_per_step": 4 * world},
            "gpu_launches": launches, "clocks": clock_info, "roofline": roofline, "loss": last_loss,
            "step_stats": stats, "exposed_allreduce_ms_last_step": exposed_ar,
            "masked_rows_head": None if ms_mr != ms_mr else {
                "value": tokens_per_step * mr_steps / (ms_mr / 1e3), "unit": "tokens/s", "head_rows": head_cap,
                "of_rows": mbs * T, "steps": mr_steps,
                "note": "same step with the head GEMMs / CE restricted to the rows inside the MLM mask (identical loss "
                        "and gradients); reported separately, not used for `value`, `e2e` or the roofline"},
            "cpu_baseline": cpu_leg,
        }
        out.update(extras)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="small", choices=sorted(CONFIGS))
    ap.add_argument("--global-batch", type=int, default=1024)
    ap.add_argument("--mini-batch-size", type=int, default=32)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-masked-rows-head", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="no encode / gpu_eager_reference / large_config legs")
    ap.add_argument("--skip-large", action="store_true", help="no BASELINE config 4 leg at N = 8")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
