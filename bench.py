#!/usr/bin/env python
"""Benchmark of the OmniBioTA MLM pre-training step (BASELINE.json metric: MLM train tokens/s).

    python bench.py --gpus N --steps K --warmup W            # this implementation (B200 kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A step = one optimizer step of training/train_encoder.py:270-318 on the omnibiote-small shape (8L/1024d/8h,
ctx 1024, vocab 65536, bf16, mini_batch_size 32, dropout 0.1 = the reference default): the global batch of 1024
sequences (BASELINE config 3) is split over the N ranks (strong scaling in the batch, as the reference does:
train_encoder.py:115-118) and accumulated in micro-batches of 32. Synthetic packed mixed nucleotide/peptide token
ids, random-init weights.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
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
N_NONEMB_SMALL = 167_789_568


def flops_per_token(n_layer, n_embd, ctx, n_nonemb):
    """The reference's own estimate (train_encoder.py:360): 6 N + 12 L C T."""
    return 6 * n_nonemb + 12 * n_layer * n_embd * ctx


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
# CPU baseline: the reference's own CPU implementation of the path. /root/reference does not exist on the GPU box,
# so the oracle port (oracle/omnibiota_oracle.py, bit-exact against the reference on CPU: tests/test_oracle.py)
# is timed, on a bounded sample of the workload.
# -------------------------------------------------------------------------------------------------------------------
def cpu_reference_step_tokens_per_s(steps, warmup, batch=2, T=1024, seed=0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import omnibiota_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    cfg = SMALL
    C, V, L, H = cfg["n_embd"], cfg["vocab_size"], cfg["n_layer"], cfg["n_head"]
    wm = C / 24
    bf = torch.bfloat16
    p = {"transformer.wte.weight": torch.randn(V, C).to(bf)}
    lin = lambda o, i: ((torch.rand(o, i) * 2 - 1) / np.sqrt(i)).to(bf)  # nn.Linear default init range
    for l in range(L):
        pre = f"transformer.h.{l}."
        p[pre + "ln_1.weight"] = torch.ones(C, dtype=bf)
        p[pre + "attn.freqs_cis"] = orc.precompute_freqs_cis(C // H, T).real.to(bf)
        p[pre + "attn.c_attn.weight"] = lin(3 * C, C)
        p[pre + "attn.c_proj.weight"] = lin(C, C)
        p[pre + "ln_2.weight"] = torch.ones(C, dtype=bf)
        p[pre + "mlp.c_fc.weight"] = lin(4 * C, C)
        p[pre + "mlp.c_proj.weight"] = lin(C, 4 * C)
    p["transformer.ln_f.weight"] = torch.ones(C, dtype=bf)
    p["lm_head.weight"] = (lin(V, C).float() * np.sqrt(wm)).to(bf)
    names = [k for k in p if "freqs" not in k]
    for k in names:
        p[k].requires_grad_(True)
    m = {k: torch.zeros_like(p[k]) for k in names}
    v = {k: torch.zeros_like(p[k]) for k in names}
    rng = np.random.RandomState(seed)
    ids = torch.from_numpy(synth_ids(batch, T, rng))
    lr = 1e-2 * np.sqrt(1024) / 32

    def one_step(step):
        lm, masked = orc.mlm_mask(ids, rng)
        mask = orc.create_attention_mask(torch.ones(batch, T, T, dtype=bf) * -1e9, ids, padding=False)
        mask = mask.unsqueeze(1).expand(-1, H, -1, -1)
        logits = orc.forward(p, L, H, masked, mask, readout_width_mult=wm)
        loss = orc.mlm_loss(logits, ids, lm, 1)
        grads = torch.autograd.grad(loss, [p[k] for k in names])
        _, coef = orc.clip_grad_norm(grads, 1.0)
        with torch.no_grad():
            for k, g in zip(names, grads):
                g = (g * coef).to(bf)
                lr_k, wd_k = orc.mu_lr_wd(k, p[k].shape, lr, 1e-2, C)
                np_, m[k], v[k] = orc.adamw_step(p[k].detach(), g, m[k], v[k], step, lr_k, wd_k)
                p[k].data.copy_(np_)
        return float(loss)

    for s in range(warmup):
        one_step(s + 1)
    t0 = time.perf_counter()
    for s in range(steps):
        one_step(warmup + s + 1)
    dt = time.perf_counter() - t0
    return batch * T * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    tps, spt = cpu_reference_step_tokens_per_s(steps, warmup)
    cores = os.cpu_count() or 1
    out = {
        "impl": "reference", "metric": "mlm_train_tokens_per_s", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": spt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "omnibiote-small MLM pretraining step (fwd+bwd+clip+muP AdamW), 8L/1024d/8h ctx 1024 "
                               "vocab 65536, bf16, CPU reference path", "sample": "micro-batch of 2 x 1024 tokens per step"},
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of 2 x 1024 tokens (oracle port of training/model.py + "
                                   "train_encoder.py:273-318, torch CPU bf16, all host threads)"},
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# -------------------------------------------------------------------------------------------------------------------
def build_model(device, dropout):
    from omnibiote_b200.model import OmniBioTA, OmniBioTAConfig
    from omnibiote_b200.mup import set_base_shapes
    import copy
    import contextlib
    import io
    import warnings
    cfg = OmniBioTAConfig()
    for k, v in SMALL.items():
        setattr(cfg, k, v)
    cfg.dropout = dropout
    cfg.flash = True
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = OmniBioTA(cfg)
        c2 = copy.copy(cfg); c2.n_embd, c2.n_head = 24, 3
        c3 = copy.copy(cfg); c3.n_embd, c3.n_head = 48, 12
        set_base_shapes(m, OmniBioTA(c2), delta=OmniBioTA(c3))  # train_encoder.py:158-166
        m.to(torch.bfloat16).to(device)
    return m


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

    torch.manual_seed(0)
    model = build_model(device, args.dropout).train()
    if world > 1:  # DDP's initial parameter broadcast (train_encoder.py:185)
        for p_ in model.parameters():
            dist.broadcast(p_.data, 0)
    T, mbs = SMALL["block_size"], args.mini_batch_size
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

    # ---- device-resident timing: `value`
    ops.PROFILE_GEMM = []
    ops.LAUNCHES = 0
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        loss = trainer.step(dev_ids)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ops.LAUNCHES
    gemm_prof, ops.PROFILE_GEMM = ops.PROFILE_GEMM, None
    clock_info = clocks.stop() if rank == 0 else None
    last_loss = float(loss) / trainer.n_accum

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

    # ---- optional second device-resident measurement: the masked-rows-only head (same loss and gradients, the head
    # GEMMs / CE run on the ~15 % of rows inside the MLM mask). Reported under its own key, never as `value`: the
    # roofline fraction and the headline count the head dense over all positions, as the reference executes it.
    ms_mr = float("nan")
    if not args.skip_masked_rows_head:
        from omnibiote_b200 import functional as Fn
        trainer.head_cap = Fn.masked_rows_capacity(mbs * T, trainer.mask_prob)
        trainer.step(dev_ids)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(args.steps):
            trainer.step(dev_ids)
        m1.record()
        barrier()
        ms_mr = m0.elapsed_time(m1)
        trainer.check_head_overflow()
        head_cap, trainer.head_cap = trainer.head_cap, 0

    times = torch.tensor([ms, ms_e2e, ms_mr], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_mr = float(times[0]), float(times[1]), float(times[2])

    if rank == 0:
        tokens_per_step = args.global_batch * T
        value = tokens_per_step * args.steps / (ms / 1e3)
        e2e = tokens_per_step * args.steps / (ms_e2e / 1e3)
        ftok = flops_per_token(SMALL["n_layer"], SMALL["n_embd"], T, N_NONEMB_SMALL)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
        # dominant kernel: the tcgen05 GEMM. achieved = algorithmic FLOPs of all GEMM launches / their summed duration
        g_flops = sum(f for f, _, _ in gemm_prof)
        g_ms = sum(a.elapsed_time(b) for _, a, b in gemm_prof)
        roofline = {
            "kernel": "gemm_bf16_kernel (tcgen05/TMEM/TMA)", "bound": "tensor",
            "achieved": g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else None, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (g_flops / (g_ms / 1e3) / 1e12 / peak_tf) if g_ms > 0 else None,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the c_fc forward shape (32768x4096x1024,
            # ncu --set full, profiles/r01_gemm_cg2_v3.source.txt) against 344.0e6 algorithmic bytes (A + B + D):
            # no operand is re-read from HBM
            "traffic": 306.3e6, "traffic_shape": "32768x4096x1024 (c_fc forward), algorithmic 344.0e6 B",
            "peak_source": peak_src, "launches": len(gemm_prof), "share_of_step": g_ms / ms if ms > 0 else None,
            "step_model_flops_frac_of_2.25PF": value * ftok / world / 2.25e15,
            "step_model_flops_frac_of_measured_sustained": value * ftok / world / (peak_tf * 1e12),
        }
        cpu_tps, cpu_spt = (None, None)
        if world == 1 and not args.skip_cpu_baseline:
            cpu_tps, cpu_spt = cpu_reference_step_tokens_per_s(2, 1)
        out = {
            "metric": "mlm_train_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "omnibiote-small MLM pretraining step (mask + fwd + bwd + clip + muP AdamW), "
                                   "8L/1024d/8h ctx 1024 vocab 65536", "global_batch": args.global_batch,
                       "mini_batch_size": mbs, "grad_accum_per_rank": trainer.n_accum, "seq_len": T,
                       "dropout": args.dropout, "parallelism": f"dp{world}",
                       "l2": "inputs + activations per micro-batch (>4 GiB logits) exceed the 126 MB L2"},
            "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": int(host_ids.numel() * 8) * world,
                    "d2h_bytes_per_step": 4 * world},
            "gpu_launches": launches, "clocks": clock_info, "roofline": roofline, "loss": last_loss,
            "masked_rows_head": None if ms_mr != ms_mr else {
                "value": tokens_per_step * args.steps / (ms_mr / 1e3), "unit": "tokens/s", "head_rows": head_cap,
                "of_rows": mbs * T,
                "note": "same step with the head GEMMs / CE restricted to the rows inside the MLM mask (identical loss "
                        "and gradients); reported separately, not used for `value`, `e2e` or the roofline"},
            "cpu_baseline": None if cpu_tps is None else {
                "value": cpu_tps, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                "sample": "2 steps of 2 x 1024 tokens (oracle port, torch CPU bf16, all host threads)"},
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--global-batch", type=int, default=1024)
    ap.add_argument("--mini-batch-size", type=int, default=32)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-masked-rows-head", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
