"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):  ``python oracle/gen_golden.py``
It imports ``/root/reference/training/model.py`` as-is (with ``oracle/mup`` standing in for the missing third-party
``mup`` package), builds tiny random-init models exactly like ``training/train_encoder.py:144-170`` does
(target / base n_embd=24 / delta n_embd=48 -> set_base_shapes -> .to(dtype)), and records inputs, weights, outputs
and gradients. The fixtures travel to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OBT_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)                           # oracle/mup
sys.path.insert(0, os.path.join(REF, "training"))  # reference model.py

from model import OmniBioTA, OmniBioTAConfig  # noqa: E402  (the reference)
from mup import set_base_shapes, MuAdamW  # noqa: E402
import omnibiota_oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
EOS, MASK, PAD = 3, 2, 1


def build_reference(n_layer, n_embd, n_head, vocab, block_size, dtype, seed=0):
    torch.manual_seed(seed)
    np.random.seed(seed)
    cfg = OmniBioTAConfig()
    cfg.vocab_size, cfg.dropout, cfg.block_size = vocab, 0.0, block_size
    cfg.n_embd, cfg.n_layer, cfg.n_head = n_embd, n_layer, n_head
    cfg.flash, cfg.checkpoint_freq = True, 0
    m = OmniBioTA(cfg)
    cfg.n_embd, cfg.n_head = 24, 3
    base = OmniBioTA(cfg)
    cfg.n_embd, cfg.n_head = 48, 12
    delta = OmniBioTA(cfg)
    set_base_shapes(m, base, delta=delta)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.to(dtype)
    return m


def synth_ids(B, T, vocab, rng, padded):
    """Packed documents [tag][body]...[EOS]; padded rows end with whole documents then PAD (loader.py:139-146)."""
    ids = np.full((B, T), PAD, dtype=np.int64)
    for b in range(B):
        pos = 0
        while pos < T:
            n = int(rng.randint(4, max(5, T // 2)))
            doc = [int(rng.choice([4, 6, 18]))] + list(rng.randint(20, vocab, size=n)) + [EOS]
            if padded and pos + len(doc) > T:
                break
            doc = doc[: T - pos]
            ids[b, pos:pos + len(doc)] = doc
            pos += len(doc)
    return torch.from_numpy(ids)


def ref_masks(ids, n_head, dtype):
    B, T = ids.shape
    out = {}
    m = torch.ones((B, T, T), dtype=dtype) * -1e9
    out["doc"] = orc_to4(create_ref_mask(m, ids, False), n_head)
    m = torch.ones((B, T, T), dtype=dtype) * -1e9
    out["docpad"] = orc_to4(create_ref_mask(m, ids, True), n_head)
    return out


def create_ref_mask(m, ids, padding):
    # the reference's TorchScript builder, imported from its training script would pull in the data loader and
    # wandb; its text (train_encoder.py:25-57) is exercised through exec of just those lines instead.
    mod = _load_lines(os.path.join(REF, "training", "train_encoder.py"), 24, 57, "import torch\n", "ref_mask_builder")
    return mod.create_attention_mask(m, ids, padding=padding)


def ref_pad_attn(ids, dtype):
    mod = _load_lines(os.path.join(REF, "evals", "gue.py"), 14, 21, "import torch\nPAD_TOKEN = 1\n", "ref_pad_attn")
    B, T = ids.shape
    m = torch.zeros((B, T, T), dtype=torch.float32)
    return mod.pad_attn(m, ids).to(dtype)


_MODS = {}


def _load_lines(path, start, end, header, modname):
    """Import lines [start, end) of a reference file as a module (TorchScript needs real source on disk)."""
    if modname in _MODS:
        return _MODS[modname]
    import importlib.util
    import tempfile
    src = open(path).read().split("\n")
    tmp = os.path.join(tempfile.mkdtemp(prefix="obt_ref_"), modname + ".py")
    with open(tmp, "w") as f:
        f.write(header + "\n".join(src[start:end]) + "\n")
    spec = importlib.util.spec_from_file_location(modname, tmp)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    _MODS[modname] = mod
    return mod


def orc_to4(m3, n_head):
    return m3.unsqueeze(1).expand(-1, n_head, -1, -1)


def make_case(name, n_layer, n_embd, n_head, vocab, block_size, B, T, dtype, with_grads=True, seed=0):
    rng = np.random.RandomState(seed + 17)
    model = build_reference(n_layer, n_embd, n_head, vocab, block_size, dtype, seed)
    model.eval()
    wm = model.lm_head.width_mult()
    case = {
        "name": name,
        "cfg": dict(n_layer=n_layer, n_embd=n_embd, n_head=n_head, vocab_size=vocab, block_size=block_size),
        "dtype": str(dtype).replace("torch.", ""),
        "width_mult": float(wm),
        "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
    }
    ids = synth_ids(B, T, vocab, rng, padded=False)
    ids_pad = synth_ids(B, T, vocab, rng, padded=True)
    case["ids"], case["ids_pad"] = ids, ids_pad
    masks = ref_masks(ids, n_head, dtype)
    masks_p = ref_masks(ids_pad, n_head, dtype)
    pad_m = ref_pad_attn(ids_pad, dtype)
    case["mask_doc"] = masks["doc"][:, 0].clone()
    case["mask_docpad"] = masks_p["docpad"][:, 0].clone()
    case["mask_pad"] = pad_m.clone()
    with torch.no_grad():
        case["emb_none"] = model(ids, return_embeddings=True)
        case["logits_none"] = model(ids)
        case["emb_doc"] = model(ids, attn_mask=masks["doc"], return_embeddings=True)
        case["logits_doc"] = model(ids, attn_mask=masks["doc"])
        case["emb_docpad"] = model(ids_pad, attn_mask=masks_p["docpad"], return_embeddings=True)
        case["emb_pad"] = model(ids_pad, attn_mask=orc_to4(pad_m, n_head), return_embeddings=True)
        for method in ["mean", "first", "last", "max", "all"]:
            case["encode_" + method] = model.encode(ids, method=method)
        # odd, unaligned length, batch 1, no mask (the evals' call pattern)
        case["ids_odd"] = ids[:1, : T - 3].clone()
        case["emb_odd"] = model(case["ids_odd"], return_embeddings=True)

    if with_grads:
        # one accumulation micro-step of train_encoder.py:273-308 with n_accum = 2
        mrng = np.random.RandomState(seed + 99)
        lm, masked = orc.mlm_mask(ids, mrng)
        case["mlm_mask"], case["ids_masked"] = lm, masked
        model.train()
        model.zero_grad(set_to_none=True)
        logits = model.forward(masked, attn_mask=masks["doc"])
        loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)), ids.view(-1), reduction="none") / 2
        loss *= lm.view(-1).float()
        loss = loss.sum() / lm.view(-1).sum()
        loss.backward()
        case["loss"] = loss.detach().clone()
        case["grads"] = {n: p.grad.clone() for n, p in model.named_parameters()}
        # clip + MuAdamW step (train_encoder.py:195-201,316-317), lr 1e-2 * sqrt(1024)/32, wd 1e-2
        opt = MuAdamW(model.parameters(), lr=1e-2, weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8)
        case["opt_groups"] = [dict(lr=g["lr"], weight_decay=g["weight_decay"], n=len(g["params"])) for g in opt.param_groups]
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        case["grad_norm"] = gn.detach().clone()
        opt.step()
        case["params_after_step"] = {n: p.detach().clone() for n, p in model.named_parameters()}
    return case


def adamw_case():
    """torch.optim.AdamW on bf16 tensors for 3 steps: pins oracle.adamw_step and the fused kernel's rounding."""
    torch.manual_seed(5)
    p = torch.nn.Parameter((torch.randn(4096) * 0.5).to(torch.bfloat16))
    opt = torch.optim.AdamW([p], lr=3e-3, weight_decay=0.1, betas=(0.9, 0.999), eps=1e-8)
    out = {"p0": p.detach().clone(), "g": [], "p": [], "m": [], "v": [], "lr": 3e-3, "wd": 0.1}
    for s in range(3):
        g = (torch.randn(4096) * (0.1 + s)).to(torch.bfloat16)
        p.grad = g.clone()
        opt.step()
        st = opt.state[p]
        out["g"].append(g)
        out["p"].append(p.detach().clone())
        out["m"].append(st["exp_avg"].clone())
        out["v"].append(st["exp_avg_sq"].clone())
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = [
        make_case("bf16_h2", 1, 256, 2, 320, 64, 3, 40, torch.bfloat16),
        make_case("bf16_h1", 2, 128, 1, 256, 64, 2, 48, torch.bfloat16, seed=1),
        make_case("fp32_h1", 1, 128, 1, 256, 64, 2, 24, torch.float32, with_grads=False, seed=2),
    ]
    for c in cases:
        path = os.path.join(OUT, c["name"] + ".pt")
        torch.save(c, path)
        print(path, os.path.getsize(path) // 1024, "KiB")
    path = os.path.join(OUT, "adamw_bf16.pt")
    torch.save(adamw_case(), path)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
