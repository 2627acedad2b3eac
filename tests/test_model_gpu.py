"""GPU: the drop-in OmniBioTA module on the B200 kernels against golden outputs of the unmodified reference
(tests/golden, bf16 CPU run) and against the CPU oracle on fresh seeded inputs.

Tolerances (bf16, stated per SURVEY Appendix C.2): relative Frobenius error <= 1.5e-2 per tensor
(<= 5e-2 for wte.weight.grad), loss within 1 bf16 ulp; mask / pooling index selection bit-exact."""
import copy

import pytest
import torch

from conftest import rel_err, max_abs

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
TOL = 1.5e-2


def build(case):
    from omnibiote_b200.model import OmniBioTA, OmniBioTAConfig
    from omnibiote_b200.mup import set_base_shapes
    cfg = OmniBioTAConfig()
    for k, v in case["cfg"].items():
        setattr(cfg, k, v)
    cfg.dropout = 0.0
    cfg.flash = True
    m = OmniBioTA(cfg)
    c2 = copy.copy(cfg); c2.n_embd, c2.n_head = 24, 3
    base = OmniBioTA(c2)
    c3 = copy.copy(cfg); c3.n_embd, c3.n_head = 48, 12
    delta = OmniBioTA(c3)
    set_base_shapes(m, base, delta=delta)
    assert abs(m.lm_head.width_mult() - case["width_mult"]) < 1e-9
    m.to(BF)
    missing = m.load_state_dict(case["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.cuda()


def m4(mask3, H):
    return mask3.cuda().unsqueeze(1).expand(-1, H, -1, -1)


@pytest.mark.parametrize("name", ["bf16_h2", "bf16_h1"])
def test_forward_matches_reference_golden(golden, name):
    c = golden(name)
    H = c["cfg"]["n_head"]
    model = build(c).eval()
    assert list(model.state_dict().keys()) == list(c["state_dict"].keys())
    with torch.no_grad():
        ids, ids_pad = c["ids"].cuda(), c["ids_pad"].cuda()
        report = {
            "emb_none": rel_err(model(ids, return_embeddings=True), c["emb_none"]),
            "logits_none": rel_err(model(ids), c["logits_none"]),
            "emb_doc": rel_err(model(ids, attn_mask=m4(c["mask_doc"], H), return_embeddings=True), c["emb_doc"]),
            "logits_doc": rel_err(model(ids, attn_mask=m4(c["mask_doc"], H)), c["logits_doc"]),
            "emb_docpad": rel_err(model(ids_pad, attn_mask=m4(c["mask_docpad"], H), return_embeddings=True), c["emb_docpad"]),
            "emb_pad": rel_err(model(ids_pad, attn_mask=m4(c["mask_pad"], H), return_embeddings=True), c["emb_pad"]),
            "emb_odd": rel_err(model(c["ids_odd"].cuda(), return_embeddings=True), c["emb_odd"]),
        }
    print(name, {k: f"{v:.2e}" for k, v in report.items()})
    for k, v in report.items():
        assert v < TOL, (k, v)


@pytest.mark.parametrize("name", ["bf16_h2", "bf16_h1"])
def test_encode_all_methods(golden, name):
    c = golden(name)
    model = build(c).eval()
    ids = c["ids"].cuda()
    with torch.no_grad():
        emb = model.encode(ids, "all")
        assert rel_err(emb, c["encode_all"]) < TOL
        # index selection is bit-exact with respect to our own embeddings
        assert torch.equal(model.encode(ids, "first"), emb[:, 0])
        assert torch.equal(model.encode(ids, "last"), emb[:, -1])
        assert torch.equal(model.encode(ids, "max"), emb.max(dim=1)[0])
        assert max_abs(model.encode(ids, "mean"), emb.float().mean(dim=1)) <= 2 ** -8 * float(emb.abs().max())
        for method in ["mean", "first", "last", "max"]:
            assert rel_err(model.encode(ids, method), c["encode_" + method]) < TOL, method
    with pytest.raises(AssertionError):
        model.encode(ids, "median")
    with pytest.raises(AssertionError):
        model(torch.zeros(1, c["cfg"]["block_size"] + 1, dtype=torch.long, device="cuda"))


@pytest.mark.parametrize("name", ["bf16_h2", "bf16_h1"])
@pytest.mark.parametrize("path", ["dropin", "fused", "masked_rows"])
def test_mlm_loss_and_gradients(golden, name, path):
    """train_encoder.py:296-308 on the same weights / batch / MLM mask as the golden reference run (n_accum = 2)."""
    c = golden(name)
    H = c["cfg"]["n_head"]
    model = build(c).train()
    ids, masked, lm = c["ids"].cuda(), c["ids_masked"].cuda(), c["mlm_mask"].cuda()
    mask = m4(c["mask_doc"], H)
    if path == "dropin":
        logits = model.forward(masked, attn_mask=mask)
        loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)), ids.view(-1), reduction="none") / 2
        loss *= lm.view(-1).float()
        loss = loss.sum() / lm.view(-1).sum()
    elif path == "fused":
        loss, _ = model.mlm_loss(masked, ids, lm, attn_mask=mask, n_accum=2)
    else:  # head restricted to the rows inside the MLM mask: same loss, same gradients
        cap = -(-int(lm.sum()) // 8) * 8 + 8
        loss, _ = model.mlm_loss(masked, ids, lm, attn_mask=mask, n_accum=2, masked_rows_cap=cap)
        assert model.head_rows_meta.tolist() == [int(lm.sum()), 0]
    loss.backward()
    ulp = 2 ** -7 * float(c["loss"])
    assert abs(float(loss) - float(c["loss"])) <= ulp, (float(loss), float(c["loss"]))
    report = {}
    for n, p in model.named_parameters():
        assert p.grad is not None, n
        report[n] = rel_err(p.grad, c["grads"][n])
    print(name, path, {k: f"{v:.2e}" for k, v in report.items()})
    for n, v in report.items():
        assert v < (5e-2 if "wte" in n else 2.5e-2), (n, v)


def test_clip_and_muadamw_step_matches_reference(golden):
    """clip_grad_norm_(1.0) + MuAdamW step (train_encoder.py:195-201,316-317) from the reference's own gradients."""
    from omnibiote_b200.optim import MuAdamW
    c = golden("bf16_h2")
    model = build(c)
    for n, p in model.named_parameters():
        p.grad = c["grads"][n].clone().cuda()
    opt = MuAdamW(model.parameters(), lr=1e-2, weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8)
    got = [(round(g["lr"], 10), round(g["weight_decay"], 10), len(g["params"])) for g in opt.param_groups]
    want = [(round(g["lr"], 10), round(g["weight_decay"], 10), g["n"]) for g in c["opt_groups"]]
    assert got == want
    norm = opt.clip_and_step(max_norm=1.0)
    assert abs(float(norm[0]) - float(c["grad_norm"])) <= 0.02 * float(c["grad_norm"])
    report = {}
    for n, p in model.named_parameters():
        # first step moves every weight by ~lr: compare the update, not the weight
        upd = p.detach().float().cpu() - c["state_dict"][n].float()
        ref = c["params_after_step"][n].float() - c["state_dict"][n].float()
        report[n] = rel_err(upd, ref)
    print("update rel err", {k: f"{v:.3f}" for k, v in report.items()})
    for n, p in model.named_parameters():
        assert report[n] < 0.02, (n, report[n])  # measured on B200: <= 0.004 (profiles/r02s_tolerances.log)
        assert max_abs(p, c["params_after_step"][n]) <= 2 ** -6 * float(c["params_after_step"][n].abs().max()), n


def test_module_behaviours_used_by_callers(golden):
    """deepcopy / state_dict round trip / .train() / .eval() / named_parameters filter (gue.py:51,60-64)."""
    c = golden("bf16_h1")
    model = build(c)
    clone = copy.deepcopy(model)
    ids = c["ids"].cuda()
    with torch.no_grad():
        assert torch.equal(model.eval()(ids, return_embeddings=True), clone.eval()(ids, return_embeddings=True))
    assert model.transformer.wte.weight.shape[-1] == c["cfg"]["n_embd"]
    assert model.transformer.h[0].attn.n_head == c["cfg"]["n_head"]
    assert any("wte" in n for n, _ in model.named_parameters())
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    for k, v in c["state_dict"].items():
        assert torch.equal(sd[k], v), k
    assert model.get_num_params() == sum(v.numel() for k, v in c["state_dict"].items()
                                         if "freqs" not in k and "wte" not in k)
    with pytest.raises(RuntimeError):
        clone.cpu()(c["ids"], return_embeddings=True)  # no CPU fallback


def test_activation_checkpointing_same_gradients(golden):
    c = golden("bf16_h1")
    H = c["cfg"]["n_head"]
    ids, lm = c["ids"].cuda(), c["mlm_mask"].cuda()
    grads = []
    for freq in (0, 1):
        model = build(c).train()
        model.config.checkpoint_freq = freq
        loss, _ = model.mlm_loss(c["ids_masked"].cuda(), ids, lm, attn_mask=m4(c["mask_doc"], H), n_accum=2)
        loss.backward()
        grads.append({n: p.grad.clone() for n, p in model.named_parameters()})
    for n in grads[0]:
        assert torch.equal(grads[0][n], grads[1][n]), n


def test_reference_checkpoint_pickle_runs_on_the_kernels():
    """A whole-module pickle of the unmodified reference (tests/golden/ref_checkpoint.pt) is read without the
    reference's code, moved to the GPU and must reproduce the oracle's logits for the same weights."""
    import io
    import os
    from conftest import GOLDEN
    import omnibiota_oracle as orc
    from omnibiote_b200 import checkpoint
    c = torch.load(os.path.join(GOLDEN, "ref_checkpoint.pt"), map_location="cpu", weights_only=False)["bf16"]
    model = checkpoint.load_reference_checkpoint(io.BytesIO(c["pickle"])).cuda().eval()
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(4, 96, (3, 24), generator=g)
    got = model(ids.cuda()).float().cpu()
    p = {k: v for k, v in c["state_dict"].items()}
    want = orc.forward(p, 2, 4, ids, None, readout_width_mult=c["width_mult"]).float()
    assert rel_err(got, want) < 1.5e-2, rel_err(got, want)
