"""GPU: tcgen05 GEMM (obt_gemm_bf16 through the C ABI) against an fp32 torch matmul of the same bf16 operands."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ops():
    from omnibiote_b200 import _lib, ops
    return _lib.load(), ops


def _mk(M, N, K, a_mn, b_mn, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    af = a.float().t() if a_mn else a.float()
    bf = b.float().t() if b_mn else b.float()
    return a, b, af @ bf.t()


def _check(out, ref, what, ulps=1, mag=None):
    out_f = out.float()
    # one bf16 rounding of an fp32-accumulated result: |err| <= 2^-8 |ref| (+ accumulation-order noise).
    # mag: magnitude of the intermediate that was rounded (a residual sum can cancel: the rounding error of the
    # bf16 accumulator is relative to |acc|, not to |acc + resid|)
    tol = (ref.abs() if mag is None else torch.maximum(ref.abs(), mag)) * 2 ** -7 * ulps + 1e-2
    bad = (out_f - ref).abs() > tol
    if bad.any():
        idx = bad.nonzero()
        rows = torch.unique(idx[:, 0] // 32)[:16].tolist()
        cols = torch.unique(idx[:, 1] // 32)[:16].tolist()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements off; rel={rel_err(out_f, ref):.3e}; "
                             f"bad row-blocks(32) {rows} col-blocks(32) {cols}; first {idx[0].tolist()} "
                             f"got {out_f[tuple(idx[0])].item()} want {ref[tuple(idx[0])].item()}")
    assert rel_err(out_f, ref) < 4e-3, what


SHAPES = [(128, 256, 64), (256, 512, 128), (384, 256, 512), (300, 520, 200), (1024, 1024, 1024), (6, 1024, 1024),
          (2048, 768, 256)]
LAYOUTS = [(False, False), (False, True), (True, True), (True, False)]


@pytest.mark.parametrize("cg", [1, 2, 3])  # 1 CTA / cta_group::2 pair / 2-CTA cluster with multicast B
@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_plain(cg, a_mn, b_mn, M, N, K):
    lib, ops = _ops()
    if a_mn and M % 8:
        pytest.skip("MN-major A needs M % 8 == 0 (TMA pitch)")
    if b_mn and N % 8:
        pytest.skip("MN-major B needs N % 8 == 0 (TMA pitch)")
    lib.obt_gemm_set_cta_group(cg)
    try:
        a, b, ref = _mk(M, N, K, a_mn, b_mn)
        out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, allow_splitk=False)
        torch.cuda.synchronize()
        _check(out, ref, f"cg{cg} a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}")
    finally:
        lib.obt_gemm_set_cta_group(0)


@pytest.mark.parametrize("cg", [1, 2, 3])  # 1 CTA / cta_group::2 pair / 2-CTA cluster with multicast B
def test_gemm_strided_operands_and_output(cg):
    lib, ops = _ops()
    lib.obt_gemm_set_cta_group(cg)
    try:
        M, N, K = 512, 256, 384
        big_a = (torch.randn(M, 3 * K, device="cuda") * 0.5).to(torch.bfloat16)
        a = big_a[:, K:2 * K]  # column slice: pitch 3K
        b = (torch.randn(N, K, device="cuda") * 0.5).to(torch.bfloat16)
        big_o = torch.zeros(M, 2 * N, dtype=torch.bfloat16, device="cuda")
        ops.gemm(a, b, out=big_o[:, N:], allow_splitk=False)
        torch.cuda.synchronize()
        _check(big_o[:, N:], a.float() @ b.float().t(), "strided")
        assert float(big_o[:, :N].abs().max()) == 0.0
    finally:
        lib.obt_gemm_set_cta_group(0)


@pytest.mark.parametrize("cg", [1, 2, 3])  # 1 CTA / cta_group::2 pair / 2-CTA cluster with multicast B
def test_gemm_epilogues(cg):
    lib, ops = _ops()
    lib.obt_gemm_set_cta_group(cg)
    try:
        M, N, K = 640, 512, 256
        a, b, ref = _mk(M, N, K, False, False, seed=3)
        res = (torch.randn(M, N, device="cuda")).to(torch.bfloat16)
        # residual
        out = ops.gemm(a, b, epilogue=ops.EPI_RESID, aux_in=res, allow_splitk=False)
        want = (res.float() + ref.to(torch.bfloat16).float())
        _check(out, want, "resid", ulps=2, mag=ref.abs())
        # in-place accumulate (D aliases aux_in)
        acc = res.clone()
        ops.gemm(a, b, out=acc, epilogue=ops.EPI_RESID, aux_in=acc, allow_splitk=False)
        _check(acc, want, "accumulate in place", ulps=2, mag=ref.abs())
        # gelu (reference expression, constant 1.41421)
        u = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        g = ops.gemm(a, b, epilogue=ops.EPI_GELU, aux_out=u, allow_splitk=False)
        _check(u, ref, "gelu pre-activation")
        uf = u.float()
        _check(g, uf * 0.5 * (1.0 + torch.erf(uf / 1.41421)), "gelu")
        # gelu backward
        x = uf.clone().requires_grad_(True)
        (x * 0.5 * (1.0 + torch.erf(x / 1.41421))).sum().backward()
        d = ops.gemm(a, b, epilogue=ops.EPI_GELU_BWD, aux_in=u, allow_splitk=False)
        _check(d, ref.to(torch.bfloat16).float() * x.grad, "gelu bwd")
        # the pair the block uses: forward saves gelu'(U), backward multiplies by it
        dg = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        g2 = ops.gemm(a, b, epilogue=ops.EPI_GELU_DG, aux_out=dg, allow_splitk=False)
        assert torch.equal(g2, g)
        _check(dg, x.grad, "gelu derivative")
        d2 = ops.gemm(a, b, epilogue=ops.EPI_MUL, aux_in=dg, allow_splitk=False)
        _check(d2, ref.to(torch.bfloat16).float() * dg.float(), "mul epilogue")
        _check(d2, ref.to(torch.bfloat16).float() * x.grad, "gelu bwd via saved derivative", ulps=2)
    finally:
        lib.obt_gemm_set_cta_group(0)


@pytest.mark.parametrize("cg", [1, 2, 3])
def test_gemm_rowmask_epilogue(cg):
    """Epilogue 8 (MLM head): rows outside the uint8 mask are stored as exact zeros, the others as rb(acc)."""
    lib, ops = _ops()
    lib.obt_gemm_set_cta_group(cg)
    try:
        M, N, K = 700, 520, 256
        a, b, ref = _mk(M, N, K, False, False, seed=11)
        g = torch.Generator(device="cuda").manual_seed(1)
        mask = (torch.rand(M, generator=g, device="cuda") < 0.15).to(torch.uint8)
        out = ops.gemm(a, b, epilogue=ops.EPI_ROWMASK, aux_in=mask, allow_splitk=False)
        plain = ops.gemm(a, b, allow_splitk=False)
        assert torch.equal(out[mask.bool()], plain[mask.bool()])
        assert float(out[~mask.bool()].abs().max()) == 0.0
    finally:
        lib.obt_gemm_set_cta_group(0)


@pytest.mark.parametrize("cg", [1, 2, 3])  # 1 CTA / cta_group::2 pair / 2-CTA cluster with multicast B
def test_gemm_splitk_wgrad_shape(cg):
    lib, ops = _ops()
    lib.obt_gemm_set_cta_group(cg)
    try:
        # dW[N',K'] = dY[Mtok,N']^T X[Mtok,K']: long reduction, small output -> split-K + accumulate into grad
        Mtok, Nn, Kk = 8192, 512, 256
        dy = (torch.randn(Mtok, Nn, device="cuda") * 0.1).to(torch.bfloat16)
        x = (torch.randn(Mtok, Kk, device="cuda") * 0.5).to(torch.bfloat16)
        ref = dy.float().t() @ x.float()
        out = ops.gemm(dy, x, a_mn=True, b_mn=True)
        _check(out, ref, "split-K plain")
        grad = (torch.randn(Nn, Kk, device="cuda")).to(torch.bfloat16)
        want = grad.float() + ref.to(torch.bfloat16).float()
        ops.gemm(dy, x, out=grad, a_mn=True, b_mn=True, epilogue=ops.EPI_RESID, aux_in=grad)
        _check(grad, want, "split-K accumulate", ulps=2, mag=ref.abs())  # two roundings: rb(acc) then rb(old + .)
    finally:
        lib.obt_gemm_set_cta_group(0)


@pytest.mark.parametrize("cg", [1, 2, 3])
@pytest.mark.parametrize("B,T,H", [(2, 200, 2), (3, 256, 3), (4, 1024, 8)])
def test_gemm_delta_epilogue(cg, B, T, H):
    """Epilogue 11: D = rb(acc) and delta[b, h, t] = sum over head h's 128 columns of D * aux_in (the attention
    backward's rowsum(dO * O), emitted by the GEMM that produces dO)."""
    lib, ops = _ops()
    lib.obt_gemm_set_cta_group(cg)
    try:
        M, N, K = B * T, H * 128, 256
        a, b, ref = _mk(M, N, K, False, True, seed=5)
        y = torch.randn(M, N, device="cuda").to(torch.bfloat16)
        delta = torch.full((B, H, T), float("nan"), device="cuda")
        out = ops.gemm(a, b, b_mn=True, epilogue=ops.EPI_DELTA, aux_in=y, delta=(delta, T))
        assert torch.equal(out, ops.gemm(a, b, b_mn=True, allow_splitk=False))          # D is the plain result
        want = (out.float() * y.float()).view(B, T, H, 128).sum(-1).permute(0, 2, 1)    # [B, H, T]
        assert rel_err(delta, want) < 1e-5 and bool(torch.isfinite(delta).all())
    finally:
        lib.obt_gemm_set_cta_group(0)


def test_gemm_rejects_bad_arguments():
    lib, ops = _ops()
    a = torch.zeros(16, 12, dtype=torch.bfloat16, device="cuda")  # K=12: pitch not a multiple of 8
    b = torch.zeros(16, 12, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        ops.gemm(a, b)
    with pytest.raises(RuntimeError):
        ops.gemm(a.cpu(), b.cpu())
