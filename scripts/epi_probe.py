"""Times every GEMM epilogue on one shape (CUDA events, 10 iterations)."""
import sys

import torch

sys.path.insert(0, ".")
from omnibiote_b200 import ops  # noqa: E402

M, N, K = [int(x) for x in sys.argv[1:4]]
only = sys.argv[4] if len(sys.argv) > 4 else None
a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
b = (torch.randn(N, K, device="cuda") * 0.5).to(torch.bfloat16)
aux = torch.randn(M, N, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
u = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
cases = {
    "plain": dict(epilogue=ops.EPI_PLAIN),
    "resid": dict(epilogue=ops.EPI_RESID, aux_in=aux),
    "gelu": dict(epilogue=ops.EPI_GELU, aux_out=u),
    "gelu_bwd": dict(epilogue=ops.EPI_GELU_BWD, aux_in=aux),
    "resid_dropout": dict(epilogue=ops.EPI_RESID_DROPOUT, aux_in=aux, drop_p=0.1, seed=1, offset=0),
    "mul": dict(epilogue=ops.EPI_MUL, aux_in=aux),
    "gelu_dg": dict(epilogue=ops.EPI_GELU_DG, aux_out=u),
}
if N % 384 == 0 and M % 1024 == 0:  # c_attn: rotary (cosine-scaling form of a bf16 model) on the q | k two thirds
    cos = torch.rand(1024, 64, device="cuda")
    cases["rope"] = dict(epilogue=ops.EPI_ROPE, rope=(cos, None, 1024, 128, 2 * N // 3))
if N % 128 == 0 and M % 1024 == 0:
    cases["delta"] = dict(epilogue=ops.EPI_DELTA, aux_in=aux,
                          delta=(torch.empty(M // 1024, N // 128, 1024, device="cuda"), 1024))
for name, kw in cases.items():
    if only and name != only:
        continue
    for _ in range(3):
        ops.gemm(a, b, out=out, allow_splitk=False, **kw)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.gemm(a, b, out=out, allow_splitk=False, **kw)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"{name:14s} {M}x{N}x{K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s")
