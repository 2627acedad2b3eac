"""One GEMM configuration per process (so a hang or fault is contained): prints error statistics and, on mismatch,
a coarse map of which 32x32 blocks of the first tiles are wrong (descriptor / layout debugging aid)."""
import sys

import torch

sys.path.insert(0, ".")
from omnibiote_b200 import _lib, ops  # noqa: E402

cg, a_mn, b_mn, M, N, K = [int(x) for x in sys.argv[1:7]]
lib = _lib.load()
lib.obt_gemm_set_cta_group(cg)
g = torch.Generator(device="cuda").manual_seed(0)
a = (torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda") * 0.5).to(torch.bfloat16)
b = (torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda") * 0.5).to(torch.bfloat16)
ref = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t())
out = ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), allow_splitk=False)
torch.cuda.synchronize()
err = (out.float() - ref).abs()
tol = ref.abs() * 2 ** -7 + 1e-2
bad = err > tol
print(f"cg={cg} a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}: bad={int(bad.sum())}/{bad.numel()} maxerr={float(err.max()):.4g} "
      f"rel={float((out.float()-ref).norm()/ref.norm()):.3e}")
if bad.any():
    mb, nb = min(M, 256) // 32, min(N, 256) // 32
    for i in range(mb):
        print(" ".join(f"{float(bad[i*32:(i+1)*32, j*32:(j+1)*32].float().mean()):.2f}" for j in range(nb)))
    print("out[0,:8]", out[0, :8].float().tolist())
    print("ref[0,:8]", ref[0, :8].tolist())
    sys.exit(1)
if len(sys.argv) > 7:  # timing
    for _ in range(3):
        ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), allow_splitk=False)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), allow_splitk=False)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"  time {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
    a32 = a.float().t().contiguous().to(torch.bfloat16) if a_mn else a
    b32 = b.float().t().contiguous().to(torch.bfloat16) if b_mn else b
    for _ in range(3):
        torch.matmul(a32, b32.t())
    s.record()
    for _ in range(10):
        torch.matmul(a32, b32.t())
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"  cuBLAS (torch.matmul) {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
