"""In-tree build of the C-ABI CUDA library (``libomnibiote_b200.so``) for sm_100a.

nvcc cross-compiles without a GPU; the resulting ``.so`` is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libomnibiote_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build omnibiote_b200 CUDA library")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    stamp = BUILD_DIR / "digest.txt"
    return not (LIB_PATH.exists() and stamp.exists() and stamp.read_text() == _digest())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ and link the shared library. Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    BUILD_DIR.mkdir(exist_ok=True)
    # one builder at a time: under torchrun every rank gets here on a fresh checkout; the others wait on the lock and
    # then find the finished library (the link goes to a temporary name and is renamed into place, so a concurrent
    # dlopen never sees a half-written file)
    with open(BUILD_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB_PATH
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    nvcc = _nvcc()
    headers_mtime = max(p.stat().st_mtime for p in CSRC.glob("*.cuh"))

    def compile_one(src: Path) -> Path:
        obj = BUILD_DIR / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, headers_mtime):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp_lib = BUILD_DIR / f"libomnibiote_b200.{os.getpid()}.so.tmp"
    cmd = [nvcc, "-shared", "-o", str(tmp_lib), *map(str, objs), "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp_lib, LIB_PATH)
    (BUILD_DIR / "digest.txt").write_text(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
