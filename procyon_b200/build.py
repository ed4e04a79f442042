"""Builds libprocyon_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

The shared library is the product's only compute path: there is no CPU or PyTorch fallback. nvcc
cross-compiles for sm_100a without a GPU, so this runs in the CPU-only container and the resulting .so
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD = PKG_DIR / "build"
LIB_PATH = PKG_DIR / "libprocyon_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
]
# --use_fast_math would turn erff/expf/division approximate in places where parity matters:
# kernels that need exact functions call the precise intrinsics explicitly, so keep the default math
# and only enable fast FMA contraction (nvcc default). Remove the flag globally:
COMMON_FLAGS.remove("--use_fast_math")


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "procyon_b200.h"]:
        h.update(hdr.read_bytes())
    h.update(" ".join(ARCH_FLAGS + COMMON_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: Path, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".sha")
    dig = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [NVCC, *ARCH_FLAGS, *COMMON_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = BUILD / (src.stem + ".ptxas.log")
    log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        for line in (res.stdout + res.stderr).splitlines():
            if "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
                print(f"[{src.name}] {line.strip()}")
    stamp.write_text(dig)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    if force:
        for f in BUILD.glob("*.sha"):
            f.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    link_stamp = BUILD / "link.sha"
    dig = hashlib.sha256(b"".join(o.read_bytes() for o in objs)).hexdigest()
    if LIB_PATH.exists() and link_stamp.exists() and link_stamp.read_text() == dig:
        return LIB_PATH
    cmd = [NVCC, *ARCH_FLAGS, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-cudart", "static"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    link_stamp.write_text(dig)
    return LIB_PATH


if __name__ == "__main__":
    p = build(verbose=True, force="--force" in sys.argv)
    print(p)
