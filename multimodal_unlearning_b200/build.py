"""In-tree build of libclear_b200.so (sm_100a only).

`python -m multimodal_unlearning_b200.build` compiles every `csrc/*.cu` with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` and links one shared library next to this file.
nvcc cross-compiles without a GPU, so this runs in the CPU build container; the resulting .so travels
to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libclear_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(ARCH + CFLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "clear_b200.h"]):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".sha")
    dig = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [NVCC, *ARCH, *CFLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    stamp.write_text(dig)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    if force:
        for f in BUILD.glob("*.sha"):
            f.unlink()
    srcs = _sources()
    if not srcs:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if LIB.exists() and LIB.stat().st_mtime >= newest and not force:
        return LIB
    cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), ]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(lib)
