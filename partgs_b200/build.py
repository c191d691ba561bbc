"""In-tree build of libpartgs_b200.so (hand-written CUDA for sm_100a behind a C ABI).

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so that it
travels with the repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
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
OBJ_DIR = PKG_DIR / "csrc" / "build"
LIB_PATH = PKG_DIR / "libpartgs_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# Same fp defaults as the reference build (fmad on, no fast-math): bit-exact radii /
# keys depend on identical FMA contraction.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    *os.environ.get("PGS_NVCC_EXTRA", "").split(),   # experiments only (e.g. -DPGS_BWD_MIN_CTAS=5)
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _stamp(src: Path) -> str:
    h = hashlib.sha1()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG_DIR.parent / "include" / "partgs_b200.h"]):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ_DIR / (src.stem + ".o")
    stamp_file = OBJ_DIR / (src.stem + ".stamp")
    stamp = _stamp(src)
    if obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj
    cmd = [NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = OBJ_DIR / (src.stem + ".ptxas.log")
    log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        for line in (res.stdout + res.stderr).splitlines():
            if "error" in line.lower() or "spill" in line.lower() and "0 bytes spill" not in line:
                print(f"[{src.name}] {line}")
    stamp_file.write_text(stamp)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    if force:
        for f in OBJ_DIR.glob("*.stamp"):
            f.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if (not LIB_PATH.exists()) or LIB_PATH.stat().st_mtime < newest or force:
        cmd = [NVCC, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    p = build(verbose=True, force="--force" in sys.argv)
    print(p)
