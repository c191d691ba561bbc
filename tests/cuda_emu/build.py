"""Compile a product .cu file for the CPU emulator (tests/cuda_emu/emu.h): the source is used as it is, only the
`kernel<<<grid, block, smem, stream>>>(args);` statements become EMU_LAUNCH(kernel, grid, block, args);.  The result is
a shared library exporting the file's `pgs::launch_*` functions through the extern "C" wrappers given in `exports`.
TEST INFRASTRUCTURE (only tests/ use it)."""
from __future__ import annotations

import hashlib
import re
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "partgs_b200" / "csrc"
OUT = HERE / "_build"
CUDA_INC = "/usr/local/cuda/include"

_LAUNCH = re.compile(r"(\w+(?:<[^<>;]*>)?)\s*<<<\s*(.*?)\s*>>>\s*\((.*?)\)\s*;", re.S)


def _split_args(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src: str) -> str:
    def sub(m):
        cfg = _split_args(m.group(2))
        return f"EMU_LAUNCH({m.group(1)}, {cfg[0]}, {cfg[1]}, {m.group(3)});"
    return _LAUNCH.sub(sub, src)


def build(cu_name: str, exports: str) -> Path:
    """-> path of the emulator build of partgs_b200/csrc/<cu_name>; `exports` is C++ text appended to the translation
    unit (extern "C" wrappers)."""
    src = (CSRC / cu_name).read_text()
    body = rewrite_launches(src)
    assert "<<<" not in body, "unconverted kernel launch"
    tu = ('#include "emu.h"\nnamespace pgs { void count_launch(int) {} }\n' + body + "\n" + exports + "\n")
    deps = "".join(p.read_text() for p in [HERE / "emu.h", CSRC / "common.cuh", CSRC / "kernels.h"])
    tag = hashlib.sha1((tu + deps).encode()).hexdigest()[:16]
    OUT.mkdir(exist_ok=True)
    so = OUT / f"{Path(cu_name).stem}_{tag}.so"
    if so.exists():
        return so
    cpp = OUT / f"{Path(cu_name).stem}_{tag}.cpp"
    cpp.write_text(tu)
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-w", "-ffp-contract=off", f"-I{HERE}",
           f"-I{CSRC}", f"-I{CUDA_INC}", str(cpp), "-o", str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + res.stderr[-4000:])
    return so
