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



class EmuUnavailable(RuntimeError):
    """The host cannot build the emulator at all (no g++ / no CUDA toolkit headers): tests skip, they do not fail."""


def _require_toolchain():
    import shutil
    if shutil.which("g++") is None:
        raise EmuUnavailable("g++ not found")
    if not (Path(CUDA_INC) / "cuda_runtime.h").exists():
        raise EmuUnavailable(f"{CUDA_INC}/cuda_runtime.h not found")


_DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];")
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
    src = src.replace("<< <", "<<<").replace(">> >", ">>>")   # the reference spells its launches with spaces
    def sub(m):
        cfg = _split_args(m.group(2))
        smem = cfg[2] if len(cfg) > 2 else "0"
        return f"EMU_LAUNCH(({m.group(1)}), {cfg[0]}, {cfg[1]}, {smem}, {m.group(3)});"
    src = _LAUNCH.sub(sub, src)
    # extern __shared__ [__align__(n)] T name[];  ->  T* name = (T*)emu::g_dyn_smem;
    return _DYN_SMEM.sub(lambda m: f"{m.group(1)}* {m.group(2)} = ({m.group(1)}*)emu::g_dyn_smem;", src)


def build(cu_name: str, exports: str) -> Path:
    """-> path of the emulator build of partgs_b200/csrc/<cu_name>; `exports` is C++ text appended to the translation
    unit (extern "C" wrappers)."""
    _require_toolchain()
    src = (CSRC / cu_name).read_text()
    body = rewrite_launches(src)
    assert "<<<" not in body, "unconverted kernel launch"
    tu = ('#include "emu.h"\nnamespace pgs { void count_launch(int) {} }\n' + body + "\n" + exports + "\n")
    deps = "".join(p.read_text() for p in [HERE / "emu.h", CSRC / "common.cuh", CSRC / "kernels.h"])
    suffix, san = _sanitize_flags()
    tag = hashlib.sha1((tu + deps).encode()).hexdigest()[:16] + suffix
    OUT.mkdir(exist_ok=True)
    so = OUT / f"{Path(cu_name).stem}_{tag}.so"
    if so.exists():
        return so
    cpp = OUT / f"{Path(cu_name).stem}_{tag}.cpp"
    cpp.write_text(tu)
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-Wl,-Bsymbolic", "-w", "-ffp-contract=off", *san, f"-I{HERE}",
           f"-I{CSRC}", f"-I{CUDA_INC}", str(cpp), "-o", str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + res.stderr[-4000:])
    return so


def _sanitize_flags():
    """PGS_EMU_SANITIZE=address|thread builds the emulated library with that sanitizer (run pytest with the matching
    runtime preloaded, see tools/emu_sanitize.sh): memcheck / racecheck of the kernels on the CPU."""
    import os
    kind = os.environ.get("PGS_EMU_SANITIZE", "")
    if kind not in ("address", "thread"):
        return "", []
    return "_" + kind, [f"-fsanitize={kind}", "-fno-omit-frame-pointer", "-g"]


def build_full() -> Path:
    """The WHOLE product library (every partgs_b200/csrc/*.cu, including the C-ABI layer api.cu) for the emulator:
    same exported `pgs_*` symbols as libpartgs_b200.so, host memory instead of device memory, synchronous streams."""
    from concurrent.futures import ThreadPoolExecutor
    _require_toolchain()
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "partgs_b200.h",
                                                                      HERE / "emu.h", HERE / "emu_runtime.cpp"])
    h = hashlib.sha1()
    for f in srcs + hdrs:
        h.update(f.read_bytes())
    suffix, san = _sanitize_flags()
    tag = h.hexdigest()[:16] + suffix
    OUT.mkdir(exist_ok=True)
    so = OUT / f"libpartgs_b200_emu_{tag}.so"
    if so.exists():
        return so
    work = OUT / f"full_{tag}"
    work.mkdir(exist_ok=True)
    flags = ["-std=c++17", "-O1", "-fPIC", "-pthread", "-w", "-ffp-contract=off", "-DPGS_EMU", f"-I{HERE}", f"-I{CSRC}",
             f"-I{CUDA_INC}", "-include", "emu.h", *san]

    def compile_one(src: Path) -> Path:
        body = rewrite_launches(src.read_text())
        assert "<<<" not in body, f"unconverted kernel launch in {src.name}"
        body = body.replace('"../../include/partgs_b200.h"', f'"{ROOT / "include" / "partgs_b200.h"}"')
        cpp = work / (src.stem + ".cpp")
        cpp.write_text(body)
        obj = work / (src.stem + ".o")
        res = subprocess.run(["g++", *flags, "-c", str(cpp), "-o", str(obj)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"emulator build of {src.name} failed:\n" + res.stderr[-3000:])
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, srcs))
    rt = work / "emu_runtime.o"
    res = subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-w", *san, f"-I{CUDA_INC}", "-c",
                          str(HERE / "emu_runtime.cpp"), "-o", str(rt)], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulator runtime build failed:\n" + res.stderr[-3000:])
    res = subprocess.run(["g++", "-shared", "-pthread", "-Wl,-Bsymbolic", *san, "-o", str(so), *map(str, objs), str(rt)],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulator link failed:\n" + res.stderr[-3000:])
    return so


def build_reference(fork: str = "part") -> Path:
    """The UNMODIFIED reference rasteriser (cuda_rasterizer/{forward,backward,rasterizer_impl}.cu of the chosen fork,
    read from /root/reference where they lie) compiled for the emulator, with host stand-ins for CUB, cooperative
    groups (ref_shim/) and GLM (oracle/glm_shim), plus ref_wrap_<fork>.cpp: extern "C" entry points over
    CudaRasterizer::Rasterizer::forward / backward.  Build container only; used by tools/make_golden_ref_emu.py."""
    from concurrent.futures import ThreadPoolExecutor
    _require_toolchain()
    if fork == "knn":
        ref = Path("/root/reference/submodules/simple-knn")
        names = ("simple_knn.cu",)
    else:
        sub = {"part": "diff-surfel-rasterization_part", "base": "diff-surfel-rasterization"}[fork]
        ref = Path("/root/reference/submodules") / sub / "cuda_rasterizer"
        names = ("forward.cu", "backward.cu", "rasterizer_impl.cu")
    if not ref.is_dir():
        raise EmuUnavailable(f"{ref} not found (reference tree not mounted)")
    srcs = [ref / n for n in names]
    wrap = HERE / f"ref_wrap_{fork}.cpp"
    h = hashlib.sha1()
    for f in srcs + sorted(ref.glob("*.h")) + [wrap, HERE / "emu.h", HERE / "emu_runtime.cpp"] + \
            sorted((HERE / "ref_shim").rglob("*.*")):
        h.update(f.read_bytes())
    tag = h.hexdigest()[:16]
    OUT.mkdir(exist_ok=True)
    so = OUT / f"libref_{fork}_emu_{tag}.so"
    if so.exists():
        return so
    work = OUT / f"ref_{fork}_{tag}"
    work.mkdir(exist_ok=True)
    flags = ["-std=c++17", "-O1", "-fPIC", "-pthread", "-w", "-ffp-contract=off", f"-I{HERE}", f"-I{HERE / 'ref_shim'}",
             f"-I{ROOT / 'oracle' / 'glm_shim'}", f"-I{ref}", f"-I{CUDA_INC}", "-include", "emu.h", "-include", "cstdint",
             "-include", "cfloat"]

    def compile_one(src: Path) -> Path:
        body = rewrite_launches(src.read_text()) if src.suffix == ".cu" else src.read_text()
        assert "<<<" not in body, f"unconverted kernel launch in {src.name}"
        cpp = work / (src.stem + ".cpp")
        cpp.write_text(body)
        obj = work / (src.stem + ".o")
        res = subprocess.run(["g++", *flags, "-c", str(cpp), "-o", str(obj)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"emulator build of reference {src.name} failed:\n" + res.stderr[-3000:])
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, srcs + [wrap]))
    rt = work / "emu_runtime.o"
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-w", f"-I{CUDA_INC}", "-c", str(HERE / "emu_runtime.cpp"), "-o",
                    str(rt)], check=True)
    res = subprocess.run(["g++", "-shared", "-pthread", "-Wl,-Bsymbolic", "-o", str(so), *map(str, objs), str(rt)],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulator link of the reference failed:\n" + res.stderr[-3000:])
    return so
