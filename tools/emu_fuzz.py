"""Differential fuzzing WITHOUT a GPU: the product rasteriser and the UNMODIFIED reference CUDA source, both executed
by the CPU emulator (tests/cuda_emu), on randomised adversarial scenes — huge and tiny splats, edge-on surfels,
opacities at 0 / 1, surfels at the near plane and behind the camera, zero-norm-ish rotations, ragged image sizes,
long per-tile lists, every SH degree, both forks.  Build container only (needs /root/reference).

    python tools/emu_fuzz.py [--fork base|part] [--seeds 0:40]

Both sides run IEEE fp32 without the GPU's FMA contraction but the product pins a few FMAs explicitly, so single
threshold decisions (alpha >= 1/255, T < 1e-4, radius rounding) may flip: integer state is compared with a small
allowance, images / gradients through robust quantiles.  Prints one line per seed and a summary; exit code 1 if any
seed is outside the tolerances."""
import argparse
import ctypes as C
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tests" / "cuda_emu"))
import build as emu_build  # noqa: E402

from partgs_b200 import _lib, synth  # noqa: E402
from test_emu_raster import HostAlloc, _p  # noqa: E402


def f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)


def scene_for(seed, part):
    g = np.random.default_rng(seed)
    P = int(g.integers(40, 400))
    W, H = int(g.integers(8, 72)), int(g.integers(8, 56))
    S = int(g.integers(1, 17)) if part else 0
    sc = synth.make_point_scene(P, seed=seed, S=S, device="cpu")
    cam = synth.make_cameras(1, W, H, seed=seed + 1000, device="cpu")[0]
    m3, scales, rot, op, sh = (f32(sc[k]) for k in ("means3D", "scales", "rotations", "opacities", "shs"))
    mode = seed % 6
    scales *= np.exp(g.uniform(np.log(0.3), np.log(30.0), size=(P, 1))).astype(np.float32)      # tiny .. huge
    if mode == 1:
        scales[:, 1] *= 1e-3                                                                    # needles
    if mode == 2:
        op[g.random(P) < 0.3] = 1.0; op[g.random(P) < 0.1] = 0.0                                # saturated / invisible
    if mode == 3:
        m3 += f32(cam.campos)[None] * g.uniform(0.0, 1.2, size=(P, 1)).astype(np.float32)       # towards / past the camera
    if mode == 4:
        rot *= g.choice([1e-4, 1.0, 1e4], size=(P, 1)).astype(np.float32)                       # badly scaled quaternions
    if mode == 5:
        m3 *= 0.05; scales *= 0.3                                                               # everything in a few tiles
    deg = int(g.integers(0, 4))
    sem = f32(sc["semantics"]) if part else None
    bg = g.random(3).astype(np.float32)
    ng = 8 if part else 7
    gr = synth.upstream_grads(W, H, seed + 7, n_aux=ng, S=S, device="cpu")
    return SimpleNamespace(P=P, W=W, H=H, S=S, deg=deg, m3=m3, scales=scales, rot=rot, op=op, sh=sh, sem=sem, bg=bg,
                           vm=f32(cam.viewmatrix), pm=f32(cam.projmatrix), cp=f32(cam.campos), tx=cam.tanfovx,
                           ty=cam.tanfovy, gc=f32(gr["color"]), ga=f32(gr["allmap"]),
                           gs=f32(gr["semantic"]) if part else None, mode=mode)


def run_reference(ref, s, part):
    P, M, W, H, S = s.P, s.sh.shape[1], s.W, s.H, s.S
    color = np.zeros((3, H, W), np.float32); allmap = np.zeros((8 if part else 7, H, W), np.float32)
    semantic = np.zeros((max(S, 1), H, W), np.float32); radii = np.zeros(P, np.int32)
    names = ["means2D", "normal", "opacity", "colors"] + (["semantics"] if part else []) + ["means3D", "transMat", "sh",
                                                                                           "scales", "rotations"]
    shapes = dict(means2D=(P, 3), normal=(P, 3), opacity=(P, 1), colors=(P, 3), semantics=(P, max(S, 1)),
                  means3D=(P, 3), transMat=(P, 9), sh=(P, M, 3), scales=(P, 2), rotations=(P, 4))
    d = {k: np.zeros(shapes[k], np.float32) for k in names}
    cf = C.c_float
    if part:
        ref.ref_part_forward.restype = C.c_int
        R = ref.ref_part_forward(P, s.deg, M, _p(s.bg), W, H, S, _p(s.m3), _p(s.sh), None, _p(s.op), _p(s.sem),
                                 _p(s.scales), cf(1.0), _p(s.rot), _p(s.vm), _p(s.pm), _p(s.cp), cf(s.tx), cf(s.ty),
                                 _p(color), _p(semantic), _p(allmap), _p(radii))
        if R > 0:
            ref.ref_part_backward(P, s.deg, M, R, _p(s.bg), W, H, S, _p(s.m3), _p(s.sh), None, _p(s.sem), _p(s.scales),
                                  cf(1.0), _p(s.rot), _p(s.vm), _p(s.pm), _p(s.cp), cf(s.tx), cf(s.ty), _p(radii),
                                  _p(s.gc), _p(s.gs), _p(s.ga), *[_p(d[k]) for k in names])
    else:
        ref.ref_base_forward.restype = C.c_int
        R = ref.ref_base_forward(P, s.deg, M, _p(s.bg), W, H, _p(s.m3), _p(s.sh), None, _p(s.op), _p(s.scales), cf(1.0),
                                 _p(s.rot), None, _p(s.vm), _p(s.pm), _p(s.cp), cf(s.tx), cf(s.ty), _p(color),
                                 _p(allmap), _p(radii))
        if R > 0:
            ref.ref_base_backward(P, s.deg, M, R, _p(s.bg), W, H, _p(s.m3), _p(s.sh), None, _p(s.scales), cf(1.0),
                                  _p(s.rot), None, _p(s.vm), _p(s.pm), _p(s.cp), cf(s.tx), cf(s.ty), _p(radii),
                                  _p(s.gc), _p(s.ga), *[_p(d[k]) for k in names])
    return R, color, semantic, allmap, radii, d


def run_product(emu, s, part):
    P, M, W, H, S = s.P, s.sh.shape[1], s.W, s.H, s.S
    color = np.full((3, H, W), np.nan, np.float32); allmap = np.full((8 if part else 7, H, W), np.nan, np.float32)
    semantic = np.full((max(S, 1), H, W), np.nan, np.float32); radii = np.full(P, -7, np.int32)
    al = HostAlloc()
    shapes = dict(means2D=(P, 3), opacity=(P, 1), colors=(P, 3), semantics=(P, max(S, 1)), means3D=(P, 3),
                  transMat=(P, 9), sh=(P, M, 3), scales=(P, 2), rotations=(P, 4))
    g = {k: np.zeros(v, np.float32) for k, v in shapes.items()}
    scratch = np.zeros(emu.pgs_dsr_backward_scratch_bytes(P) + 256, np.uint8)
    sp = (scratch.ctypes.data + 255) // 256 * 256
    if part:
        R = emu.pgs_dsrp_forward(al.cb, 1, al.cb, 2, al.cb, 3, P, s.deg, M, _p(s.bg), W, H, S, _p(s.m3), _p(s.sh), None,
                                 _p(s.sem), _p(s.op), _p(s.scales), 1.0, _p(s.rot), None, _p(s.vm), _p(s.pm), _p(s.cp),
                                 s.tx, s.ty, 0, _p(color), _p(semantic), _p(allmap), _p(radii), 1, None)
        assert R >= 0, emu.pgs_last_error()
        if R > 0:
            rc = emu.pgs_dsrp_backward(P, s.deg, M, R, _p(s.bg), W, H, S, _p(s.m3), _p(s.sh), None, _p(s.sem),
                                       _p(s.scales), 1.0, _p(s.rot), None, _p(s.vm), _p(s.pm), _p(s.cp), s.tx, s.ty,
                                       _p(radii), al.ptr(1), al.ptr(2), al.nbytes(2), al.ptr(3), _p(s.gc), _p(s.gs),
                                       _p(s.ga), _p(g["means2D"]), sp, _p(g["opacity"]), _p(g["colors"]),
                                       _p(g["semantics"]), _p(g["means3D"]), _p(g["transMat"]), _p(g["sh"]),
                                       _p(g["scales"]), _p(g["rotations"]), 1, None)
            assert rc >= 0, emu.pgs_last_error()
    else:
        R = emu.pgs_dsr_forward(al.cb, 1, al.cb, 2, al.cb, 3, P, s.deg, M, _p(s.bg), W, H, _p(s.m3), _p(s.sh), None,
                                _p(s.op), _p(s.scales), 1.0, _p(s.rot), None, _p(s.vm), _p(s.pm), _p(s.cp), s.tx, s.ty,
                                0, _p(color), _p(allmap), _p(radii), 1, None)
        assert R >= 0, emu.pgs_last_error()
        if R > 0:
            rc = emu.pgs_dsr_backward(P, s.deg, M, R, _p(s.bg), W, H, _p(s.m3), _p(s.sh), None, _p(s.scales), 1.0,
                                      _p(s.rot), None, _p(s.vm), _p(s.pm), _p(s.cp), s.tx, s.ty, _p(radii), al.ptr(1),
                                      al.ptr(2), al.nbytes(2), al.ptr(3), _p(s.gc), _p(s.ga), _p(g["means2D"]), sp,
                                      _p(g["opacity"]), _p(g["colors"]), _p(g["means3D"]), _p(g["transMat"]),
                                      _p(g["sh"]), _p(g["scales"]), _p(g["rotations"]), 1, None)
            assert rc >= 0, emu.pgs_last_error()
    return R, color, semantic, allmap, radii, g


def robust(a, b, q=0.995):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    fin = np.isfinite(a) & np.isfinite(b)
    if (np.isfinite(a) != np.isfinite(b)).mean() > 0.002:
        return float("inf")
    if not fin.any():
        return 0.0
    d = np.abs(a[fin] - b[fin])
    return float(np.quantile(d, q) / (np.abs(b[fin]).max() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fork", default="base"); ap.add_argument("--seeds", default="0:24")
    a = ap.parse_args()
    part = a.fork == "part"
    lo, hi = (int(v) for v in a.seeds.split(":"))
    ref = C.CDLL(str(emu_build.build_reference(a.fork)))
    emu = C.CDLL(str(emu_build.build_full()))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(emu, name); fn.restype, fn.argtypes = res, args
    bad = 0
    for seed in range(lo, hi):
        s = scene_for(seed, part)
        Rr, cr, sr, ar, rr, dr = run_reference(ref, s, part)
        Ro, co, so, ao, ro, do = run_product(emu, s, part)
        n_radii = int((rr != ro).sum())
        errs = dict(color=robust(co, cr), allmap=robust(np.delete(ao, 6, 0), np.delete(ar, 6, 0)),
                    dist=robust(ao[6], ar[6], 0.98))
        if part and s.S:
            errs["semantic"] = robust(so[:s.S], sr[:s.S])
        if Rr > 0 and Ro > 0:
            for k in ("means3D", "opacity", "scales", "rotations", "sh") + (("semantics",) if part else ()):
                errs["d_" + k] = robust(do[k], dr[k], 0.99)
        ok = (n_radii <= max(2, s.P // 100) and abs(Ro - Rr) <= max(4, Rr // 200) and errs["color"] <= 1e-4 and
              errs["allmap"] <= 2e-4 and errs["dist"] <= 2e-2 and all(v <= 2e-3 for k, v in errs.items() if k.startswith("d_")) and
              errs.get("semantic", 0) <= 1e-4)
        bad += not ok
        print(f"seed {seed:3d} mode {s.mode} P {s.P:3d} {s.W:2d}x{s.H:2d} S {s.S:2d} deg {s.deg} R {Rr:6d}/{Ro:6d} radii!= {n_radii} "
              + " ".join(f"{k}={v:.1e}" for k, v in errs.items()) + ("" if ok else "   <-- OUTSIDE TOLERANCE"), flush=True)
    print(f"{hi - lo - bad} of {hi - lo} seeds within tolerance")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
